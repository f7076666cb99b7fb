// The per-step pipeline as CUDA kernels for sm_100a.
//
// One model step (main layers 1-7, /root/reference/src/model.py:2140-2163) is
//   k_begin_step    zero the per-step counters
//   k_risk_resolve  sparse: exact death sub-step of the few agents that CAN die in play
//   k_play_travel   dense sweep 1: roll + roles + 8 play sub-steps + mover decision + travel claim; a mover alone
//                   on a target only its tile can reach relocates inside the tile
//   k_move_commit   sparse: resolve the other travel claims of round 1 (one thread per claimant), winners relocate
//   k_move_retry    sparse: travel rounds 2..8 for the claim losers
//   k_claim_punish  dense sweep 2: eligibility + reproduction claim (judged inside the tile where only the tile can
//                   reach the target), fused with cost of living, death, counts and the next step's at-risk list
//                   (a winning parent corrects its own outcome afterwards)
//   k_repro_commit  sparse: resolve the other reproduction claims of round 1; births and retry entries of all
//   k_repro_retry   sparse: reproduction rounds 2..8 (gated by the carrying capacity)
//   k_finalize      sparse: newborn cull (exact k-th selection), the surviving newborns enter the planes
// Every dense sweep is a pure GATHER on the state planes: a cell reads its neighbourhood and writes only itself.
// Claims go through ONE extra plane of 64-bit words with atomic max (the north star's "atomicMin on packed
// (priority, id) per target cell"), tagged by step / phase / round so that it never needs clearing.
// The persistent sparse kernels are templated on their block size: one 1024-thread CTA on small worlds,
// 512-thread CTAs (128 registers available) in the cooperative multi-CTA launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pd_device.cuh"
#include "pd_dense.cuh"

namespace pd {

// ------------------------------------------------------------------------------------------
__global__ void k_set_step(Counters* ctr, unsigned int step) { ctr->step_no = step; }

__global__ void k_begin_step(Counters* ctr, int count_step) {
    // zero everything up to (not including) risk_n
    unsigned int* w = reinterpret_cast<unsigned int*>(ctr);
    const size_t nwords = offsetof(Counters, risk_n) / sizeof(unsigned int);
    for (size_t i = threadIdx.x; i < nwords; i += blockDim.x) w[i] = 0u;
    if (threadIdx.x == 0) {
        ctr->agents_start = ctr->n_agents;
        if (count_step) ctr->agent_steps += ctr->n_agents;
    }
}

// ------------------------------------------------------------------------------------------
// k_risk_resolve: the agents whose energy is so low that a run of negative payoffs could kill
// them during play (list built by the previous step).  A responder that dies in sub-step c
// (:695-697) neither challenges nor responds afterwards, so its neighbours' games depend on
// WHEN it dies.  The result (D code = sub-step of death + 1) is written into the agent's tf byte so
// the dense play sweep knows for every neighbour whether it is still alive in sub-step c.
//   Phase 1, no barriers: an at-risk agent none of whose challengers is itself at risk meets all
//   its challengers for sure, so it plays its sub-steps by itself, all loads of one agent in flight
//   together.
//   Phase 2, the rest (an at-risk challenger may die before it challenges): eight lock-stepped rounds
//   over those few reproduce exit_play_fn's iterations (:1527-1556) exactly.
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS) k_risk_resolve(const __grid_constant__ Params p) {
    Counters* ctr = p.ctr;
    const unsigned int n = ctr->risk_n;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) return;
    uint32_t* dep = p.risk_next;  // scratch: the next step's list is only filled after this kernel
    // Phase 1 pays off when a lock-stepped round would need several entries per thread; with fewer entries than
    // threads (small worlds) the rounds are pure barrier latency and the extra pass only adds to it, so every
    // entry goes straight to phase 2.
    const bool split = n > stride;
    for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
        const unsigned int i = i0 + threadIdx.x;
        bool dependent = false;
        if (i < n) {
            const uint32_t cell = p.risk_cur[i];
            const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
            const uint32_t t = __ldcg(p.tfA + cell);
            if ((t & TF_OCC) && !split) {
                dependent = true;
                p.risk_e[i] = __ldcg(p.E0 + cell);
            } else if (t & TF_OCC) {  // else: a stale entry
                uint32_t nc[8], tn[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    nc[d] = nb_local(p, x, y, d);
                    tn[d] = __ldcg(p.tfA + nc[d]);
                }
                float e = __ldcg(p.E0 + cell);
                const uint32_t my_strat = __ldcg(p.strat + cell);
                const uint64_t g = gcell_of(p, x, y);
                const uint32_t roll = play_word(p, g), coins = roll;
                uint32_t chal = 0;  // bit d: my neighbour d challenges me (:413), in sub-step 7-d
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    if (!(tn[d] & TF_OCC)) continue;
                    if (roll_higher(play_word(p, nb_gcell(p, x, y, d)), roll, d)) chal |= 1u << d;
                }
#pragma unroll
                for (int d = 0; d < 8; ++d)
                    if (((chal >> d) & 1u) && __ldcg(p.E0 + nc[d]) <= p.risk_thr) dependent = true;
                if (dependent) {
                    p.risk_e[i] = e;
                } else {
                    for (int c = 0; c < 8; ++c) {
                        const int d = 7 - c;
                        if (!((chal >> d) & 1u)) continue;
                        e = __fadd_rn(e, game_payoff(p, g, c, coins, my_strat, t & TF_TRAIT, __ldcg(p.strat + nc[d]), tn[d] & TF_TRAIT));
                        if (e <= 0.0f) {
                            p.tfA[cell] = (uint8_t)(t | ((uint32_t)(c + 1) << TF_D_SHIFT));
                            break;
                        }
                        e = fminf(e, p.max_energy);
                    }
                }
            }
        }
        const uint32_t slot = warp_append(&ctr->dep_n, dependent);
        if (dependent) dep[slot] = i;  // dep_n <= n <= list_cap
    }
    grid_sync(ctr->barrier, gridDim.x);
    const unsigned int ndep = __ldcg(&ctr->dep_n);
    if (ndep == 0) return;
    for (int c = 0; c < 8; ++c) {
        const int d = 7 - c;  // my direction towards whoever challenges me in sub-step c (:477)
        for (unsigned int k = t0; k < ndep; k += stride) {
            const unsigned int i = __ldcg(&dep[k]);
            const uint32_t cell = p.risk_cur[i];
            const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
            const uint32_t t = __ldcg(p.tfA + cell);
            if (t & TF_D) continue;  // already dead
            const uint32_t nc = nb_local(p, x, y, d);
            const uint32_t tn = __ldcg(p.tfA + nc);
            if (!(tn & TF_OCC)) continue;
            const uint64_t g = gcell_of(p, x, y);
            const uint32_t wr = play_word(p, g);  // (recomputed: cheaper than a list round trip for these few)
            if (!roll_higher(play_word(p, nb_gcell(p, x, y, d)), wr, d)) continue;  // I am the challenger (:413)
            const uint32_t dn = (tn & TF_D) >> TF_D_SHIFT;
            if (dn != 0u && (int)dn - 1 < c) continue;  // challenger died in an earlier sub-step
            const float e = __fadd_rn(p.risk_e[i], game_payoff(p, g, c, wr, p.strat[cell], t & TF_TRAIT, p.strat[nc], tn & TF_TRAIT));
            if (e <= 0.0f) p.tfA[cell] = (uint8_t)(t | ((uint32_t)(c + 1) << TF_D_SHIFT));
            else p.risk_e[i] = fminf(e, p.max_energy);
        }
        grid_sync(ctr->barrier, gridDim.x);
    }
}

// ------------------------------------------------------------------------------------------
// Claim resolution (sparse).  The dense sweeps post round 1's claims on the claim plane and list the claimants;
// a claimant has won iff its target's word is still its own (pd_device.cuh claim_word).
//   k_move_commit / k_repro_commit   round 1: independent threads over the listed claimants
//   k_move_retry / k_repro_retry     rounds 2..8 for the losers (~5 % of the movers, ~1/3 of the attempting
//                                    parents -- and at saturation the population gate stops reproduction after
//                                    round 1), persistent, two phases per round separated by grid barriers:
//       request  (:784-872, :1042-1134)  probe from the saved direction, post the claim
//       response (:881-960, :1170-1205)  winners commit, losers stay on the list
// Everything happens in place on tfB / E1 / strat: a winner only writes its own cell and its (empty) target,
// nobody reads either in the same phase.
// Retry-list entry state: 0x80 active | 0x40 fresh (claimed in round 1) | (probes used - 1) << 3 | start direction.
// ------------------------------------------------------------------------------------------
// A CTA takes chunks of COMMIT_CHUNK consecutive list entries, COMMIT_K per thread with all their loads in flight
// together, stages what it appends to the global lists (losers, newborns) in shared memory and reserves the
// space with ONE atomic per chunk and list: a warp-level append per entry makes ~10^6 atomics per step on one
// address, which the L2 serialises (that alone was 2/3 of k_repro_commit's time, profiles/README.md).
#ifndef PD_COMMIT_K
#define PD_COMMIT_K 4
#endif
constexpr int COMMIT_NT = 256, COMMIT_K = PD_COMMIT_K, COMMIT_CHUNK = COMMIT_NT * COMMIT_K;

// k_move_commit: move_response (:881-960), round 1.  A winner relocates -- energy, strategies and trait (they
// came with the list entry) go to the target cell -- and leaves a GHOST that keeps blocking movers this step
// (quirk B-13); a loser joins the retry list with the direction it claimed.  One random read per claimant (the
// claim word), four random writes per winner.
__global__ void __launch_bounds__(COMMIT_NT) k_move_commit(const __grid_constant__ Params p) {
    __shared__ uint32_t s_lost_cell[COMMIT_CHUNK];
    __shared__ uint8_t s_lost_state[COMMIT_CHUNK];
    __shared__ unsigned int s_n[2];  // 0 losers of this chunk, 1 their base in the retry list
    Counters* ctr = p.ctr;
    unsigned int n = ctr->mv_n;
    if (n > p.list_cap) n = p.list_cap;
    const uint32_t step = __ldg(&ctr->step_no);
    const int tid = threadIdx.x;
    unsigned int moved = 0;
    for (unsigned int c0 = blockIdx.x * COMMIT_CHUNK; c0 < n; c0 += gridDim.x * COMMIT_CHUNK) {
        if (tid == 0) s_n[0] = 0;
        __syncthreads();
        uint32_t cell[COMMIT_K], tgt[COMMIT_K], aux[COMMIT_K];
        unsigned long long top[COMMIT_K];
        float e[COMMIT_K];
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            const unsigned int i = c0 + k * COMMIT_NT + tid;
            cell[k] = i < n ? p.cl_cell[i] : 0u;
            aux[k] = i < n ? (uint32_t)p.cl_aux[i] : 0u;
            e[k] = i < n ? p.cl_e[i] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            tgt[k] = nb_local(p, cell[k] & (p.W - 1), cell[k] >> p.log2W, (int)(aux[k] & 7u));
            top[k] = __ldcg(p.claim + tgt[k]);
        }
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            const bool valid = c0 + k * COMMIT_NT + tid < n;
            const int d = (int)(aux[k] & 7u);
            const uint32_t x = cell[k] & (p.W - 1), y = cell[k] >> p.log2W;
            const unsigned long long mine = claim_word(step, CLAIM_TRAVEL, rng(p, gcell_of(p, x, y), S_TRAVEL).x, d);
            const bool won = valid && top[k] == mine;
            if (won) {  // :946-956
                p.E1[tgt[k]] = e[k];
                p.strat[tgt[k]] = (uint8_t)(aux[k] >> 5);
                p.tfB[tgt[k]] = (uint8_t)(TF_OCC | ((aux[k] >> 3) & TF_TRAIT));
                p.tfB[cell[k]] = (uint8_t)TF_GHOST;
                if (row_owned(p, tgt[k] >> p.log2W)) ++moved;
            }
            const bool lost = valid && !won;  // :940-943 -> retry with the same roll
            const uint32_t slot = wappend(&s_n[0], lost);
            if (lost) {
                s_lost_cell[slot] = cell[k];
                s_lost_state[slot] = (uint8_t)(0xC0u | (uint32_t)d);
            }
        }
        __syncthreads();
        const unsigned int n_lost = s_n[0];
        if (n_lost) {  // uniform across the CTA
            if (tid == 0) s_n[1] = atomicAdd(&ctr->rt_n[0], n_lost);
            __syncthreads();
            const unsigned int base = s_n[1];
            if (base + n_lost > p.list_cap) {
                if (tid == 0) atomicExch(&ctr->overflow, 1u);
            } else {
                for (unsigned int k = tid; k < n_lost; k += COMMIT_NT) {
                    p.rt_cell[base + k] = s_lost_cell[k];
                    p.rt_state[base + k] = s_lost_state[k];
                }
            }
        }
    }
    if (p.stats) {
        for (int o = 16; o > 0; o >>= 1) moved += __shfl_down_sync(0xFFFFFFFFu, moved, o);
        if ((tid & 31) == 0 && moved) atomicAdd(&ctr->st_moved, (unsigned long long)moved);
    }
}

// k_repro_commit: god_multiply (:1144-1335), round 1.  A winner pays the reproduction cost (parent_outcome: its
// layer-7 result is re-derived from its energy, which came with the list entry) and its child goes on the newborn
// list -- as (cell, direction) only: near the carrying capacity almost every newborn is culled in the same step,
// so the child is made (make_child) in k_finalize if it survives, from the parent's unchanged E1 / strat / tfB
// entries; it blocks its cell when retry rounds must see it.  A child born into a ghost row belongs to the
// neighbour slab: here it only blocks its cell.  One random read per claimant, one random write per winner.
// Two lists: the claimants that went through the claim plane (front of the claim list), and the ones k_claim_punish
// has already judged inside its tiles (LOCAL: back of the list, verdict in the entry, the winner has paid) -- for
// those only the appends are left: no random access at all.
struct CommitShared {
    uint32_t lost_cell[COMMIT_CHUNK];
    uint8_t lost_state[COMMIT_CHUNK];
    uint32_t nb_cell[COMMIT_CHUNK];
    uint8_t nb_dir[COMMIT_CHUNK];
    uint32_t nb_prio[COMMIT_CHUNK];
    unsigned int hist[2048];  // this CTA's share of the cull histogram
    unsigned int n[4];        // 0 losers of this chunk, 1 newborns, 2 / 3 their bases in the global lists
};

template <bool LOCAL>
__device__ __forceinline__ void repro_commit_list(const Params& p, CommitShared& s, unsigned int n, uint32_t step) {
    Counters* ctr = p.ctr;
    const int tid = threadIdx.x;
    for (unsigned int c0 = blockIdx.x * COMMIT_CHUNK; c0 < n; c0 += gridDim.x * COMMIT_CHUNK) {
        if (tid < 2) s.n[tid] = 0;
        __syncthreads();
        uint32_t cell[COMMIT_K], tgt[COMMIT_K], aux[COMMIT_K];
        unsigned long long top[COMMIT_K];
        float e[COMMIT_K];
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            const unsigned int i = c0 + k * COMMIT_NT + tid;
            const unsigned int at = LOCAL ? p.list_cap - 1u - i : i;
            cell[k] = i < n ? p.cl_cell[at] : 0u;
            aux[k] = i < n ? (uint32_t)p.cl_aux[at] : 0u;
            e[k] = !LOCAL && i < n ? p.cl_e[at] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            tgt[k] = nb_local(p, cell[k] & (p.W - 1), cell[k] >> p.log2W, (int)(aux[k] & 7u));
            top[k] = LOCAL ? 0ull : __ldcg(p.claim + tgt[k]);
        }
#pragma unroll
        for (int k = 0; k < COMMIT_K; ++k) {
            const bool valid = c0 + k * COMMIT_NT + tid < n;
            const int d = (int)(aux[k] & 7u);
            const uint32_t x = cell[k] & (p.W - 1), y = cell[k] >> p.log2W;
            bool won, born = false;
            if (LOCAL) {
                won = valid && (aux[k] & 8u) != 0u;
                born = won && row_owned(p, tgt[k] >> p.log2W);
            } else {
                const unsigned long long mine = claim_word(step, CLAIM_REPRO, rng(p, gcell_of(p, x, y), S_REPRO).x, d);
                won = valid && top[k] == mine;
                if (won) {  // REPRODUCTION_COMPLETE (:1211-1213, :1331)
                    const uint32_t trait = (aux[k] >> 3) & TF_TRAIT;
                    if (row_owned(p, y)) parent_outcome(p, cell[k], e[k], trait);
                    born = row_owned(p, tgt[k] >> p.log2W);
                    if (!born) p.tfB[tgt[k]] = (uint8_t)(TF_OCC | TF_NEWBORN | trait);
                }
            }
            const uint32_t bslot = wappend(&s.n[1], born);
            if (born) {
                const uint32_t prio = cull_prio(p, tgt[k]);
                atomicAdd(&s.hist[prio >> 21], 1u);
                s.nb_cell[bslot] = tgt[k];
                s.nb_dir[bslot] = (uint8_t)d;
                s.nb_prio[bslot] = prio;
            }
            const bool lost = valid && !won;  // stays ATTEMPTING_REPRODUCTION (:1203-1205)
            const uint32_t lslot = wappend(&s.n[0], lost);
            if (lost) {
                s.lost_cell[lslot] = cell[k];
                s.lost_state[lslot] = (uint8_t)(0xC0u | (uint32_t)d);
            }
        }
        __syncthreads();
        const unsigned int n_lost = s.n[0], n_born = s.n[1];
        if (tid == 0 && n_lost) s.n[2] = atomicAdd(&ctr->rt_n[1], n_lost);
        if (tid == 32 && n_born) s.n[3] = atomicAdd(&ctr->nb_n, n_born);
        __syncthreads();
        if (n_lost) {  // uniform across the CTA
            const unsigned int base = s.n[2];
            if (base + n_lost > p.list_cap) {
                if (tid == 0) atomicExch(&ctr->overflow, 1u);
            } else {
                for (unsigned int k = tid; k < n_lost; k += COMMIT_NT) {
                    p.rt_cell[base + k] = s.lost_cell[k];
                    p.rt_state[base + k] = s.lost_state[k];
                }
            }
        }
        if (n_born) {
            const unsigned int base = s.n[3];
            if (base + n_born > p.list_cap) {
                if (tid == 0) atomicExch(&ctr->overflow, 1u);
            } else {
                for (unsigned int k = tid; k < n_born; k += COMMIT_NT) {
                    p.nb_cell[base + k] = s.nb_cell[k];
                    p.nb_dir[base + k] = s.nb_dir[k];
                    p.nb_prio[base + k] = s.nb_prio[k];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(COMMIT_NT) k_repro_commit(const __grid_constant__ Params p) {
    __shared__ CommitShared s;
    Counters* ctr = p.ctr;
    unsigned int n = ctr->rp_n, n_loc = ctr->lc_n;
    if (n > p.list_cap) n = p.list_cap;
    if (n_loc > p.list_cap) n_loc = p.list_cap;
    const uint32_t step = __ldg(&ctr->step_no);
    const int tid = threadIdx.x;
    if (blockIdx.x * COMMIT_CHUNK >= n && blockIdx.x * COMMIT_CHUNK >= n_loc) return;
    // the two ends of the claim list have met: entries were overwritten
    if (blockIdx.x == 0 && tid == 0 && (unsigned long long)n + n_loc > p.list_cap) atomicExch(&ctr->overflow, 1u);
    for (int i = tid; i < 2048; i += COMMIT_NT) s.hist[i] = 0;
    repro_commit_list<false>(p, s, n, step);
    repro_commit_list<true>(p, s, n_loc, step);
    __syncthreads();
    for (int i = tid; i < 2048; i += COMMIT_NT)
        if (s.hist[i]) atomicAdd(&ctr->hist[0][0] + i, s.hist[i]);
}

// ------------------------------------------------------------------------------------------
// k_move_retry: travel rounds 2..8 (exit_move_fn, :1559-1576) for the losers of round 1, on tfB / E1 / strat.
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS) k_move_retry(const __grid_constant__ Params p) {
    Counters* ctr = p.ctr;
    unsigned int n = ctr->rt_n[0];
    if (n > p.list_cap) n = p.list_cap;
    if (n == 0) return;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t step = __ldg(&ctr->step_no);
    uint8_t* tf = p.tfB;
    unsigned int moved = 0;
    for (int round = 1; round < 8; ++round) {
        // ---- request: blocked = OCC|GHOST (a cell vacated this step keeps blocking movers, B-13)
        for (unsigned int i = t0; i < n; i += stride) {
            const uint32_t st = p.rt_state[i];
            if (!(st & 0x80u)) continue;
            const uint32_t cell = p.rt_cell[i];
            const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
            uint32_t blocked = 0;
#pragma unroll
            for (int d = 0; d < 8; ++d)
                if (__ldcg(tf + nb_local(p, x, y, d)) & (TF_OCC | TF_GHOST)) blocked |= 1u << d;
            int l = (int)(st & 7u), seq = (int)((st >> 3) & 7u) + 1;
            const Words w = rng(p, gcell_of(p, x, y), S_TRAVEL);  // the roll of the first attempt is kept (:794-807)
            if (st & 0x40u) {  // fresh from round 1: it claimed direction l after probe_pos probes (:848-849)
                seq = probe_pos((int)(w.y >> 29), l);
                l = (l + 1) & 7;
            }
            const int dir = probe(blocked, l, seq);
            if (dir < 0) {
                p.rt_state[i] = 0;  // no free space: stays put (:842-846)
            } else {
                post_claim(p.claim + nb_local(p, x, y, dir), claim_word(step, CLAIM_TRAVEL + (uint32_t)round, w.x, dir));
                p.rt_state[i] = (uint8_t)(0x80u | (uint32_t)((dir + 1) & 7) | ((uint32_t)(seq - 1) << 3));
                p.rt_aux[i] = (uint8_t)dir;
            }
        }
        grid_sync(ctr->barrier, gridDim.x);
        // ---- response
        unsigned int still = 0;
        for (unsigned int i = t0; i < n; i += stride) {
            if (!(p.rt_state[i] & 0x80u)) continue;
            const uint32_t cell = p.rt_cell[i];
            const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
            const int d = (int)(p.rt_aux[i] & 7u);
            const uint32_t tgt = nb_local(p, x, y, d);
            const unsigned long long top = __ldcg(p.claim + tgt);
            const unsigned long long mine = claim_word(step, CLAIM_TRAVEL + (uint32_t)round, rng(p, gcell_of(p, x, y), S_TRAVEL).x, d);
            if (top == mine) {  // the winner relocates (:946-956)
                const uint32_t t = __ldcg(tf + cell);
                p.E1[tgt] = __ldcg(p.E1 + cell);
                p.strat[tgt] = __ldcg(p.strat + cell);
                tf[tgt] = (uint8_t)(TF_OCC | (t & TF_TRAIT));
                tf[cell] = (uint8_t)TF_GHOST;
                p.rt_state[i] = 0;
                if (row_owned(p, tgt >> p.log2W)) ++moved;
            } else {
                ++still;
            }
        }
        if (still) atomicAdd(&ctr->mv_active[round], still);
        grid_sync(ctr->barrier, gridDim.x);
        if (__ldcg(&ctr->mv_active[round]) == 0u) break;
    }
    if (p.stats && moved) atomicAdd(&ctr->st_moved, (unsigned long long)moved);
}

// ------------------------------------------------------------------------------------------
// k_repro_retry: reproduction rounds 2..8 (exit_god_fn, :1607-1627), only while the population (olds + births
// so far) does not exceed max_agents (:1616, :1621-1625).  On tfB; E1 is read-only (a winner's layer-7 outcome
// is re-derived by parent_outcome).  Rounds [round0, round1) run in this launch (a row slab runs one
// round per launch: the caller sums the population over the slabs in between).
// ------------------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(BS) k_repro_retry(const __grid_constant__ Params p) {
    Counters* ctr = p.ctr;
    unsigned int n = ctr->rt_n[1];
    if (n > p.list_cap) n = p.list_cap;
    if (n == 0) return;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t step = __ldg(&ctr->step_no);
    uint8_t* tf = p.tfB;
    for (int round = (int)p.round0; round < (int)p.round1; ++round) {
        // overpopulated gate, evaluated after every round (:1616, :1621-1625); in slab mode the caller
        // has summed olds and births over all slabs into glb[] before this launch
        const unsigned long long pop = p.glb ? __ldcg(&p.glb[0]) + __ldcg(&p.glb[1])
                                             : __ldcg(&ctr->n_old) + (unsigned long long)__ldcg(&ctr->nb_n);
        if (pop > p.max_agents) break;
        if (round == 1) {
            // the newborns of round 1 are only on the newborn list so far: they must block their cells now
            unsigned int nb1 = __ldcg(&ctr->nb_n);
            if (nb1 > p.list_cap) nb1 = p.list_cap;
            for (unsigned int i = t0; i < nb1; i += stride) {
                const uint32_t c = p.nb_cell[i];
                tf[c] = (uint8_t)(TF_OCC | TF_NEWBORN | (__ldcg(tf + parent_of(p, c, p.nb_dir[i])) & TF_TRAIT));
            }
            if (t0 == 0) ctr->nb_flagged = 1u;
            grid_sync(ctr->barrier, gridDim.x);
        }
        // ---- go forth (:1042-1134)
        for (unsigned int i = t0; i < n; i += stride) {
            const uint32_t st = p.rt_state[i];
            if (!(st & 0x80u)) continue;
            const uint32_t cell = p.rt_cell[i];
            const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
            uint32_t blocked = 0;
#pragma unroll
            for (int d = 0; d < 8; ++d)
                if (__ldcg(tf + nb_local(p, x, y, d)) & TF_OCC) blocked |= 1u << d;
            int l = (int)(st & 7u), seq = (int)((st >> 3) & 7u) + 1;
            const Words w = rng(p, gcell_of(p, x, y), S_REPRO);  // the roll of the first attempt is kept (:1074-1076)
            if (st & 0x40u) {  // fresh from round 1: it claimed direction l after probe_pos probes (:1110-1111)
                seq = probe_pos((int)(w.y >> 29), l);
                l = (l + 1) & 7;
            }
            const int dir = probe(blocked, l, seq);  // also covers reproduce_sequence >= 8 (:1067)
            if (dir < 0) {
                p.rt_state[i] = 0x40u;  // REPRODUCTION_IMPOSSIBLE (:1104-1108)
            } else {
                post_claim(p.claim + nb_local(p, x, y, dir), claim_word(step, CLAIM_REPRO + (uint32_t)round, w.x, dir));
                p.rt_state[i] = (uint8_t)(0x80u | (uint32_t)((dir + 1) & 7) | ((uint32_t)(seq - 1) << 3));
                p.rt_aux[i] = (uint8_t)dir;
            }
        }
        grid_sync(ctr->barrier, gridDim.x);
        // ---- multiply (:1170-1335)
        unsigned int still = 0;
        for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride) {
            const unsigned int i = i0 + threadIdx.x;
            bool born = false;
            uint32_t tgt = 0, dir = 0;
            if (i < n && (p.rt_state[i] & 0x80u)) {
                const uint32_t cell = p.rt_cell[i];
                const uint32_t x = cell & (p.W - 1), y = cell >> p.log2W;
                const int d = (int)(p.rt_aux[i] & 7u);
                tgt = nb_local(p, x, y, d);
                const uint64_t g = gcell_of(p, x, y);
                const unsigned long long top = __ldcg(p.claim + tgt);
                const unsigned long long mine = claim_word(step, CLAIM_REPRO + (uint32_t)round, rng(p, g, S_REPRO).x, d);
                if (top == mine) {
                    const uint32_t t = __ldcg(tf + cell);
                    if (row_owned(p, y)) parent_outcome(p, cell, __ldcg(p.E1 + cell), t & TF_TRAIT);  // :1211-1213
                    tf[tgt] = (uint8_t)(TF_OCC | TF_NEWBORN | (t & TF_TRAIT));  // blocks its cell in later rounds
                    dir = (uint32_t)d;  // (REPRODUCTION_COMPLETE: it leaves the list -- one child per step, B-6)
                    born = row_owned(p, tgt >> p.log2W);  // a child in the ghost zone belongs to the neighbour slab
                    p.rt_state[i] = 0x40u;
                } else {
                    ++still;
                }
            }
            const uint32_t slot = warp_append(&ctr->nb_n, born);
            if (born) {
                if (slot < p.list_cap) {
                    const uint32_t prio = cull_prio(p, tgt);
                    atomicAdd(&ctr->hist[0][0] + (prio >> 21), 1u);
                    p.nb_cell[slot] = tgt;
                    p.nb_dir[slot] = (uint8_t)dir;
                    p.nb_prio[slot] = prio;
                } else atomicExch(&ctr->overflow, 1u);
            }
        }
        if (still) atomicAdd(&ctr->rp_active[round], still);
        grid_sync(ctr->barrier, gridDim.x);
        if (__ldcg(&ctr->rp_active[round]) == 0u) break;
    }
}

// ------------------------------------------------------------------------------------------
// Finalisation (sparse, persistent): the carrying-capacity cull, the living cost of the few claim
// losers it kills (deferred by k_repro_commit), the newborns' counts and the bookkeeping for the next step.
// Cull (replaces the list-index rule :1343/:1354, SURVEY B-4): if olds + births exceed
// max_agents, only the K = max_agents - olds newborns with the highest (Philox cull priority,
// cell) keys survive.  The K-th largest 64-bit key is found exactly by a radix select: one
// 2048-bin histogram of the top 11 bits over all newborns, compaction of the threshold bin's
// few candidates, then (CTA 0) one 11-bit pass over the candidates and a rank count of what is left.
// Whole world: one kernel (k_finalize).  Row slabs: three kernels (k_fin_hist, k_fin_cand,
// k_fin_apply) with the caller summing the histogram and gathering the candidates in between.
// ------------------------------------------------------------------------------------------
// exit_god_fn's decision after round 1 (:1614-1627), from the words the caller summed over the slabs: another
// round is due while the population is within the cap and somebody is still trying
__device__ __forceinline__ bool gate_open(const Params& p) {
    return __ldcg(&p.glb[0]) + __ldcg(&p.glb[1]) <= p.max_agents && __ldcg(&p.glb[2]) > 0ull;
}

struct CullPlan {
    bool cull;                 // olds + births exceed the cap
    unsigned long long keep;   // newborns allowed to survive (all slabs together)
};
__device__ __forceinline__ CullPlan cull_plan(const Params& p) {
    const unsigned long long n_old = p.glb ? __ldcg(&p.glb[0]) : p.ctr->n_old;
    const unsigned long long births = p.glb ? __ldcg(&p.glb[1]) : (unsigned long long)p.ctr->nb_n;
    CullPlan c;
    c.cull = n_old + births > p.max_agents;
    c.keep = 0;
    if (c.cull) {
        c.keep = p.max_agents > n_old ? p.max_agents - n_old : 0ull;
        if (c.keep >= births) c.cull = false;
    }
    return c;
}

// Largest bin b such that the number of entries in bins >= b reaches `keep`, found by warp 0 in parallel
// (each lane owns a contiguous chunk of bins, a warp scan finds the chunk, its lane finishes serially).
// Writes the bin to *s_digit and keep - (entries in bins > b) to *s_krem.  Call with all threads; nbins % 32 == 0.
__device__ __forceinline__ void find_threshold_bin(const unsigned int* s_hist, int nbins, unsigned long long keep,
                                                   unsigned int* s_digit, unsigned long long* s_krem) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x, per = nbins / 32;
        // lane 0 owns the TOP chunk so that a prefix scan over lanes is a suffix scan over bins
        const int hi = nbins - 1 - lane * per;
        unsigned long long mine = 0;
        for (int j = 0; j < per; ++j) mine += s_hist[hi - j];
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned int reach = __ballot_sync(0xFFFFFFFFu, incl >= keep);
        const int owner = reach ? __ffs(reach) - 1 : 31;  // first chunk (from the top) where the count is reached
        if (lane == owner) {
            unsigned long long cum = incl - mine;
            int digit = hi - per + 1;  // lowest bin of my chunk if never reached (cannot happen for keep <= total)
            for (int j = 0; j < per; ++j) {
                const unsigned long long h = s_hist[hi - j];
                if (cum + h >= keep) { digit = hi - j; break; }
                cum += h;
            }
            *s_digit = (unsigned int)digit;
            *s_krem = keep - cum;
        }
    }
    __syncthreads();
}

// threshold bin of the (summed) histogram; every CTA computes it identically.  Returns the bin;
// *krem = how many of the bin's candidates survive.
__device__ __forceinline__ unsigned int fin_digit(const unsigned int* hist, unsigned long long keep,
                                                  unsigned int* s_hist, unsigned int* s_digit,
                                                  unsigned long long* s_krem) {
    for (unsigned int i = threadIdx.x; i < 2048; i += blockDim.x) s_hist[i] = __ldcg(&hist[i]);
    __syncthreads();
    find_threshold_bin(s_hist, 2048, keep, s_digit, s_krem);
    return *s_digit;
}

// krem-th largest of `n` candidate keys (all with the same top 11 bits `digit`), by CTA 0 only.  One pass
// histograms the next 11 bits (s_hist has 2048 words); the threshold bin then holds n/2048 keys on average:
// they are compacted into shared memory and the answer is the key with exactly krem-1 larger ones (keys are
// unique: the cell index is in the low bits).  Only if that bin is improbably full do 8-bit passes over all
// candidates finish the job.  key_at(k) fetches candidate k.
template <typename KeyAt>
__device__ __forceinline__ unsigned long long fin_select(unsigned int n, unsigned int digit, unsigned int* s_hist,
                                                         unsigned long long* s_prefix, unsigned long long* s_krem,
                                                         unsigned int smem_cap, KeyAt key_at) {
    constexpr unsigned int SK = 1024;
    __shared__ unsigned long long s_keys[SK];
    __shared__ unsigned int s_m;
    __shared__ unsigned int s_dg_store;
    unsigned int* s_dg = &s_dg_store;
    for (unsigned int i = threadIdx.x; i < 2048; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    for (unsigned int k = threadIdx.x; k < n; k += 4 * blockDim.x) {  // 4 independent loads in flight
        unsigned long long key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) key[u] = (k + u * blockDim.x < n) ? key_at(k + u * blockDim.x) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (k + u * blockDim.x < n && (unsigned int)(key[u] >> 53) == digit)
                atomicAdd(&s_hist[(unsigned int)(key[u] >> 42) & 2047u], 1u);
    }
    __syncthreads();
    find_threshold_bin(s_hist, 2048, *s_krem, s_dg, s_krem);
    const unsigned long long prefix22 = ((unsigned long long)digit << 11) | (unsigned long long)*s_dg;
    const unsigned int m = s_hist[*s_dg];
    if (m <= (smem_cap < SK ? smem_cap : SK)) {  // smem_cap < SK only in tests of the fallback below
        for (unsigned int k = threadIdx.x; k < n; k += 4 * blockDim.x) {
            unsigned long long key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) key[u] = (k + u * blockDim.x < n) ? key_at(k + u * blockDim.x) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u * blockDim.x < n && (key[u] >> 42) == prefix22) s_keys[atomicAdd(&s_m, 1u)] = key[u];
        }
        __syncthreads();
        const unsigned long long krem = *s_krem;
        for (unsigned int i = threadIdx.x; i < m; i += blockDim.x) {
            const unsigned long long mine = s_keys[i];
            unsigned long long larger = 0;
            for (unsigned int j = 0; j < m; ++j) larger += s_keys[j] > mine ? 1ull : 0ull;
            if (larger + 1ull == krem) *s_prefix = mine;
        }
        __syncthreads();
        return *s_prefix;
    }
    if (threadIdx.x == 0) *s_prefix = prefix22;  // the 22 bits decided so far
    __syncthreads();
    int decided = 22;
    while (decided < 64) {
        const int bits = (64 - decided) >= 8 ? 8 : (64 - decided);
        const int shift = 64 - decided - bits;
        if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        for (unsigned int k = threadIdx.x; k < n; k += 4 * blockDim.x) {  // 4 independent loads in flight
            unsigned long long key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) key[u] = (k + u * blockDim.x < n) ? key_at(k + u * blockDim.x) : ~0ull;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u * blockDim.x < n && (key[u] >> (shift + bits)) == prefix)
                    atomicAdd(&s_hist[(unsigned int)(key[u] >> shift) & ((1u << bits) - 1u)], 1u);
        }
        __syncthreads();
        if (bits == 8) {
            find_threshold_bin(s_hist, 256, *s_krem, s_dg, s_krem);
            if (threadIdx.x == 0) *s_prefix = (prefix << 8) | (unsigned long long)*s_dg;
        } else if (threadIdx.x == 0) {  // the last, narrower digit
            unsigned long long krem = *s_krem, cum = 0;
            int dg = 0;
            for (int b = (1 << bits) - 1; b >= 0; --b) {
                const unsigned long long h = s_hist[b];
                if (cum + h >= krem) { dg = b; break; }
                cum += h;
            }
            *s_krem = krem - cum;
            *s_prefix = (prefix << bits) | (unsigned long long)dg;
        }
        __syncthreads();
        decided += bits;
    }
    return *s_prefix;
}

// candidates of the threshold bin `digit`: their keys are appended to out[0 .. cap) through *counter.  The list
// walk keeps four independent loads in flight per thread (these kernels run a few CTAs per SM over ~10^7 entries:
// one load per round trip is pure latency)
__device__ __forceinline__ void fin_candidates(const Params& p, unsigned int nb, unsigned int digit,
                                               unsigned int* counter, unsigned long long* out, unsigned int cap) {
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int i0 = blockIdx.x * blockDim.x * 4u; i0 < nb; i0 += stride * 4u) {
        uint32_t prio[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned int i = i0 + u * blockDim.x + threadIdx.x;
            prio[u] = i < nb ? p.nb_prio[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned int i = i0 + u * blockDim.x + threadIdx.x;
            const bool is_cand = i < nb && (prio[u] >> 21) == digit;
            const uint32_t slot = warp_append(counter, is_cand);
            if (is_cand) {
                if (slot < cap) out[slot] = cull_key_of(p, prio[u], p.nb_cell[i]);
                else atomicExch(&p.ctr->overflow, 1u);  // > cap newborns in one of 2048 bins: not a real world
            }
        }
    }
}

// newborns: cull or admit (OWNED rows only; no living cost in the birth step, :1343).  With a cull only the
// newborns of the bins >= `digit` can survive, so the walk reads the priorities alone (four loads in flight).
__device__ __forceinline__ void fin_apply(const Params& p, unsigned int nb, bool cull, unsigned long long keep,
                                          unsigned long long tau, unsigned int digit) {
    Counters* ctr = p.ctr;
    const unsigned int stride = gridDim.x * blockDim.x;
    unsigned int n_admitted = 0;
    for (unsigned int i0 = blockIdx.x * blockDim.x * 4u; i0 < nb; i0 += stride * 4u) {
        uint32_t prio[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned int i = i0 + u * blockDim.x + threadIdx.x;
            prio[u] = i < nb ? p.nb_prio[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned int i = i0 + u * blockDim.x + threadIdx.x;
            bool at_risk = false;
            uint32_t cell = 0;
            if (i < nb && (!cull || (keep != 0 && (prio[u] >> 21) >= digit))) {
                cell = p.nb_cell[i];
                if (!cull || cull_key_of(p, prio[u], cell) >= tau) {
                    // Only now is the child made (:1216-1328: the parent's trait, its own energy and strategies) and
                    // enters the planes.  Its parent's E1 / strat / tfB entries are as they were when it won its claim
                    // (a parent neither moves nor is overwritten during the reproduction rounds, dead or alive).
                    const uint32_t par = parent_of(p, cell, p.nb_dir[i]);
                    const uint32_t trait = (uint32_t)p.tfB[par] & TF_TRAIT;
                    float ce;
                    uint32_t cs;
                    make_child(p, gcell_of(p, par & (p.W - 1), par >> p.log2W), p.strat[par], trait,
                               __fsub_rn(p.E1[par], p.reproduce_cost), ce, cs);
                    p.tfA[cell] = (uint8_t)(TF_OCC | trait);
                    p.E0[cell] = ce;
                    p.strat[cell] = (uint8_t)cs;
                    at_risk = ce <= p.risk_thr;
                    if (p.want_counts) atomicAdd(&ctr->bins[count_bin(cs, trait)], 1ull);
                    ++n_admitted;
                }  // (a culled newborn never entered tfA / E0: nothing to undo)
            }
            const uint32_t slot = warp_append(&ctr->risk_next_n, at_risk);
            if (at_risk) {
                if (slot < p.list_cap) p.risk_next[slot] = cell;
                else atomicExch(&ctr->overflow, 1u);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) n_admitted += __shfl_down_sync(0xFFFFFFFFu, n_admitted, o);  // one atomic per warp
    if ((threadIdx.x & 31) == 0 && n_admitted) atomicAdd(&ctr->admitted, (unsigned long long)n_admitted);
}

// bookkeeping by one thread after everything else of the step is done
__device__ __forceinline__ void fin_bookkeeping(const Params& p) {
    Counters* ctr = p.ctr;
    const unsigned long long ld = __ldcg(&ctr->living_deaths);
    const unsigned long long b = __ldcg(&ctr->nb_n);  // births = the newborn list
    const unsigned long long culled = b - __ldcg(&ctr->admitted);
    ctr->n_agents = __ldcg(&ctr->n_old) + b - culled - ld;  // this slab's agents
    unsigned int rn = __ldcg(&ctr->risk_next_n);
    ctr->risk_n = rn > p.list_cap ? p.list_cap : rn;
    if (p.want_counts)
        for (int k = 0; k < 16; ++k) ctr->last_bins[k] = __ldcg(&ctr->bins[k]);
    unsigned long long* s = ctr->last_stats;
    s[0] = ctr->agents_start; s[1] = ctr->n_agents;
    s[2] = __ldcg(&ctr->st_play_deaths); s[3] = __ldcg(&ctr->st_movers);
    s[4] = __ldcg(&ctr->st_moved); s[5] = __ldcg(&ctr->st_travel_deaths);
    s[6] = b; s[7] = culled; s[8] = ld;
    s[9] = rn; s[10] = __ldcg(&ctr->rt_n[0]); s[11] = __ldcg(&ctr->rt_n[1]);  // losers of round 1 (travel, reproduction)
    ctr->step_no += 1u;  // the step is complete
    if (p.want_counts && p.step_log) {  // one row per counted step, read back whenever the host wants
        const unsigned int row = ctr->log_n;
        if (row < p.log_cap) {
            unsigned long long* r = p.step_log + (size_t)row * 18u;
            r[0] = ctr->step_no;
            r[1] = ctr->n_agents;
            for (int k = 0; k < 16; ++k) r[2 + k] = ctr->last_bins[k];
            ctr->log_n = row + 1u;
        } else {
            ctr->log_dropped += 1u;
        }
    }
    if (p.glb) {  // publish this slab's totals; the caller sums them over the slabs when it wants them
        p.glb[3] = ctr->n_agents;
        for (int k = 0; k < 16; ++k) p.glb[4 + k] = p.want_counts ? ctr->last_bins[k] : 0ull;
    }
}

template <int BS>
__global__ void __launch_bounds__(BS, BS == 512 ? 2 : 1) k_finalize(const __grid_constant__ Params p) {
    Counters* ctr = p.ctr;
    __shared__ unsigned int s_hist[2048];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_krem;
    __shared__ unsigned int s_digit;
    unsigned int nb = ctr->nb_n;
    if (nb > p.list_cap) nb = p.list_cap;
    unsigned int* g_hist = &ctr->hist[0][0];  // 2048 words
    unsigned long long* cand = p.cand;
    const unsigned int cand_cap = p.list_cap / 2u;
    const CullPlan plan = cull_plan(p);
    unsigned long long tau = 0;                // survive iff key >= tau
    unsigned int digit = 0;
    if (plan.cull && plan.keep > 0) {
        // (the histogram of the priorities' top 11 bits was built as the newborns were listed)
        digit = fin_digit(g_hist, plan.keep, s_hist, &s_digit, &s_krem);
        fin_candidates(p, nb, digit, &ctr->cand_n, cand, cand_cap);
        grid_sync(ctr->barrier, gridDim.x);
        if (blockIdx.x == 0) {
            unsigned int nc = __ldcg(&ctr->cand_n);
            if (nc > cand_cap) nc = cand_cap;
            const unsigned long long t = fin_select(nc, digit, s_hist, &s_prefix, &s_krem, p.select_cap,
                                                    [&](unsigned int k) { return __ldcg(&cand[k]); });
            if (threadIdx.x == 0) ctr->tau = t;
        }
        grid_sync(ctr->barrier, gridDim.x);
        tau = __ldcg(&ctr->tau);
    }
    fin_apply(p, nb, plan.cull, plan.keep, tau, digit);
    grid_sync(ctr->barrier, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) fin_bookkeeping(p);
}

// ---- row slabs, part 1: this slab's histogram (built as the newborns were listed) into the bound exchange
// buffer, for the caller to sum over the slabs
__global__ void __launch_bounds__(256) k_fin_hist(const __grid_constant__ Params p) {
    for (unsigned int i = threadIdx.x; i < 2048; i += blockDim.x) p.ghist[i] = (&p.ctr->hist[0][0])[i];
}

// ---- part 2 (after the histogram was summed): this slab's candidate keys into cand_out
template <int BS>
__global__ void __launch_bounds__(BS, BS == 512 ? 2 : 1) k_fin_cand(const __grid_constant__ Params p) {
    __shared__ unsigned int s_hist[2048];
    __shared__ unsigned long long s_krem;
    __shared__ unsigned int s_digit;
    Counters* ctr = p.ctr;
    if (p.speculative && gate_open(p)) return;  // a further reproduction round is due: the caller comes back later
    const CullPlan plan = cull_plan(p);
    if (!(plan.cull && plan.keep > 0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) p.cand_out[0] = 0ull;
        return;
    }
    unsigned int nb = ctr->nb_n;
    if (nb > p.list_cap) nb = p.list_cap;
    const unsigned int digit = fin_digit(p.ghist, plan.keep, s_hist, &s_digit, &s_krem);
    fin_candidates(p, nb, digit, &ctr->cand_n, p.cand_out + 1, p.cand_cap);
    grid_sync(ctr->barrier, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned int n = __ldcg(&ctr->cand_n);
        p.cand_out[0] = n < p.cand_cap ? n : p.cand_cap;
    }
}

// ---- part 3 (after all slabs' candidates were gathered): threshold, apply, bookkeeping
template <int BS>
__global__ void __launch_bounds__(BS, BS == 512 ? 2 : 1) k_fin_apply(const __grid_constant__ Params p) {
    __shared__ unsigned int s_hist[2048];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_krem;
    __shared__ unsigned int s_digit;
    __shared__ unsigned int s_cum[PD_MAX_WORLD + 1];
    Counters* ctr = p.ctr;
    if (p.speculative && gate_open(p)) return;  // (see k_fin_cand)
    unsigned int nb = ctr->nb_n;
    if (nb > p.list_cap) nb = p.list_cap;
    const CullPlan plan = cull_plan(p);
    unsigned long long tau = 0;
    unsigned int digit = 0;
    if (plan.cull && plan.keep > 0) {
        digit = fin_digit(p.ghist, plan.keep, s_hist, &s_digit, &s_krem);
        if (blockIdx.x == 0) {
            // candidates of slab w live at cand_all[w * (1 + cap) + 1 ...]; flatten by a linear search
            // over the (at most 8) block counts
            // exclusive prefix of the block counts in shared memory (a search through global memory costs up
            // to `world` dependent L2 round trips per key: +0.4 ms per step on 8 GPUs)
            const unsigned long long blk = 1ull + p.cand_cap;
            if (threadIdx.x == 0) {
                unsigned int cum = 0;
                for (uint32_t w = 0; w < p.world; ++w) {
                    s_cum[w] = cum;
                    cum += (unsigned int)__ldcg(&p.cand_all[w * blk]);
                }
                s_cum[p.world] = cum;
            }
            __syncthreads();
            const unsigned int total = s_cum[p.world];
            const unsigned long long t = fin_select(total, digit, s_hist, &s_prefix, &s_krem, p.select_cap, [&](unsigned int k) {
                uint32_t w = 0;
                while (k >= s_cum[w + 1]) ++w;
                return __ldcg(&p.cand_all[w * blk + 1 + (k - s_cum[w])]);
            });
            if (threadIdx.x == 0) ctr->tau = t;
        }
        grid_sync(ctr->barrier, gridDim.x);
        tau = __ldcg(&ctr->tau);
    }
    fin_apply(p, nb, plan.cull, plan.keep, tau, digit);
    grid_sync(ctr->barrier, gridDim.x);
    if (blockIdx.x == 0 && threadIdx.x == 0) fin_bookkeeping(p);
}

// ---- row slabs: publish this slab's olds / births / still-active repro losers for the caller to sum
__global__ void k_publish_gate(const __grid_constant__ Params p, int round) {
    if (threadIdx.x == 0) {
        p.glb[0] = p.ctr->n_old;
        p.glb[1] = p.ctr->nb_n;  // births
        p.glb[2] = round == 0 ? p.ctr->rt_n[1] : p.ctr->rp_active[round];  // parents still trying
    }
}

// ---- row slabs: the at-risk agents of the freshly received ghost rows join the current list
__global__ void __launch_bounds__(256) k_ghost_risk_scan(const __grid_constant__ Params p) {
    const size_t W = p.W;
    const size_t n_top = (size_t)p.own_y0 * W, n_bot = (size_t)(p.H - p.own_y1) * W;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x; i0 < n_top + n_bot; i0 += stride) {
        const size_t i = i0 + threadIdx.x;
        bool at_risk = false;
        size_t cell = 0;
        if (i < n_top + n_bot) {
            cell = i < n_top ? i : (size_t)p.own_y1 * W + (i - n_top);
            at_risk = (p.tfA[cell] & TF_OCC) && p.E0[cell] <= p.risk_thr;
        }
        const uint32_t slot = warp_append(&p.ctr->risk_n, at_risk);
        if (at_risk) {
            if (slot < p.list_cap) p.risk_cur[slot] = (uint32_t)cell;
            else atomicExch(&p.ctr->overflow, 1u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_scan_state: rebuild the at-risk list, the population count and the strategy counts from the
// bound planes (after init or after the caller wrote the planes); canonicalises tf.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_scan_state(const __grid_constant__ Params p, size_t cells) {
    __shared__ unsigned int s_bins[16];
    __shared__ unsigned int s_n;
    if (threadIdx.x < 16) s_bins[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    unsigned int n = 0;
    const size_t stride = (size_t)gridDim.x * NT;
    for (size_t i0 = (size_t)blockIdx.x * NT; i0 < cells; i0 += stride) {
        const size_t i = i0 + threadIdx.x;
        bool at_risk = false;
        if (i < cells && row_owned(p, (uint32_t)(i >> p.log2W))) {
            const uint32_t t = p.tfA[i];
            if (t & TF_OCC) {
                ++n;
                p.tfA[i] = (uint8_t)(t & (TF_OCC | TF_TRAIT));
                at_risk = p.E0[i] <= p.risk_thr;
                atomicAdd(&s_bins[count_bin(p.strat[i], t & TF_TRAIT)], 1u);
            } else {
                p.tfA[i] = 0;
                p.E0[i] = 0.0f;
            }
        }
        const uint32_t slot = warp_append(&p.ctr->risk_next_n, at_risk);
        if (at_risk) {
            if (slot < p.list_cap) p.risk_next[slot] = (uint32_t)i;
            else atomicExch(&p.ctr->overflow, 1u);
        }
    }
    if (n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0 && s_n) atomicAdd(&p.ctr->n_old, (unsigned long long)s_n);
    if (threadIdx.x < 16 && s_bins[threadIdx.x])
        atomicAdd(&p.ctr->bins[threadIdx.x], (unsigned long long)s_bins[threadIdx.x]);
}
__global__ void k_scan_finish(Counters* ctr, uint32_t list_cap) {
    if (threadIdx.x == 0) {
        ctr->n_agents = ctr->n_old;
        const unsigned int rn = ctr->risk_next_n;
        ctr->risk_n = rn > list_cap ? list_cap : rn;
        for (int k = 0; k < 16; ++k) ctr->last_bins[k] = ctr->bins[k];
    }
}

// ------------------------------------------------------------------------------------------
// Population init (init_fn, :1423-1524) on the GPU.  Exactly n cells -- those with the n
// smallest (placement priority, cell) keys -- are occupied: an 8-pass radix select over all
// cells finds the n-th smallest key, then one sweep writes the agents.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long place_key(const Params& p, uint64_t g) {
    return ((unsigned long long)rng_init(p, g, I_PLACE).x << 32) | (uint32_t)g;
}
// one histogram pass; prefix/k state lives in ctr->hist bookkeeping done by k_init_pick
__global__ void __launch_bounds__(NT) k_init_hist(const __grid_constant__ Params p, size_t cells, int pass,
                                                   unsigned long long prefix) {
    __shared__ unsigned int s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const size_t stride = (size_t)gridDim.x * NT;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < cells; i += stride) {
        const uint64_t g = (uint64_t)i;  // GLOBAL cell index: every slab scans the whole world and finds the same threshold
        const unsigned long long key = place_key(p, g);
        if (pass == 0 || (key >> (shift + 8)) == prefix)
            atomicAdd(&s_hist[(unsigned int)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&p.ctr->hist[pass][threadIdx.x], s_hist[threadIdx.x]);
}
__global__ void __launch_bounds__(NT) k_init_write(const __grid_constant__ Params p, size_t cells,
                                                    unsigned long long tau, int none) {
    const size_t stride = (size_t)gridDim.x * NT;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < cells; i += stride) {
        const uint64_t g = gcell_of(p, (uint32_t)(i & (p.W - 1)), (uint32_t)(i >> p.log2W));
        const Words w = rng_init(p, g, I_PLACE);
        const unsigned long long key = ((unsigned long long)w.x << 32) | (uint32_t)g;
        uint8_t tf = 0, st = 0;
        float e = 0.0f;
        if (!none && key <= tau) {
            const uint32_t trait = w.y >> 30;  // :1467
            const float z = normal_from_words(rng_init(p, g, I_ENERGY));
            e = __fadd_rn(__fmul_rn(z, p.e_sigma), p.e_mu);  // :1460-1462
            e = fmaxf(e, p.e_min);
            if (p.max_energy > 0.0f) e = fminf(e, p.max_energy);  // :1463-1464
            const Words sw = rng_init(p, g, I_STRAT);
            const uint32_t words[4] = {sw.x, sw.y, sw.z, sw.w};
            uint32_t draw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                draw[k] = (uint32_t)((unsigned long long)words[k] >= p.strat_thr[0]) +
                          (uint32_t)((unsigned long long)words[k] >= p.strat_thr[1]) +
                          (uint32_t)((unsigned long long)words[k] >= p.strat_thr[2]);
            uint32_t s;
            if (p.pure) s = draw[0] * 0x55u;                                   // :1473-1477
            else if (p.per_trait) s = draw[0] | (draw[1] << 2) | (draw[2] << 4) | (draw[3] << 6);  // :1478-1482
            else s = ((draw[1] * 0x55u) & ~(3u << (2 * trait))) | (draw[0] << (2 * trait));       // :1483-1497
            st = (uint8_t)s;
            tf = (uint8_t)(TF_OCC | trait);
        }
        p.tfA[i] = tf;
        p.strat[i] = st;
        p.E0[i] = e;
    }
}

// ---- row slabs: halo exchange through peer memory.  A strip = ghost_rows rows of the three state planes, packed
// as [energy f32 | strat u8 | tf u8].  k_halo_pack copies this slab's first / last owned rows into two strips of
// an outbox the neighbours can read (symmetric memory over NVLink); k_halo_unpack fills the ghost rows from the
// neighbours' strips -- the loads go over NVLink, 16 bytes per thread and iteration.
__global__ void __launch_bounds__(256) k_halo_copy(const __grid_constant__ Params p, uint4* a_strip, uint4* b_strip,
                                                   uint32_t a_row, uint32_t b_row, uint32_t rows, int unpack) {
    // strip a <-> local rows [a_row, a_row + rows), strip b <-> [b_row, b_row + rows)
    const size_t cells = (size_t)rows << p.log2W;
    const size_t ne = cells / 4, nb = cells / 16;   // 16-byte chunks of the energy part / of a byte plane's part
    const size_t per = ne + 2 * nb;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per; i += stride) {
        const bool second = i >= per;
        const size_t k = second ? i - per : i;
        uint4* strip = second ? b_strip : a_strip;
        const size_t row0 = (size_t)(second ? b_row : a_row) << p.log2W;
        uint4* plane;
        if (k < ne) plane = reinterpret_cast<uint4*>(p.E0 + row0) + k;
        else if (k < ne + nb) plane = reinterpret_cast<uint4*>(p.strat + row0) + (k - ne);
        else plane = reinterpret_cast<uint4*>(p.tfA + row0) + (k - ne - nb);
        if (unpack) *plane = strip[k];
        else strip[k] = *plane;
    }
}

// Philox known-answer probe for the tests
__global__ void k_philox_kat(const uint32_t* in, uint32_t* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const Words w = philox4x32_10(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3], in[6 * i + 4], in[6 * i + 5]);
        out[4 * i] = w.x; out[4 * i + 1] = w.y; out[4 * i + 2] = w.z; out[4 * i + 3] = w.w;
    }
}

}  // namespace pd

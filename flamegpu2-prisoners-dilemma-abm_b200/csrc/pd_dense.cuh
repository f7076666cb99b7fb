// The two dense sweeps of a step.
//
//   k_play_travel   tfA, strat (1-halo), E0 -> tfB, E1     search/game list/8 play sub-steps/first move_request:
//                                                          a mover posts its claim on the target's word of the claim
//                                                          plane and goes on the claim list
//   k_claim_punish  tfB (1-halo), E1        -> tfA, E0     neighbourhood_update + first god_go_forth: a parent posts
//                                                          its claim and goes on the claim list; fused with
//                                                          environmental_punishment + counts (cell-local)
//
// Round 1 of the claims (move_response / god_multiply) is decided WHERE IT CAN BE DECIDED: a target cell all of
// whose eight neighbours lie in the claimant's tile -- every cell but the tile's outermost ring, 86 % of them --
// cannot have a rival outside the tile, so such a claim is settled in shared memory inside the sweep that makes it
// (a byte mask per cell for the movers, who relocate within the tile before it is stored; an atomic max per cell
// for the parents, who pay at once).  Only claims on the rim, movers with rivals, and rounds 2..8 go through the
// claim plane and the SPARSE kernels (pd_kernels.cuh): a claimant has won iff its target's claim word is still
// its own (atomic max, pd_device.cuh claim_word), one independent thread per claimant, no barriers.
// (History: round 1 as two more dense sweeps that staged every tile for a 4-5 % claimant rate cost 1.8 ms of a
// 5.3 ms step; all claims through the plane cost 0.6 + 0.4 ms in the commit kernels, bound by random DRAM sectors;
// now 0.18 + 0.16 ms -- profiles/README.md.)
//
// Design rules learned from the profiles:
//   * u8 planes enter shared memory as 16-byte chunks (interior at column 16 of a TW+32 wide row);
//     the common path touches 4 cells per thread as one 32-bit word / one float4;
//   * anything that needs a Philox call (rolls, movers, parents) is first COMPACTED into a shared-memory
//     work list and then processed with all 32 lanes busy;
//   * global list appends are aggregated per CTA (one atomicAdd per list per CTA);
//   * ballots / shuffles in a hot loop, nvcc's warp-aggregated atomicAdd(&shared_counter, 1) in divergent code
//     (use satom_add) and extra __syncthreads() phases with little work in them all cost more than their
//     instruction count suggests.
// Every sweep is a pure gather on the planes: a cell is written by its own tile only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pd_device.cuh"

namespace pd {

constexpr int NT = 256;  // threads per CTA of the dense sweeps
#ifndef PD_PLAY_CTAS
#define PD_PLAY_CTAS 5   // resident CTAs per SM the play sweep is compiled for (register budget)
#endif
#ifndef PD_CLAIM_CTAS
#define PD_CLAIM_CTAS 6   // (6 CTAs at 40 registers 0.87 ms, 8 at 32 registers 0.89 ms)
#endif
constexpr uint32_t FULL = 0xFFFFFFFFu;

// living cost for one agent (environmental_punishment, :1357-1368); returns false if it dies
__device__ __forceinline__ bool punish(const Params& p, float& e) {
    e = __fsub_rn(fminf(e, p.max_energy), p.cost_of_living);
    return e > 0.0f;
}

// Shared-memory atomic add that the compiler leaves alone.  For `atomicAdd(&counter, 1)` in divergent code nvcc
// emits a warp-aggregation sequence (vote, popc, leader election, shuffle: ~20 issue slots per site); with the
// handful of lanes that append to a queue at any one site here, letting the hardware serialise the lanes of one
// ATOMS is cheaper (profiles/README.md).
__device__ __forceinline__ unsigned int satom_add(unsigned int* counter, unsigned int v) {
    unsigned int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                 : "=r"(old)
                 : "r"((unsigned int)__cvta_generic_to_shared(counter)), "r"(v)
                 : "memory");
    return old;
}

__device__ __forceinline__ void satom_or(unsigned int* word, unsigned int v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(word)), "r"(v) : "memory");
}

// warp-aggregated append to a shared-memory counter; all 32 lanes must call
__device__ __forceinline__ uint32_t wappend(unsigned int* counter, bool pred) {
    const unsigned int m = __ballot_sync(FULL, pred);
    if (m == 0u) return 0u;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = satom_add(counter, (unsigned int)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    return base + (unsigned int)__popc(m & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------
// k_play_travel (dense sweep 1).  Fuses search_for_neighbours (:352-372), get_game_list
// (:373-440), the 8 iterations of play_challenge/play_response (:450-734) with exit_play_fn's
// terminal rule, and the first move_request (:784-872).
//
// Per cell of tile + 1-ring ONE 32-bit word R is staged in shared memory:
//     empty cell: bit 31 clear (the rest is 0 or leftover roll / strategy bits)
//     occupied:   bit 31 | roll (16 bits) << 15 | has-D-code << 11 | strategies (8) << 3 | trait (2) << 1
// so one shared-memory load per neighbour answers "occupied?", "does it challenge me?" (an unsigned
// compare of the whole word against one of two thresholds: a tie of the rolls goes to the neighbour in a
// direction >= 4) and yields both strategies of the game.  The words are built straight from global memory,
// four cells (one flag word, one strategy word, ONE Philox call) per thread -- no flag / strategy tiles in
// shared memory and a single barrier before the agents run; the same pass compacts the tile's agents into a
// work list.  An agent then plays its (up to) 8 games as the responder in sub-step order straight from the
// words: the strategies, the two RANDOM coins (the low half of the agent's play word, kept next to its list
// entry) and -- with noise -- the two flips form the index of a 256-entry payoff table in shared memory, so a
// game costs one table look-up and no random call.
// ------------------------------------------------------------------------------------------

constexpr uint32_t R_OCC = 0x80000000u;
constexpr uint32_t R_DFLAG = 0x800u;
constexpr uint32_t R_LOW = 0x7FFFu;  // everything below the roll

// Thread-level append of up to four values (byte k of `occ4` has bit 7 set = v[k] goes on the list): ONE shared
// atomic per thread with its count, then plain stores.  No ballots, and with a per-lane count ptxas leaves the
// atomic alone (for a uniform increment it wraps it in a vote / elect / shuffle sequence of ~15 issue slots).
__device__ __forceinline__ void tappend4(unsigned int* cnt, uint32_t* list, uint32_t occ4, const uint32_t (&v)[4]) {
    if (!occ4) return;
    unsigned int slot = satom_add(cnt, (unsigned int)__popc(occ4));
    if (occ4 & 0x80u) list[slot++] = v[0];
    if (occ4 & 0x8000u) list[slot++] = v[1];
    if (occ4 & 0x800000u) list[slot++] = v[2];
    if (occ4 & 0x80000000u) list[slot] = v[3];
}

// entry i of the payoff table: bits 0-1 the challenger's strategy against my trait, 2-3 my strategy against
// its trait (:599-600), 4 the challenger's RANDOM coin, 5 mine (:631, :664), 6 the challenger's noise flip,
// 7 mine (:638-640, :672-674) -> my payoff (:677-689)
__host__ __device__ __forceinline__ float lut_entry(const float (&pay)[4], uint32_t i) {
    const uint32_t theirs = i & 3u, mine = (i >> 2) & 3u;
    bool ch_coop = theirs == 3u ? ((i >> 4) & 1u) != 0u : theirs != 1u;
    bool i_coop = mine == 3u ? ((i >> 5) & 1u) != 0u : mine != 1u;
    if ((i >> 6) & 1u) ch_coop = !ch_coop;
    if ((i >> 7) & 1u) i_coop = !i_coop;
    return pay[((uint32_t)i_coop << 1) | (uint32_t)ch_coop];
}

// The (up to) 8 games of one agent as the responder, in sub-step order (:677-710).  rn[d] = word of neighbour d
// (0 if it does not play against me), tA / tB = the thresholds a challenger's word exceeds.  Returns the energy
// after the last game; `dead` latches a death (:695-697: what a dead agent's energy does afterwards is never
// used), `any` = at least one game.
template <bool NOISY>
__device__ __forceinline__ float play_games(const Params& p, const float* s_lut, const uint32_t (&rn)[8], uint32_t rme,
                                            uint32_t tA, uint32_t tB, uint32_t coins, uint64_t my_g, float e,
                                            bool& dead, bool& any) {
    // my strategy byte replicated through 64 bits and moved up by 4: a funnel shift by the neighbour's word --
    // whose low 5 bits are 2 * its trait plus a multiple of 8 -- drops my strategy against that trait at bits 4-5
    const uint32_t rep = ((rme >> 3) & 0xFFu) * 0x01010101u;
    const uint32_t xlo = rep << 4, xhi = rep >> 4;  // (rep:rep << 4) = xhi:xlo
    const uint32_t my_sh = (rme & 6u) | 1u;         // the neighbour's strategy against my trait -> bits 2-3
    const char* lut = reinterpret_cast<const char*>(s_lut);
    Words nw = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int d = 7 - c;  // my direction towards whoever challenges me in sub-step c (:477)
        const uint32_t r = rn[d];
        const bool hi = r > (d < 4 ? tA : tB);  // it challenges me (:413)
        const uint32_t m = __funnelshift_r(xlo, xhi, r);
        const uint32_t t = r >> my_sh;
        uint32_t idx = (m & 0x30u) | (t & ~0x30u);
        const uint32_t cs = c < 3 ? coins << (6 - 2 * c) : coins >> (2 * c - 6);
        idx = (cs & 0xC0u) | (idx & ~0xC0u);
        idx &= 0xFCu;
        if (NOISY) {
            if ((c & 1) == 0) nw = rng(p, my_g, S_NOISE0 + (uint32_t)(c >> 1));  // :611-612
            const uint32_t my_roll = (c & 1) ? nw.z : nw.x, ch_roll = (c & 1) ? nw.w : nw.y;
            if ((unsigned long long)ch_roll < p.noise_thr) idx |= 0x100u;  // :638-640
            if ((unsigned long long)my_roll < p.noise_thr) idx |= 0x200u;  // :672-674
        }
        const float en = fminf(__fadd_rn(e, *reinterpret_cast<const float*>(lut + idx)), p.max_energy);  // :679-702
        e = hi ? en : e;
        dead |= e <= 0.0f;  // :695-697 (an agent enters the step with energy > 0)
        any |= hi;
    }
    return e;
}

template <int TW, int TH>
__global__ void __launch_bounds__(NT, PD_PLAY_CTAS) k_play_travel(const __grid_constant__ Params p) {
    // R words: TH + 2 rows of GROW groups of 4 cells -- the groups left and right of the tile hold the ring columns
    constexpr int GROW = TW / 4 + 2, SWR = 4 * GROW, NC = TW * TH;
    constexpr int NI = NC / 4, NITEM = NI + 2 * GROW + 2 * TH;  // interior groups, then the ring's: top, bottom, sides
    constexpr int LOG_TW = TW == 128 ? 7 : 4;
    static_assert(TW == 128 || TW == 16, "tile width");
    static_assert(NI % 32 == 0, "whole warps of interior groups");
    __shared__ __align__(16) uint32_t s_R[(TH + 2) * SWR];
    __shared__ __align__(16) float s_e[NC];
    __shared__ __align__(16) uint8_t s_out[NC];
    // the agents of the tile: tile-local cell index | the agent's 16 coin bits << 16.  Later: the tile's claimants.
    __shared__ uint32_t s_cells[NC];
    __shared__ uint16_t s_mover[NC];
    __shared__ float s_lut[256];
    // travel claims on cells ALL of whose neighbours lie in this tile (everything but the tile's outermost cells):
    // bit j of a cell's byte = "the agent on its neighbour j claims it".  Nobody outside the tile can claim such a
    // cell, so an uncontested claim is settled right here, in shared memory (below)
    __shared__ __align__(16) uint8_t s_cmask[NC];
    __shared__ unsigned int s_n[8];  // 0 agents, 1 base of this tile's travel-list entries, 2 movers, 3 claimants that go
                                     // through the claim plane, 4 local claimants, 5 local claimants with rivals
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const bool noisy = p.noise_thr != 0ull;
    if (tid < 8) s_n[tid] = 0;
    for (int i = tid; i < NC / 16; i += NT) {
        reinterpret_cast<uint4*>(s_out)[i] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(s_cmask)[i] = make_uint4(0, 0, 0, 0);
    }
    // ---- the word of every cell in reach: per-step die roll (:364) recomputed, not communicated.  One group of 4
    // cells per thread and iteration: flag word + strategy word from global memory, one Philox call, 4 words out.
    constexpr int NIT = (NITEM + NT - 1) / NT;
    uint32_t tf4[NIT], st4[NIT];
    int sw[NIT];          // shared-memory word index of the group, -1 = no item
    uint32_t gq[NIT];     // global column of the group's first cell
    uint32_t gyv[NIT];    // global row
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int j = tid + it * NT;
        int ly, lx4;
        if (j < NI) { ly = j / (TW / 4); lx4 = 4 * (j - ly * (TW / 4)); }
        else {
            const int r = j - NI;
            if (r < GROW) { ly = -1; lx4 = 4 * r - 4; }
            else if (r < 2 * GROW) { ly = TH; lx4 = 4 * (r - GROW) - 4; }
            else { const int k = r - 2 * GROW; ly = k >> 1; lx4 = (k & 1) ? TW : -4; }
        }
        sw[it] = j < NITEM ? (ly + 1) * SWR + 4 + lx4 : -1;
        gq[it] = (x0 + (uint32_t)lx4) & (p.W - 1);
        gyv[it] = (p.row0 + y0 + (uint32_t)ly) & (p.HW - 1);
        tf4[it] = 0; st4[it] = 0;
        const int yl = (int)y0 + ly;  // local row: wraps for a whole world; a slab reads rows beyond its planes as empty
        const bool in = j < NITEM && (p.wrap_y || (yl >= 0 && yl < (int)p.H));
        if (in) {
            const size_t off = ((size_t)(p.wrap_y ? ((uint32_t)yl & (p.H - 1)) : (uint32_t)yl) << p.log2W) + gq[it];
            tf4[it] = *reinterpret_cast<const uint32_t*>(p.tfA + off);
            st4[it] = *reinterpret_cast<const uint32_t*>(p.strat + off);
        }
    }
    // the energies go straight to shared memory (cp.async), requested after the flag / strategy words (which the next
    // phase waits for); they are first needed by the agents
    for (int i = tid; i < NC / 4; i += NT) {
        const int ly = (i * 4) >> LOG_TW, lx = (i * 4) & (TW - 1);
        const float* src = p.E0 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(s_e + i * 4)), "l"(src));
    }
    asm volatile("cp.async.commit_group;");
    static_assert(NT == 256, "one table entry per thread");
    s_lut[tid] = __ldg(p.lut + tid);  // lut_entry() of every index, filled by pd_create
    __syncthreads();  // the list counters are zero (the loads above are in flight meanwhile)
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int j = tid + it * NT;
        if (it * NT >= NITEM) break;
        if (j < NITEM) {
            const uint32_t occ4 = tf4[it] & 0x80808080u;
            uint4 r4 = make_uint4(0, 0, 0, 0);
            uint32_t v[4] = {0, 0, 0, 0};
            if (occ4) {
                const uint64_t g4 = (((uint64_t)gyv[it] << p.log2W) | gq[it]) >> 2;
                const Words w = rng(p, g4, S_ROLL);
                const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                // Branch-free, straight from the two plane words (shift + mask-and-merge per field).  Only bit 31 says
                // "occupied": an EMPTY cell of the group keeps roll and strategy bits (its flag byte is 0, so no trait,
                // no D flag) -- such a word compares below every threshold and is never looked at otherwise.
                static_assert(ROLL_SHIFT == 16 && TF_OCC == 0x80u && TF_TRAIT == 0x03u, "word layout");
                const uint32_t T = tf4[it], S = st4[it];
                uint32_t rr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t r = ((k == 3 ? T : T << (24 - 8 * k)) & R_OCC) | ((ww[k] >> 1) & 0x7FFF8000u);
                    r |= (k == 0 ? S << 3 : S >> (8 * k - 3)) & 0x7F8u;
                    r |= (k == 0 ? T << 1 : T >> (8 * k - 1)) & 0x6u;
                    rr[k] = r;
                    v[k] = ((uint32_t)(j * 4 + k)) | (ww[k] << 16);  // (interior: j * 4 + k is the tile-local cell)
                }
                if (T & (TF_D * 0x01010101u)) {  // rare: an at-risk agent that dies during play (k_risk_resolve)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if ((T >> (8 * k)) & TF_D) rr[k] |= R_DFLAG;
                }
                r4 = make_uint4(rr[0], rr[1], rr[2], rr[3]);
            }
            *reinterpret_cast<uint4*>(s_R + sw[it]) = r4;
            if (j < NI) tappend4(&s_n[0], s_cells, occ4, v);
        }
    }
    asm volatile("cp.async.wait_all;");
    __syncthreads();
    // ---- agents: roles, games, terminal rule, travel decision -- one lane per agent, no warp-wide operations
    const unsigned int n_agent = s_n[0];
    unsigned int n_play_deaths = 0, n_movers = 0;
    for (unsigned int a = tid; a < n_agent; a += NT) {
        const uint32_t ent = s_cells[a];
        const int cl = (int)(ent & 0xFFFFu);
        const uint32_t coins = ent >> 16;
        const int ly = cl >> LOG_TW, lx = cl & (TW - 1);
        const int si = (ly + 1) * SWR + 4 + lx;
        const uint32_t rme = s_R[si];
        uint32_t rn[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) rn[d] = s_R[si + DY(d) * SWR + DX(d)];
        const uint32_t tA = rme | R_LOW, tB = (rme & ~R_LOW) - 1u;  // neighbour d challenges me iff rn[d] > (d < 4 ? tA : tB)
        const uint32_t any_nb = (rn[0] | rn[1] | rn[2]) | (rn[3] | rn[4] | rn[5]) | (rn[6] | rn[7]);
        // act[7] / act[0] of the terminal rule come from the roll order alone (get_game_list, :413): a
        // challenger that dies early still "would have challenged".
        // act[7]==1 <=> neighbour 7 exists and does not challenge me; act[0]==0 <=> neighbour 0 challenges me
        const bool act7_challenge = (rn[7] & R_OCC) && !(rn[7] > tB);
        const bool act0_respond = rn[0] > tA;
        if (any_nb & R_DFLAG) {  // rare: a neighbour dies during play (D code = sub-step of death + 1)
#pragma unroll
            for (int d = 0; d < 8; ++d)
                if (rn[d] & R_DFLAG) {
                    const uint32_t dn = ((uint32_t)p.tfA[nb_local(p, x0 + lx, y0 + ly, d)] & TF_D) >> TF_D_SHIFT;
                    if ((int)dn - 1 < 7 - d) rn[d] = 0u;  // dead before sub-step 7-d: it sends no challenge
                }
        }
        bool dead = false, any = false;
        const float e0 = s_e[cl];
        float e;
        if (noisy) e = play_games<true>(p, s_lut, rn, rme, tA, tB, coins, gcell_of(p, x0 + (uint32_t)lx, y0 + (uint32_t)ly), e0, dead, any);
        else e = play_games<false>(p, s_lut, rn, rme, tA, tB, coins, 0ull, e0, dead, any);
        if (dead) {
            s_out[cl] = (uint8_t)TF_GHOST;
            s_e[cl] = 0.0f;
            ++n_play_deaths;
        } else {
            // terminal rule (:547-556, :719-730, :486-496) or no neighbours (:429-433); games_played (:710) of an
            // agent that is still alive counts all its games as the responder
            const bool mover = !(any_nb & R_OCC) || (!any && (act7_challenge || act0_respond));
            s_out[cl] = (uint8_t)(TF_OCC | ((rme >> 1) & TF_TRAIT));
            s_e[cl] = e;
            // (+1 spelled as my word's OCC bit: for an increment it can prove uniform, ptxas wraps the atomic in a
            // vote / elect / shuffle sequence of ~15 issue slots)
            if (mover) { s_mover[satom_add(&s_n[2], rme >> 31)] = (uint16_t)cl; ++n_movers; }
        }
    }
    __syncthreads();
    // ---- movers: travel cost, death, random start, probe, claim (:794-868)
    const unsigned int n_mover = s_n[2];
    const uint32_t step = __ldg(&p.ctr->step_no);
    unsigned int n_travel_deaths = 0;
    for (unsigned int k = tid; k < n_mover; k += NT) {
        const int cl = s_mover[k];
        const int ly = cl >> LOG_TW, lx = cl & (TW - 1);
        const int si = (ly + 1) * SWR + 4 + lx;
        const float e = __fsub_rn(s_e[cl], p.travel_cost);  // :800
        if (e <= 0.0f) {                                    // :801-804
            s_out[cl] = (uint8_t)TF_GHOST;
            s_e[cl] = 0.0f;
            ++n_travel_deaths;
            continue;
        }
        s_e[cl] = e;
        uint32_t occ_mask = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) occ_mask |= (s_R[si + DY(d) * SWR + DX(d)] >> 31) << d;
        const Words w = rng(p, gcell_of(p, x0 + lx, y0 + ly), S_TRAVEL);
        int l = (int)(w.y >> 29), seq = 0;  // :810
        const int dir = probe(occ_mask, l, seq);
        if (dir >= 0) {  // (the agent list is spent: s_cells takes the claimants -- plane-bound from the front, local from the back)
            const int tx = lx + DX(dir), ty = ly + DY(dir);
            if (tx >= 1 && tx <= TW - 2 && ty >= 1 && ty <= TH - 2) {
                const int tc = ty * TW + tx;
                satom_or(reinterpret_cast<unsigned int*>(s_cmask) + (tc >> 2), 1u << (8 * (tc & 3) + (7 - dir)));
                s_cells[NC - 1 - satom_add(&s_n[4], 1u)] = (uint32_t)cl | ((uint32_t)dir << 16);
            } else {
                prefetch_claim(p.claim + nb_local(p, x0 + lx, y0 + ly, dir));
                s_cells[satom_add(&s_n[3], 1u)] = (uint32_t)cl | ((uint32_t)dir << 16);
            }
        }
    }
    __syncthreads();
    // ---- local claims (move_response, :881-960, for the targets only this tile can reach): a claimant that is alone
    // on its target's byte has won and relocates inside the tile before the tile is stored -- no claim word, no list
    // entry, no scattered writes but the strategy byte.  Rivals (they all see the same byte) go through the claim
    // plane like the claims on the tile's rim: k_move_commit decides by priority.
    const unsigned int n_local = s_n[4];
    unsigned int n_moved = 0;
    for (unsigned int k = tid; k < n_local; k += NT) {
        const uint32_t ent = s_cells[NC - 1 - k];
        const int cl = (int)(ent & 0xFFFFu), dir = (int)(ent >> 16);
        const int ly = cl >> LOG_TW, lx = cl & (TW - 1);
        const int tx = lx + DX(dir), ty = ly + DY(dir), tc = ty * TW + tx;
        if ((uint32_t)s_cmask[tc] == (1u << (7 - dir))) {  // :946-956
            s_out[tc] = s_out[cl];
            s_e[tc] = s_e[cl];
            s_out[cl] = (uint8_t)TF_GHOST;  // keeps blocking movers this step (quirk B-13)
            s_e[cl] = 0.0f;
            p.strat[((size_t)(y0 + ty) << p.log2W) + x0 + tx] = (uint8_t)(s_R[(ly + 1) * SWR + 4 + lx] >> 3);
            ++n_moved;
        } else {
            prefetch_claim(p.claim + nb_local(p, x0 + lx, y0 + ly, dir));
            s_mover[satom_add(&s_n[5], 1u)] = (uint16_t)((uint32_t)cl | ((uint32_t)dir << 11));  // (the mover list is spent)
        }
    }
    __syncthreads();
    // ---- coalesced stores
    const unsigned int n_plane = s_n[3], n_claim = n_plane + s_n[5];
    if (tid == 0 && n_claim) s_n[1] = atomicAdd(&p.ctr->mv_n, n_claim);
    for (int i = tid; i < NC / 16; i += NT) {
        const int ly = (i * 16) >> LOG_TW, lx = (i * 16) & (TW - 1);
        *reinterpret_cast<uint4*>(p.tfB + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<uint4*>(s_out)[i];
    }
    for (int i = tid; i < NC / 4; i += NT) {
        const int ly = (i * 4) >> LOG_TW, lx = (i * 4) & (TW - 1);
        *reinterpret_cast<float4*>(p.E1 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<float4*>(s_e)[i];
    }
    // ---- the claimants of this tile go on the claim list (k_move_commit): cell, claimed direction
    if (n_claim) {  // uniform across the CTA
        __syncthreads();
        const unsigned int base = s_n[1];
        if (base + n_claim > p.list_cap) {
            if (tid == 0) atomicExch(&p.ctr->overflow, 1u);
        } else {
            for (unsigned int k = tid; k < n_claim; k += NT) {
                uint32_t cl;
                int dir;
                if (k < n_plane) { const uint32_t ent = s_cells[k]; cl = ent & 0xFFFFu; dir = (int)(ent >> 16); }
                else { const uint32_t ent = s_mover[k - n_plane]; cl = ent & 0x7FFu; dir = (int)(ent >> 11); }
                const uint32_t x = x0 + (cl & (TW - 1)), y = y0 + (cl >> LOG_TW);
                // :861-868: the request goes to the target cell.  Posted here, after the CTA's last barrier (a barrier
                // waits for the CTA's outstanding reductions), with the priority drawn again rather than kept
                post_claim(p.claim + nb_local(p, x, y, dir), claim_word(step, CLAIM_TRAVEL, rng(p, gcell_of(p, x, y), S_TRAVEL).x, dir));
                p.cl_cell[base + k] = (y << p.log2W) | x;
                p.cl_e[base + k] = s_e[cl];  // (after the travel cost)
                // the mover's trait and strategies ride along (bits 1-2 and 3-10 of its word)
                p.cl_aux[base + k] = (uint16_t)((uint32_t)dir | ((s_R[((cl >> LOG_TW) + 1) * SWR + 4 + (cl & (TW - 1))] & 0x7FEu) << 2));
            }
        }
    }
    if (p.stats && row_owned(p, y0)) {  // tiles are entirely owned or entirely ghost (16-row granularity)
        if (n_play_deaths) atomicAdd(&p.ctr->st_play_deaths, (unsigned long long)n_play_deaths);
        if (n_movers) atomicAdd(&p.ctr->st_movers, (unsigned long long)n_movers);
        if (n_moved) atomicAdd(&p.ctr->st_moved, (unsigned long long)n_moved);
        if (n_travel_deaths) atomicAdd(&p.ctr->st_travel_deaths, (unsigned long long)n_travel_deaths);
    }
}

// ------------------------------------------------------------------------------------------
// k_claim_punish (dense sweep 2): neighbourhood_update (:975-1031) + first god_go_forth (:1042-1134), fused
// with environmental_punishment (:1339-1371), step_fn's counts (:1405-1419) and the next step's at-risk list.
// Reads tfB (1-halo; the travel rounds are over, GHOSTs are free cells again) and E1; writes tfA and E0, the
// canonical state of the next step.
//   * A parent that finds a free cell posts its claim on it and goes on the claim list (god_multiply is resolved
//     by k_repro_commit / k_repro_retry); the agents of the owned rows are counted (the carrying-capacity gate).
//   * The reference charges the cost of living in layer 7, after the god submodel (:2157-2163).  The only thing
//     the reproduction rounds change about it is the reproduction cost of the parents that win a claim, so
//     everybody is charged HERE as a non-parent (cell-local: min(e, max) - cost, death at <= 0) and a winner
//     re-derives its own outcome from E1, which stays read-only (parent_outcome below).  The rounds themselves
//     run on tfB, where an agent that has just died of the living cost still occupies its cell, as in the
//     reference; the newborns are exempt (they enter tfA / E0 in k_finalize, if they survive the cull).
// A slab leaves the state of its ghost rows alone: the next halo exchange overwrites them.
// ------------------------------------------------------------------------------------------
struct Outcome {
    float e;      // energy after the cost of living (0 if dead)
    bool alive, risk;
};
// environmental_punishment for one agent that entered layer 6 with energy e1 (:1357-1368); a parent paid the
// reproduction cost first (:1211-1213)
__device__ __forceinline__ Outcome outcome_of(const Params& p, float e1, bool parent) {
    float e = parent ? __fsub_rn(e1, p.reproduce_cost) : e1;
    Outcome o;
    o.alive = punish(p, e);
    o.e = o.alive ? e : 0.0f;
    o.risk = o.alive && e <= p.risk_thr;
    return o;
}

// A parent that has won its claim (god_multiply, :1211-1213) re-derives its layer-7 outcome: k_claim_punish
// charged it as a non-parent.  e1 = its energy when it entered layer 6 (E1 is read-only), `cell` its cell, owned
// rows only (nobody writes a parent's strat / tfB entries during the reproduction rounds).  Almost always only the energy changes; deaths, counts and the at-risk list are corrected if not.
__device__ __forceinline__ void parent_outcome(const Params& p, uint32_t cell, float e1, uint32_t trait) {
    const Outcome a = outcome_of(p, e1, false), b = outcome_of(p, e1, true);
    p.E0[cell] = b.e;
    if (a.alive != b.alive) {
        p.tfA[cell] = (uint8_t)(b.alive ? (TF_OCC | trait) : 0u);
        atomicAdd(&p.ctr->living_deaths, b.alive ? ~0ull : 1ull);  // (+1, or -1 if a negative cost revives it)
        if (p.want_counts) atomicAdd(&p.ctr->bins[count_bin(p.strat[cell], trait)], b.alive ? 1ull : ~0ull);
    }
    if (b.risk && !a.risk) {  // (an entry that is no longer at risk is harmless: k_risk_resolve replays it exactly)
        const unsigned int slot = atomicAdd(&p.ctr->risk_next_n, 1u);
        if (slot < p.list_cap) p.risk_next[slot] = cell;
        else atomicExch(&p.ctr->overflow, 1u);
    }
}

template <int TW, int TH>
__global__ void __launch_bounds__(NT, PD_CLAIM_CTAS) k_claim_punish(const __grid_constant__ Params p) {
    constexpr int SW = TW + 32, SH = TH + 2, NC = TW * TH;
    __shared__ __align__(16) uint8_t s_tf[SH * SW];
    __shared__ uint16_t s_par[NC];     // the agents that may reproduce
    __shared__ uint16_t s_claim[NC];   // claimants: tile-local cell | direction << 11 -- those that go through the claim
                                       // plane from the front, the local ones from the back
    __shared__ uint32_t s_prio[NC];
    __shared__ uint16_t s_risk[NC];
    // Claims on cells ALL of whose neighbours lie in this tile (everything but the tile's outermost cells) cannot
    // have rivals outside the tile: they are resolved right here, by an atomic max per cell in shared memory, and
    // never touch the claim plane.  The key is the claimant's priority with its position j in the low 3 bits --
    // unique among the rivals of a cell, and ordered like claim_word's (priority, j) unless two rivals agree in the
    // top 29 bits of their priorities.  That case is noticed when it could matter (the atomic returns a value with
    // my top bits: see below) and sends ALL local claimants of the tile through the claim plane instead, which
    // compares whole words.
    __shared__ __align__(16) unsigned int s_max[NC];
    __shared__ unsigned int s_bins[16];
    // 0 parents, 1 claimants for the claim plane, 2 base of their list entries, 3 agents, 4 at-risk, 5 their base,
    // 6 deaths, 7 local claimants, 8 base of their list entries, 12 set if two local rivals' priorities agree in
    // their top bits
    __shared__ unsigned int s_n[13];
    static_assert(NC <= 2048, "a tile-local cell index takes 11 bits");
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const bool owned = row_owned(p, y0);  // tiles are entirely owned or entirely ghost (16-row granularity)
    if (tid < 13) s_n[tid] = 0;
    if (tid < 16) s_bins[tid] = 0;
    for (int i = tid; i < NC / 4; i += NT) reinterpret_cast<uint4*>(s_max)[i] = make_uint4(0, 0, 0, 0);
    // The flag tile first -- the whole CTA waits for it at the barrier -- then the energies, all in flight together
    // (empty cells hold 0 in E1: loading them is only wasted bandwidth, and only in sparse worlds)
    constexpr int NCH = TW / 16 + 2;  // 16-byte chunks per tile row incl. the ring's
    static_assert(SH * NCH <= NT, "one chunk of the flag tile per thread");
    uint4 tile_v = make_uint4(0, 0, 0, 0);
    if (tid < SH * NCH) {
        const int r = tid / NCH, j = tid - r * NCH;
        const uint32_t cx = ((x0 >> 4) + (uint32_t)(j - 1)) & ((p.W >> 4) - 1);
        const int gy = (int)y0 + r - 1;  // a slab reads rows beyond its planes as empty (they are outside the valid zone)
        if (p.wrap_y || (gy >= 0 && gy < (int)p.H))
            tile_v = *reinterpret_cast<const uint4*>(p.tfB + ((size_t)(p.wrap_y ? (uint32_t)gy & (p.H - 1) : (uint32_t)gy) << p.log2W) +
                                                     ((size_t)cx << 4));
    }
    constexpr int NIT = (NC / 4 + NT - 1) / NT;
    float4 e_pre[NIT];
    uint32_t st_pre[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int i = tid + it * NT;
        e_pre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        st_pre[it] = 0;
        if (i < NC / 4) {
            const int ly = (i * 4) / TW, lx = (i * 4) - ly * TW;
            const size_t off = ((size_t)(y0 + ly) << p.log2W) + x0 + lx;
            e_pre[it] = *reinterpret_cast<const float4*>(p.E1 + off);
            if (p.want_counts && owned) st_pre[it] = *reinterpret_cast<const uint32_t*>(p.strat + off);
        }
    }
    if (tid < SH * NCH) *reinterpret_cast<uint4*>(s_tf + (tid / NCH) * SW + (tid % NCH) * 16) = tile_v;
    __syncthreads();
    unsigned int n_occ = 0, n_dead = 0;
    uint32_t par_bits = 0;  // bit 4 * it + k: cell k of my group of iteration it holds an agent that may reproduce
    static_assert(4 * NIT <= 32, "parent bits");
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int i = tid + it * NT;
        if (i >= NC / 4) break;  // (warp-uniform)
        const int ly = (i * 4) / TW, lx = (i * 4) - ly * TW;
        const int si = (ly + 1) * SW + lx + 16;
        const uint32_t t4 = *reinterpret_cast<const uint32_t*>(s_tf + si);
        const uint32_t occ4 = t4 & 0x80808080u;
        n_occ += __popc(occ4);
        const float4 e4 = e_pre[it];
        const float ein[4] = {e4.x, e4.y, e4.z, e4.w};
        // :978-979, :1054 -- the agents that may reproduce, compacted with four ballots per warp
        uint32_t par4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) par4 |= ein[k] >= p.reproduce_min_energy ? (0x80u << (8 * k)) : 0u;
        par4 = p.max_children > 0u ? (par4 & occ4) : 0u;
        par_bits |= ((par4 >> 7) & 1u) << (4 * it) | ((par4 >> 15) & 1u) << (4 * it + 1) | ((par4 >> 23) & 1u) << (4 * it + 2)
                    | ((par4 >> 31) & 1u) << (4 * it + 3);
        if (owned) {  // ---- cost of living, as non-parents (:1357-1368); branch-free per cell
            float ev[4];
            uint32_t dead4 = 0, risk4 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t bit = 0x80u << (8 * k);
                const float e = __fsub_rn(fminf(ein[k], p.max_energy), p.cost_of_living);
                dead4 |= e > 0.0f ? 0u : bit;
                risk4 |= e <= p.risk_thr ? bit : 0u;
                ev[k] = e;
            }
            static_assert(TF_OCC == 0x80u && TF_TRAIT == 0x03u, "flag layout");
            dead4 &= occ4;
            const uint32_t live4 = occ4 & ~dead4;
            risk4 &= live4;
#pragma unroll
            for (int k = 0; k < 4; ++k) ev[k] = (live4 >> (8 * k + 7)) & 1u ? ev[k] : 0.0f;
            n_dead += __popc(dead4);
            if (p.want_counts)
                for (uint32_t m = live4; m; m &= m - 1) {
                    const int b = (__ffs(m) - 1) >> 3;
                    atomicAdd(&s_bins[count_bin((st_pre[it] >> (8 * b)) & 0xFFu, (t4 >> (8 * b)) & TF_TRAIT)], 1u);
                }
            if (risk4)  // rare
                for (uint32_t m = risk4; m; m &= m - 1) s_risk[satom_add(&s_n[4], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
            const size_t off = ((size_t)(y0 + ly) << p.log2W) + x0 + lx;
            // OCC | trait for the survivors, 0 for the dead, the empty and the vacated
            *reinterpret_cast<uint32_t*>(p.tfA + off) = t4 & (live4 | (live4 >> 6) | (live4 >> 7));
            *reinterpret_cast<float4*>(p.E0 + off) = make_float4(ev[0], ev[1], ev[2], ev[3]);
        }
    }
    if (par_bits) {  // one shared atomic per thread with its count, then plain stores (no ballots)
        unsigned int slot = satom_add(&s_n[0], (unsigned int)__popc(par_bits));
        for (uint32_t m = par_bits; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            s_par[slot++] = (uint16_t)((tid + (b >> 2) * NT) * 4 + (b & 3));
        }
    }
    n_occ = __reduce_add_sync(FULL, n_occ);
    n_dead = __reduce_add_sync(FULL, n_dead);
    if ((tid & 31) == 0) {
        if (n_occ) satom_add(&s_n[3], n_occ);
        if (n_dead) satom_add(&s_n[6], n_dead);
    }
    __syncthreads();
    const unsigned int n_par = s_n[0];
    const uint32_t step = __ldg(&p.ctr->step_no);
    for (unsigned int k = tid; k < n_par; k += NT) {
        const int cl = s_par[k];
        const int ly = cl / TW, lx = cl - ly * TW;
        const int si = (ly + 1) * SW + lx + 16;
        uint32_t occ_mask = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d)
            if (s_tf[si + DY(d) * SW + DX(d)] & TF_OCC) occ_mask |= 1u << d;
        if (occ_mask == 0xFFu) continue;  // REPRODUCTION_IMPOSSIBLE (:1028)
        const Words w = rng(p, gcell_of(p, x0 + lx, y0 + ly), S_REPRO);
        int l = (int)(w.y >> 29), seq = 0;  // :1079
        const int dir = probe(occ_mask, l, seq);
        if (dir >= 0) {
            const int tx = lx + dxr(dir), ty = ly + dyr(dir);
            unsigned int slot;
            if (tx >= 1 && tx <= TW - 2 && ty >= 1 && ty <= TH - 2) {
                // Whoever ends up with the maximum of a cell, a rival with the same top bits either posts after it (and
                // sees it) or before it (then the later one sees that rival or something in between, which has the same
                // top bits too).  An empty entry (0) counts as a value: a false alarm only costs the detour.
                const uint32_t key = (w.x & ~7u) | (uint32_t)(7 - dir);
                const uint32_t old = atomicMax(&s_max[ty * TW + tx], key);
                if (((old ^ key) >> p.tie_shift) == 0u) s_n[12] = 1u;
                slot = NC - 1 - satom_add(&s_n[7], 1u);
            } else {
                prefetch_claim(p.claim + nb_local(p, x0 + lx, y0 + ly, dir));
                slot = satom_add(&s_n[1], 1u);
            }
            s_claim[slot] = (uint16_t)((uint32_t)cl | ((uint32_t)dir << 11));
            s_prio[slot] = w.x;
        }
    }
    __syncthreads();
    // (on the detour the local claimants join the others: entries n_plane.. of the list below)
    const bool detour = s_n[12] != 0u;
    const unsigned int n_plane = s_n[1], n_local = detour ? 0u : s_n[7], n_claim = n_plane + (detour ? s_n[7] : 0u);
    const unsigned int n_risk = s_n[4];
    // (one warp each: a thread's second atomic would wait for the first one's answer to be stored)
    if (tid == 0 && n_claim) s_n[2] = atomicAdd(&p.ctr->rp_n, n_claim);
    if (tid == 32 && n_local) s_n[8] = atomicAdd(&p.ctr->lc_n, n_local);
    if (tid == 64 && n_risk) s_n[5] = atomicAdd(&p.ctr->risk_next_n, n_risk);
    if (tid == 96) {
        if (s_n[3] && owned) atomicAdd(&p.ctr->n_old, (unsigned long long)s_n[3]);
        if (s_n[6]) atomicAdd(&p.ctr->living_deaths, (unsigned long long)s_n[6]);
    }
    if (p.want_counts && tid < 16 && s_bins[tid]) atomicAdd(&p.ctr->bins[tid], (unsigned long long)s_bins[tid]);
    if (n_claim == 0 && n_local == 0 && n_risk == 0) return;  // uniform across the CTA
    __syncthreads();
    // ---- local claims (god_multiply, :1144-1335, for the targets only this tile can reach): the winner pays right
    // away (parent_outcome: its E0 / tfA entries were written by this CTA, two barriers ago).  Winners and losers are
    // listed at the BACK of the claim list (entry list_cap - 1 - i) with the verdict in the entry; k_repro_commit
    // turns them into newborns / retry entries without a random access.
    if (n_local) {
        const unsigned int base = s_n[8];
        if (base + n_local > p.list_cap) {
            if (tid == 0) atomicExch(&p.ctr->overflow, 1u);
        } else {
            for (unsigned int k = tid; k < n_local; k += NT) {
                const unsigned int slot = NC - 1 - k;
                const uint32_t ent = s_claim[slot], cl = ent & 0x7FFu;
                const int dir = (int)(ent >> 11);
                const uint32_t lx = cl % TW, ly = cl / TW;
                const uint32_t cell = ((y0 + ly) << p.log2W) | (x0 + lx);
                const bool won = s_max[(int)cl + dyr(dir) * TW + dxr(dir)] == ((s_prio[slot] & ~7u) | (uint32_t)(7 - dir));
                if (won) {  // REPRODUCTION_COMPLETE (:1211-1213, :1331)
                    const uint32_t trait = (uint32_t)s_tf[(ly + 1) * SW + lx + 16] & TF_TRAIT;
                    if (owned) parent_outcome(p, cell, p.E1[cell], trait);
                    else  // a child born into a ghost row belongs to the neighbour slab: here it only blocks its cell
                        p.tfB[(long long)cell + dyr(dir) * (long long)p.W + dxr(dir)] = (uint8_t)(TF_OCC | TF_NEWBORN | trait);
                }
                p.cl_cell[p.list_cap - 1u - (base + k)] = cell;
                p.cl_aux[p.list_cap - 1u - (base + k)] = (uint16_t)((uint32_t)dir | (won ? 8u : 0u));
            }
        }
    }
    if (n_claim) {
        const unsigned int base = s_n[2];
        if (base + n_claim > p.list_cap) {
            if (tid == 0) atomicExch(&p.ctr->overflow, 1u);
        } else {
            for (unsigned int k = tid; k < n_claim; k += NT) {
                const unsigned int slot = k < n_plane ? k : NC - 1 - (k - n_plane);
                const uint32_t ent = s_claim[slot], cl = ent & 0x7FFu;
                const uint32_t x = x0 + cl % TW, y = y0 + cl / TW;
                const int dir = (int)(ent >> 11);
                // :1123-1130: the request goes to the target cell (posted after the CTA's last barrier: a barrier
                // waits for the CTA's outstanding reductions)
                post_claim(p.claim + nb_local(p, x, y, dir), claim_word(step, CLAIM_REPRO, s_prio[slot], dir));
                const uint32_t cell = (y << p.log2W) | x;
                p.cl_cell[base + k] = cell;
                p.cl_e[base + k] = p.E1[cell];  // (this CTA has just read the tile: an L2 hit)
                p.cl_aux[base + k] = (uint16_t)((uint32_t)dir | (((uint32_t)s_tf[(cl / TW + 1) * SW + cl % TW + 16] & TF_TRAIT) << 3));
            }
        }
    }
    if (n_risk) {
        const unsigned int base = s_n[5];
        if (base + n_risk > p.list_cap) {
            if (tid == 0) atomicExch(&p.ctr->overflow, 1u);
        } else {
            for (unsigned int k = tid; k < n_risk; k += NT) {
                const uint32_t cl = s_risk[k];
                p.risk_next[base + k] = ((y0 + cl / TW) << p.log2W) | (x0 + cl % TW);
            }
        }
    }
}

}  // namespace pd

// The four dense sweeps of a step (compacted work lists, shared-memory claim scatter).
//
//   k_play_travel   tfA, strat (1-halo), E0 -> tfB, E1     search/game list/8 play sub-steps/first move_request
//   k_move_commit   tfB (2-halo)            -> tfA          move_response round 1 (winners are pulled)
//   k_repro_claim   tfA (1-halo), E1        -> tfB          neighbourhood_update + first god_go_forth
//   k_repro_commit  tfB (2-halo), E1, strat -> tfA, E0      god_multiply round 1 + environmental_punishment + counts
//
// Design rules (the v1 profile showed all four issue-bound at 8 of 32 lanes active, DRAM traffic
// already equal to the algorithmic bytes -- profiles/README.md):
//   * u8 planes enter shared memory as 16-byte chunks (interior at column 16 of a TW+32 wide row);
//     the common path touches 4 cells per thread as one 32-bit word / one float4;
//   * anything that needs a Philox call (rolls, RANDOM/noisy games, movers, parents, contests,
//     births) is first COMPACTED into a shared-memory work list and then processed with all
//     32 lanes busy;
//   * claims are resolved by scattering one bit per claimant into a shared-memory byte per target
//     (bit j = "my neighbour j claims me"); only cells with >= 2 bits need priorities;
//   * global list appends are aggregated per CTA (one atomicAdd per list per CTA);
//   * measured since (profiles/README.md): ballots / shuffles in a hot loop, nvcc's warp-aggregated
//     atomicAdd(&shared_counter, 1) in divergent code (use satom_add) and extra __syncthreads() phases with
//     little work in them all cost more than their instruction count suggests.
// Every sweep is still a pure gather in global memory: a cell is written by its own tile only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pd_device.cuh"

namespace pd {

constexpr int NT = 256;  // threads per CTA of the dense sweeps
constexpr uint32_t FULL = 0xFFFFFFFFu;

// contest key of a claimant: (priority word, global cell) -- higher wins (:929, :1192)
struct Key {
    uint32_t prio;
    uint64_t gcell;
};
__device__ __forceinline__ bool key_gt(const Key& a, const Key& b) {
    return a.prio > b.prio || (a.prio == b.prio && a.gcell > b.gcell);
}
struct TravelKey {
    __device__ __forceinline__ Key operator()(const Params& p, uint64_t g) const {
        Key k; k.prio = rng(p, g, S_ROLL).y; k.gcell = g; return k;  // :806
    }
};
struct ReproKey {
    __device__ __forceinline__ Key operator()(const Params& p, uint64_t g) const {
        Key k; k.prio = rng(p, g, S_REPRO).x; k.gcell = g; return k;  // :1076
    }
};
__device__ __forceinline__ bool claims_dir(uint32_t t, int dir) {
    return (t & TF_CLAIM) && (int)((t & TF_DIR) >> TF_DIR_SHIFT) == dir;
}
// living cost for one agent (environmental_punishment, :1357-1368); returns false if it dies
__device__ __forceinline__ bool punish(const Params& p, float& e) {
    e = __fsub_rn(fminf(e, p.max_energy), p.cost_of_living);
    return e > 0.0f;
}

// ---- shared-memory tile of a u8 plane: (TH + 2*HALO) rows of TW+32 bytes, interior at column 16
template <int TW, int TH, int HALO>
__device__ __forceinline__ void load_tile_vec(const Params& p, const uint8_t* __restrict__ plane,
                                              uint8_t* s, uint32_t x0, uint32_t y0) {
    constexpr int NCH = TW / 16 + 2;
    constexpr int SW = TW + 32;
    constexpr int SH = TH + 2 * HALO;
    const uint32_t wch = p.W >> 4;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < SH * NCH; i += NT) {
        const int r = i / NCH, j = i - r * NCH;
        const uint32_t cx = ((x0 >> 4) + (uint32_t)(j - 1)) & (wch - 1);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (p.wrap_y) {
            const uint32_t gy = (y0 + (uint32_t)(r - HALO)) & (p.H - 1);
            v = *reinterpret_cast<const uint4*>(plane + ((size_t)gy << p.log2W) + ((size_t)cx << 4));
        } else {  // slab: rows beyond the local planes read as empty (they are outside the valid zone)
            const int gy = (int)y0 + r - HALO;
            if (gy >= 0 && gy < (int)p.H)
                v = *reinterpret_cast<const uint4*>(plane + ((size_t)gy << p.log2W) + ((size_t)cx << 4));
        }
        *reinterpret_cast<uint4*>(s + r * SW + j * 16) = v;
    }
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
//     empty cell: 0
//     occupied:   bit 31 | roll (19 bits) << 12 | has-D-code << 11 | strategies (8) << 3 | trait (2) << 1
// so one shared-memory load per neighbour answers "occupied?", "does it challenge me?" (an unsigned
// compare of the whole word against one of two thresholds: a tie of the rolls goes to the neighbour in a
// direction >= 4) and yields both strategies of the game.  An agent then plays its (up to) 8 games as the
// responder in sub-step order straight from those words: the strategies, the two RANDOM coins (bits of the
// agent's own roll call, kept next to its list entry) and -- with noise -- the two flips form the index
// of a 256-entry payoff table in shared memory, so a game costs one table look-up and no random call.
// ------------------------------------------------------------------------------------------

constexpr uint32_t R_OCC = 0x80000000u;
constexpr uint32_t R_DFLAG = 0x800u;
constexpr uint32_t R_LOW = 0xFFFu;  // everything below the roll

// append the byte indices (b0 + k) of the bytes of `occ4` that have bit 7 set; all 32 lanes call
__device__ __forceinline__ void wappend4(unsigned int* cnt, uint32_t* list, uint32_t occ4, int b0) {
    const int lane = threadIdx.x & 31;
    const int n = __popc(occ4);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    unsigned int base = 0;
    if (lane == 31) base = satom_add(cnt, (unsigned int)total);
    base = __shfl_sync(FULL, base, 31);
    unsigned int k = base + (unsigned int)(incl - n);
    while (occ4) {
        const int b = (__ffs(occ4) - 1) >> 3;
        occ4 &= occ4 - 1;
        list[k++] = (uint32_t)(b0 + b);
    }
}

// bit 7 of every non-zero byte of v
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v) {
    return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

// bit 7 of every byte of v that has two or more bits set
__device__ __forceinline__ uint32_t multi_bit_bytes(uint32_t v) {
    uint32_t c = v - ((v >> 1) & 0x55555555u);
    c = (c & 0x33333333u) + ((c >> 2) & 0x33333333u);
    c = (c + (c >> 4)) & 0x0F0F0F0Fu;        // bits set per byte, 0..8
    return (c + 0x7E7E7E7Eu) & 0x80808080u;  // >= 2
}

// Reserve slots in up to 32/BITS shared-memory lists with ONE warp scan: field k (BITS wide) of `n` is the
// number of entries this lane appends to list k, cnt[k] is that list's counter; the warp's sum per field
// must fit in BITS bits.  The start offsets come back in off[k].  All 32 lanes must call.
// Measured: faster than per-cell shared atomics in k_move_commit; no gain in k_repro_claim /
// k_repro_commit, which keep the atomics (profiles/README.md).
template <int NL, int BITS>
__device__ __forceinline__ void wreserve(unsigned int* cnt, uint32_t n, uint32_t (&off)[NL]) {
    static_assert(NL * BITS <= 32, "packed counters");
    constexpr uint32_t FM = BITS == 32 ? 0xFFFFFFFFu : ((1u << BITS) - 1u);
    const int lane = threadIdx.x & 31;
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
#pragma unroll
    for (int k = 0; k < NL; ++k) off[k] = 0;
    if (total == 0u) return;
    const uint32_t excl = incl - n;
    uint32_t base = 0;
    if (lane < NL) {
        const uint32_t t = (total >> (BITS * lane)) & FM;
        if (t) base = satom_add(&cnt[lane], t);
    }
#pragma unroll
    for (int k = 0; k < NL; ++k) off[k] = __shfl_sync(FULL, base, k) + ((excl >> (BITS * k)) & FM);
}

// entry i of the payoff table: bits 0-1 the challenger's strategy against my trait, 2-3 my strategy against
// its trait (:599-600), 4 the challenger's RANDOM coin, 5 mine (:631, :664), 6 the challenger's noise flip,
// 7 mine (:638-640, :672-674) -> my payoff (:677-689)
__device__ __forceinline__ float lut_entry(const Params& p, uint32_t i) {
    const uint32_t theirs = i & 3u, mine = (i >> 2) & 3u;
    bool ch_coop = theirs == 3u ? ((i >> 4) & 1u) != 0u : theirs != 1u;
    bool i_coop = mine == 3u ? ((i >> 5) & 1u) != 0u : mine != 1u;
    if ((i >> 6) & 1u) ch_coop = !ch_coop;
    if ((i >> 7) & 1u) i_coop = !i_coop;
    return p.pay[((uint32_t)i_coop << 1) | (uint32_t)ch_coop];
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
__global__ void __launch_bounds__(NT, 5) k_play_travel(const __grid_constant__ Params p) {
    constexpr int SW = TW + 32, SH = TH + 2, NC = TW * TH, NRING = 2 * (TW + 2) + 2 * TH;
    constexpr int LOG_TW = TW == 128 ? 7 : 4;
    static_assert(TW == 128 || TW == 16, "tile width");
    static_assert(SH * SW < 65536, "a shared-memory cell index fits 16 bits");
    __shared__ __align__(16) uint8_t s_tf[SH * SW];
    __shared__ __align__(16) uint8_t s_st[SH * SW];
    __shared__ __align__(16) uint32_t s_R[SH * SW];
    __shared__ __align__(16) float s_e[NC];
    __shared__ __align__(16) uint8_t s_out[NC];
    // the occupied cells: agents of the tile first (from 0), then ring cells.  Entry = shared-memory cell
    // index | the agent's 16 coin bits << 16 (written by whoever computes its roll)
    __shared__ uint32_t s_cells[NC + NRING];
    __shared__ uint16_t s_mover[NC];
    __shared__ float s_lut[256];
    __shared__ unsigned int s_n[4];  // 0 agents, 1 ring cells, 2 movers
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const bool noisy = p.noise_thr != 0ull;
    if (tid < 4) s_n[tid] = 0;
    static_assert(NT == 256, "one table entry per thread");
    s_lut[tid] = lut_entry(p, (uint32_t)tid);
    load_tile_vec<TW, TH, 1>(p, p.tfA, s_tf, x0, y0);
    load_tile_vec<TW, TH, 1>(p, p.strat, s_st, x0, y0);
    for (int i = tid; i < NC / 4; i += NT) {
        const int ly = (i * 4) >> LOG_TW, lx = (i * 4) & (TW - 1);
        reinterpret_cast<float4*>(s_e)[i] =
            *reinterpret_cast<const float4*>(p.E0 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx);
    }
    for (int i = tid; i < NC / 16; i += NT) reinterpret_cast<uint4*>(s_out)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < SH * SW / 4; i += NT) reinterpret_cast<uint4*>(s_R)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // ---- compaction, 4 cells per lane: the agents of this tile ...
    for (int base = 0; base < NC / 4; base += NT) {
        const int i = base + tid;  // NC/4 is a multiple of 32, so whole warps are in or out
        if (i < NC / 4) {
            const int ly = (i * 4) >> LOG_TW, lx = (i * 4) & (TW - 1);
            const int si = (ly + 1) * SW + lx + 16;
            wappend4(&s_n[0], s_cells, *reinterpret_cast<const uint32_t*>(s_tf + si) & 0x80808080u, si);
        }
    }
    __syncthreads();
    // ... and the occupied cells of the 1-ring around it (they only need their word)
    const unsigned int n_agent = s_n[0];
    for (int base = 0; base < NRING; base += NT) {
        const int i = base + tid;
        int si = 0;
        bool occ = false;
        if (i < NRING) {
            if (i < TW + 2) si = 15 + i;                                           // top row
            else if (i < 2 * (TW + 2)) si = (SH - 1) * SW + 15 + (i - (TW + 2));   // bottom row
            else {
                const int k = i - 2 * (TW + 2);                                    // left / right columns
                si = (1 + (k >> 1)) * SW + ((k & 1) ? (16 + TW) : 15);
            }
            occ = (s_tf[si] & TF_OCC) != 0;
        }
        const uint32_t slot = wappend(&s_n[1], occ);
        if (occ) s_cells[n_agent + slot] = (uint32_t)si;
    }
    __syncthreads();
    // ---- per-step die roll (:364) of every occupied cell in reach -- recomputed, not communicated
    const unsigned int n_roll = n_agent + s_n[1];
    for (unsigned int k = tid; k < n_roll; k += NT) {
        const int si = (int)s_cells[k];
        const int r = si / SW, c = si - r * SW;
        const uint32_t gx = (x0 + (uint32_t)(c - 16)) & (p.W - 1);
        const uint32_t gy = (p.row0 + y0 + (uint32_t)(r - 1)) & (p.HW - 1);
        const Words w = rng(p, ((uint64_t)gy << p.log2W) | gx, S_ROLL);
        const uint32_t t = s_tf[si];
        s_R[si] = R_OCC | ((w.x >> ROLL_SHIFT) << 12) | ((t & TF_D) ? R_DFLAG : 0u) | ((uint32_t)s_st[si] << 3) |
                  ((t & TF_TRAIT) << 1);
        s_cells[k] = (uint32_t)si | (w.w << 16);
    }
    __syncthreads();
    // ---- agents: roles, games, terminal rule, travel decision -- one lane per agent, no warp-wide operations
    unsigned int n_play_deaths = 0, n_movers = 0;
    for (unsigned int a = tid; a < n_agent; a += NT) {
        const uint32_t ent = s_cells[a];
        const int si = (int)(ent & 0xFFFFu);
        const uint32_t coins = ent >> 16;
        const int ly = si / SW - 1, lx = si - (ly + 1) * SW - 16;
        const int cl = (ly << LOG_TW) | lx;
        const uint32_t rme = s_R[si];
        uint32_t rn[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) rn[d] = s_R[si + DY(d) * SW + DX(d)];
        const uint32_t tA = rme | R_LOW, tB = (rme & ~R_LOW) - 1u;  // neighbour d challenges me iff rn[d] > (d < 4 ? tA : tB)
        const uint32_t any_nb = (rn[0] | rn[1] | rn[2]) | (rn[3] | rn[4] | rn[5]) | (rn[6] | rn[7]);
        // act[7] / act[0] of the terminal rule come from the roll order alone (get_game_list, :413): a
        // challenger that dies early still "would have challenged".
        // act[7]==1 <=> neighbour 7 exists and does not challenge me; act[0]==0 <=> neighbour 0 challenges me
        const bool act7_challenge = rn[7] != 0u && !(rn[7] > tB);
        const bool act0_respond = rn[0] > tA;
        if (any_nb & R_DFLAG) {  // rare: a neighbour dies during play (D code = sub-step of death + 1)
#pragma unroll
            for (int d = 0; d < 8; ++d)
                if (rn[d] & R_DFLAG) {
                    const uint32_t dn = ((uint32_t)s_tf[si + DY(d) * SW + DX(d)] & TF_D) >> TF_D_SHIFT;
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
            const bool mover = any_nb == 0u || (!any && (act7_challenge || act0_respond));
            s_out[cl] = (uint8_t)(TF_OCC | ((rme >> 1) & TF_TRAIT));
            s_e[cl] = e;
            if (mover) { s_mover[satom_add(&s_n[2], 1u)] = (uint16_t)cl; ++n_movers; }
        }
    }
    __syncthreads();
    // ---- movers: travel cost, death, random start, probe, claim (:794-868)
    const unsigned int n_mover = s_n[2];
    unsigned int n_travel_deaths = 0;
    for (unsigned int k = tid; k < n_mover; k += NT) {
        const int cl = s_mover[k];
        const int ly = cl >> LOG_TW, lx = cl & (TW - 1);
        const int si = (ly + 1) * SW + lx + 16;
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
        for (int d = 0; d < 8; ++d) occ_mask |= (s_R[si + DY(d) * SW + DX(d)] >> 31) << d;
        const Words w = rng(p, gcell_of(p, x0 + lx, y0 + ly), S_ROLL);
        int l = (int)(w.z >> 29), seq = 0;  // :810
        const int dir = probe(occ_mask, l, seq);
        if (dir >= 0) s_out[cl] = (uint8_t)(s_out[cl] | TF_CLAIM | ((uint32_t)dir << TF_DIR_SHIFT));
    }
    __syncthreads();
    // ---- coalesced stores
    for (int i = tid; i < NC / 16; i += NT) {
        const int ly = (i * 16) >> LOG_TW, lx = (i * 16) & (TW - 1);
        *reinterpret_cast<uint4*>(p.tfB + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<uint4*>(s_out)[i];
    }
    for (int i = tid; i < NC / 4; i += NT) {
        const int ly = (i * 4) >> LOG_TW, lx = (i * 4) & (TW - 1);
        *reinterpret_cast<float4*>(p.E1 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<float4*>(s_e)[i];
    }
    if (p.stats && row_owned(p, y0)) {  // tiles are entirely owned or entirely ghost (16-row granularity)
        if (n_play_deaths) atomicAdd(&p.ctr->st_play_deaths, (unsigned long long)n_play_deaths);
        if (n_movers) atomicAdd(&p.ctr->st_movers, (unsigned long long)n_movers);
        if (n_travel_deaths) atomicAdd(&p.ctr->st_travel_deaths, (unsigned long long)n_travel_deaths);
    }
}

// ------------------------------------------------------------------------------------------
// Claim resolution shared by k_move_commit and k_repro_commit.
// s_tf: flags with a 2-cell halo; s_cl (same geometry, zeroed by the caller) receives one bit per
// claimant: bit j of a cell's byte = "my neighbour j claims me".  A cell with one bit has its
// winner; only cells with >= 2 bits need the claimants' Philox priorities, which winner_of()
// computes on demand (and caches by collapsing the byte to the winner's bit -- every thread that
// races on it writes the same value).  A claimant with direction d won iff the winner is 7-d.
// ------------------------------------------------------------------------------------------
template <int TW, int TH, int CELLS>
__device__ __forceinline__ void scatter_claims(const uint8_t* s_tf, uint8_t* s_cl, const int* s_dxy) {
    // the tile is scanned CELLS (4 or 16) cells per thread at a time; claimants are sparse (<= ~10 % of cells).
    // Measured: 16 is faster in k_move_commit (4 % claimants), 4 in k_repro_commit (5 %, longer work per claim).
    constexpr int SW = TW + 32, SH = TH + 4, NW = CELLS / 4, NCHUNK = SH * SW / CELLS;
    static_assert(SW % CELLS == 0 && (CELLS == 4 || CELLS == 16), "a chunk stays inside one row");
    for (int w = threadIdx.x; w < NCHUNK; w += NT) {
        uint32_t tw[NW];
        if (CELLS == 16) {
            const uint4 t16 = reinterpret_cast<const uint4*>(s_tf)[w];
            tw[0] = t16.x; tw[NW > 1 ? 1 : 0] = t16.y; tw[NW > 2 ? 2 : 0] = t16.z; tw[NW > 3 ? 3 : 0] = t16.w;
            if (!((t16.x | t16.y | t16.z | t16.w) & 0x40404040u)) continue;
        } else {
            tw[0] = reinterpret_cast<const uint32_t*>(s_tf)[w];
        }
        const int r = (w * CELLS) / SW, c0 = w * CELLS - r * SW;
#pragma unroll
        for (int h = 0; h < NW; ++h) {
            uint32_t cm = tw[h] & 0x40404040u;
            while (cm) {
                const int b = (__ffs(cm) - 1) >> 3;
                cm &= cm - 1;
                const int d = (int)(((tw[h] >> (8 * b)) & TF_DIR) >> TF_DIR_SHIFT);
                const int c = c0 + h * 4 + b;
                const int rr = r + s_dxy[8 + d], cc = c + s_dxy[d];  // (shared-memory LUT: LDC would go through the ADU)
                // only targets inside tile + 1-ring matter (rows 1..TH+2, columns 15..TW+16)
                if (rr < 1 || rr > TH + 2 || cc < 15 || cc > TW + 16) continue;
                const int st = rr * SW + cc;
                atomicOr(reinterpret_cast<unsigned int*>(s_cl) + (st >> 2), (1u << (7 - d)) << (8 * (st & 3)));
            }
        }
    }
}

// winning neighbour index of the claimed cell at shared index st (highest (priority, cell), :929, :1192)
template <int TW, typename KeyFn>
__device__ __forceinline__ int winner_of(const Params& p, uint8_t* s_cl, int st, uint32_t x0, uint32_t y0, KeyFn keyfn) {
    constexpr int SW = TW + 32;
    uint32_t m = s_cl[st];
    if (!(m & (m - 1u))) return __ffs(m) - 1;  // a single claimant
    const int r = st / SW, c = st - r * SW;
    const uint32_t tx = (x0 + (uint32_t)(c - 16)) & (p.W - 1);
    // local row of the claimed cell (may be -1 or H in the ring): only used through nb_gcell, which maps
    // it to the GLOBAL row modulo the world height
    const uint32_t ty = p.wrap_y ? ((y0 + (uint32_t)(r - 2)) & (p.H - 1)) : (y0 + (uint32_t)(r - 2));
    int best = -1;
    Key bk;
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const Key k2 = keyfn(p, nb_gcell(p, tx, ty, j));
        if (best < 0 || key_gt(k2, bk)) { bk = k2; best = j; }
    }
    s_cl[st] = (uint8_t)(1u << best);
    return best;
}

// Contested cells (two or more claim bits) of tile + 1-ring are listed and resolved in a phase of their own,
// one lane per cell: the Philox rounds of a contest then run once per cell on a few full-ish warps, instead of
// once per claimant AND per claimed cell inside the divergent work loops (17 % of k_repro_commit's instructions).
// append_contested: the cells of one 4-byte word of s_cl starting at shared index si0.
__device__ __forceinline__ void append_contested(uint32_t c4, int si0, uint16_t* s_ct, unsigned int* s_nct) {
    if (!(c4 & (c4 - 1u))) return;  // at most one bit in the whole word
    for (uint32_t m = multi_bit_bytes(c4); m; m &= m - 1) s_ct[satom_add(s_nct, 1u)] = (uint16_t)(si0 + ((__ffs(m) - 1) >> 3));
}
// the ring around the tile (rows 1 and TH+2, columns 15 and TW+16 of the 2-halo layout)
template <int TW, int TH>
__device__ __forceinline__ void append_contested_ring(const uint8_t* s_cl, uint16_t* s_ct, unsigned int* s_nct) {
    constexpr int SW = TW + 32, NRING = 2 * (TW + 2) + 2 * TH;
    for (int i = threadIdx.x; i < NRING; i += NT) {
        int st;
        if (i < TW + 2) st = 1 * SW + 15 + i;
        else if (i < 2 * (TW + 2)) st = (TH + 2) * SW + 15 + (i - (TW + 2));
        else {
            const int k = i - 2 * (TW + 2);
            st = (2 + (k >> 1)) * SW + ((k & 1) ? (16 + TW) : 15);
        }
        const uint32_t m = s_cl[st];
        if (m & (m - 1u)) s_ct[satom_add(s_nct, 1u)] = (uint16_t)st;
    }
}

// copy a CTA-local list of local cell indices into its reserved range of a global list
template <int TW>
__device__ __forceinline__ void copy_cells(const Params& p, const uint16_t* s_list, unsigned int n,
                                           unsigned int base, uint32_t* gcell_list, const uint8_t* s_state,
                                           uint8_t* gstate, uint32_t x0, uint32_t y0) {
    if (n == 0) return;
    if (base + n > p.list_cap) {
        if (threadIdx.x == 0) atomicExch(&p.ctr->overflow, 1u);
        return;
    }
    for (unsigned int k = threadIdx.x; k < n; k += NT) {
        const int cl = s_list[k];
        const int ly = cl / TW, lx = cl - ly * TW;
        gcell_list[base + k] = ((y0 + ly) << p.log2W) | (x0 + lx);
        if (gstate) gstate[base + k] = s_state[k];
    }
}

// ------------------------------------------------------------------------------------------
// k_move_commit (dense sweep 2): move_response (:881-960), round 1.  tfB (2-halo) -> tfA.
// An empty cell that received claims PULLS the winner's energy/strategies; a claimant looks up
// its target's winner and either vacates (leaving a GHOST that keeps blocking movers this
// step, quirk B-13) or joins the loser list.  Claimants and claimed cells are rare, so the dense
// scan only classifies words and queues them; the queue is then processed with full lanes.
// Loser state byte: 0x80 active | 0x40 fresh | claimed direction (k_move_retry derives the rest).
// ------------------------------------------------------------------------------------------
template <int TW, int TH>
__global__ void __launch_bounds__(NT, 8) k_move_commit(const __grid_constant__ Params p) {
    constexpr int SW = TW + 32, SH = TH + 4, NC = TW * TH;
    __shared__ __align__(16) uint8_t s_tf[SH * SW];
    __shared__ __align__(16) uint8_t s_cl[SH * SW];
    __shared__ __align__(16) uint8_t s_out[NC];
    __shared__ uint16_t s_work[NC];
    __shared__ uint16_t s_lost[NC];
    __shared__ uint8_t s_lost_state[NC];
    __shared__ unsigned int s_q[4];  // queue lengths: 0 claimants, 1 claimed cells
    __shared__ int s_dxy[16], s_off[8];  // dx[8] dy[8]; shared-memory offset of neighbour d
    __shared__ unsigned int s_n[4];  // 1 lost, 2 base, 3 contested cells
    uint16_t* s_ct = s_lost;         // contested cells; consumed before the first loser is listed
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (tid < 4) s_n[tid] = s_q[tid] = 0;
    if (tid < 8) { s_dxy[tid] = DX(tid); s_dxy[8 + tid] = DY(tid); s_off[tid] = DY(tid) * SW + DX(tid); }
    load_tile_vec<TW, TH, 2>(p, p.tfB, s_tf, x0, y0);
    for (int i = tid; i < SH * SW / 16; i += NT) reinterpret_cast<uint4*>(s_cl)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    scatter_claims<TW, TH, 16>(s_tf, s_cl, s_dxy);
    __syncthreads();
    // 8 cells per thread: one pass over the tile.  Claimants are queued from the front of s_work, claimed
    // cells from the back -- two homogeneous queues -- with slots from one packed warp scan.
    for (int i = tid; i < NC / 8; i += NT) {
        const int ly = (i * 8) / TW, lx = (i * 8) - ly * TW;
        const int si = (ly + 2) * SW + lx + 16;
        const uint2 t8 = *reinterpret_cast<const uint2*>(s_tf + si);
        const uint2 c8 = *reinterpret_cast<const uint2*>(s_cl + si);
        reinterpret_cast<uint2*>(s_out)[i] = make_uint2(t8.x & 0x87878787u, t8.y & 0x87878787u);  // OCC | GHOST | trait
        const uint32_t claim[2] = {(t8.x << 1) & 0x80808080u, (t8.y << 1) & 0x80808080u};
        const uint32_t tgt[2] = {nonzero_bytes(c8.x), nonzero_bytes(c8.y)};
        append_contested(c8.x, si, s_ct, &s_n[3]);
        append_contested(c8.y, si + 4, s_ct, &s_n[3]);
        uint32_t off[2];
        wreserve<2, 16>(s_q, (uint32_t)(__popc(claim[0]) + __popc(claim[1]))
                                 | ((uint32_t)(__popc(tgt[0]) + __popc(tgt[1])) << 16), off);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            for (uint32_t m = claim[h]; m; m &= m - 1) s_work[off[0]++] = (uint16_t)(i * 8 + h * 4 + ((__ffs(m) - 1) >> 3));
            for (uint32_t m = tgt[h]; m; m &= m - 1) s_work[NC - 1 - off[1]++] = (uint16_t)(i * 8 + h * 4 + ((__ffs(m) - 1) >> 3));
        }
    }
    append_contested_ring<TW, TH>(s_cl, s_ct, &s_n[3]);
    __syncthreads();
    for (unsigned int k = tid; k < s_n[3]; k += NT) winner_of<TW>(p, s_cl, s_ct[k], x0, y0, TravelKey());  // caches the winner
    __syncthreads();
    const unsigned int n_claim = s_q[0], n_tgt = s_q[1];
    unsigned int n_moved = 0;
    for (unsigned int k = tid; k < n_claim + n_tgt; k += NT) {
        const bool is_claim = k < n_claim;
        const int cl = is_claim ? s_work[k] : s_work[NC - 1 - (k - n_claim)];
        const int ly = cl / TW, lx = cl - ly * TW;
        const int si = (ly + 2) * SW + lx + 16;
        const uint32_t t = s_tf[si];
        if (is_claim) {
            const int d = (int)((t & TF_DIR) >> TF_DIR_SHIFT);
            const int st = si + s_off[d];
            if (winner_of<TW>(p, s_cl, st, x0, y0, TravelKey()) == 7 - d) {
                s_out[cl] = (uint8_t)TF_GHOST;  // moved away; still blocks movers this step (B-13)
            } else {                            // :940-943 -> retry with the same roll
                const unsigned int slot = satom_add(&s_n[1], 1u);
                s_lost[slot] = (uint16_t)cl;
                s_lost_state[slot] = (uint8_t)(0xC0u | (uint32_t)d);
            }
        } else {  // the winner relocates here (:946-956)
            const int j = winner_of<TW>(p, s_cl, si, x0, y0, TravelKey());
            const uint32_t tn = s_tf[si + s_off[j]];
            const uint32_t x = x0 + lx, y = y0 + ly;
            const uint32_t src = nb_local(p, x, y, j);
            const size_t cell = ((size_t)y << p.log2W) | x;
            p.E1[cell] = p.E1[src];
            p.strat[cell] = p.strat[src];
            s_out[cl] = (uint8_t)(TF_OCC | (tn & TF_TRAIT));
            ++n_moved;
        }
    }
    __syncthreads();
    for (int i = tid; i < NC / 16; i += NT) {
        const int ly = (i * 16) / TW, lx = (i * 16) - ly * TW;
        *reinterpret_cast<uint4*>(p.tfA + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<uint4*>(s_out)[i];
    }
    const unsigned int n_lost = s_n[1];
    if (n_lost) {  // uniform across the CTA
        if (tid == 0) s_n[2] = atomicAdd(&p.ctr->mv_n, n_lost);
        __syncthreads();
        copy_cells<TW>(p, s_lost, n_lost, s_n[2], p.mv_cell, s_lost_state, p.mv_state, x0, y0);
    }
    if (p.stats && n_moved && row_owned(p, y0)) atomicAdd(&p.ctr->st_moved, (unsigned long long)n_moved);
}

// ------------------------------------------------------------------------------------------
// k_repro_claim (dense sweep 3): neighbourhood_update (:975-1031) + first god_go_forth
// (:1042-1134).  GHOSTs are free cells again.
// ------------------------------------------------------------------------------------------
template <int TW, int TH>
__global__ void __launch_bounds__(NT) k_repro_claim(const __grid_constant__ Params p) {
    constexpr int SW = TW + 32, SH = TH + 2, NC = TW * TH;
    __shared__ __align__(16) uint8_t s_tf[SH * SW];
    __shared__ __align__(16) uint8_t s_out[NC];
    __shared__ uint16_t s_par[NC];
    __shared__ unsigned int s_n[2];
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (tid < 2) s_n[tid] = 0;
    // the energies are requested before the flag tile arrives, so that the two latencies overlap (empty cells
    // hold 0 in E1: loading them is only wasted bandwidth, and only in sparse worlds)
    constexpr int NIT = (NC / 4 + NT - 1) / NT;
    float4 e_pre[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int i = tid + it * NT;
        e_pre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < NC / 4 && p.max_children > 0u) {
            const int ly = (i * 4) / TW, lx = (i * 4) - ly * TW;
            e_pre[it] = *reinterpret_cast<const float4*>(p.E1 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx);
        }
    }
    load_tile_vec<TW, TH, 1>(p, p.tfA, s_tf, x0, y0);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int i = tid + it * NT;
        if (i >= NC / 4) break;
        const int ly = (i * 4) / TW, lx = (i * 4) - ly * TW;
        const int si = (ly + 1) * SW + lx + 16;
        const uint32_t t4 = *reinterpret_cast<const uint32_t*>(s_tf + si);
        const uint32_t occ4 = t4 & 0x80808080u;
        reinterpret_cast<uint32_t*>(s_out)[i] = t4 & (occ4 | (occ4 >> 6) | (occ4 >> 7));  // OCC | trait of occupied cells
        if (occ4 && p.max_children > 0u) {  // :1054
            const float4 e4 = e_pre[it];
            // :978-979
            const uint32_t par4 = occ4 & ((e4.x >= p.reproduce_min_energy ? 0x80u : 0u) | (e4.y >= p.reproduce_min_energy ? 0x8000u : 0u)
                                          | (e4.z >= p.reproduce_min_energy ? 0x800000u : 0u)
                                          | (e4.w >= p.reproduce_min_energy ? 0x80000000u : 0u));
            for (uint32_t m = par4; m; m &= m - 1) s_par[satom_add(&s_n[0], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
        }
    }
    __syncthreads();
    const unsigned int n_par = s_n[0];
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
        if (dir >= 0) s_out[cl] = (uint8_t)(s_out[cl] | TF_CLAIM | ((uint32_t)dir << TF_DIR_SHIFT));
    }
    __syncthreads();
    for (int i = tid; i < NC / 16; i += NT) {
        const int ly = (i * 16) / TW, lx = (i * 16) - ly * TW;
        *reinterpret_cast<uint4*>(p.tfB + ((size_t)(y0 + ly) << p.log2W) + x0 + lx) = reinterpret_cast<uint4*>(s_out)[i];
    }
}

// ------------------------------------------------------------------------------------------
// k_repro_commit (dense sweep 4): god_multiply round 1 (:1144-1335) + environmental_punishment
// (:1339-1371) + step_fn's counts (:1405-1419).  tfB (2-halo), E1, strat -> tfA, E0, strat.
// The dense scan handles the common cells (empty, or occupied without a claim) inline and queues
// the rare ones -- claimants and cells that receive a child -- which are then processed with full
// lanes and patched into the output.  Claim losers go to the repro-loser list; they pay the living cost
// here too unless it would kill them (then it is deferred to k_finalize, see below).
// An agent that dies of living cost keeps its cell (DYING) until the retry rounds are over, exactly
// like in the reference where the punishment layer runs after the god submodel.
// ------------------------------------------------------------------------------------------
template <int TW, int TH>
__global__ void __launch_bounds__(NT, 6) k_repro_commit(const __grid_constant__ Params p) {
    constexpr int SW = TW + 32, SH = TH + 4, NC = TW * TH;
    __shared__ __align__(16) uint8_t s_tf[SH * SW];
    __shared__ __align__(16) uint8_t s_cl[SH * SW];
    __shared__ uint16_t s_work[NC];   // claimants of this tile from the front, birth cells from the back
    // cells that received a child (for the newborn list) with the child's payload.  A birth needs an empty
    // cell of the tile AND a distinct parent in tile + 1-ring, hence at most (NC + ring) / 2 per tile.
    constexpr int NBMAX = NC / 2 + TW + TH + 8;
    __shared__ uint16_t s_birth[NBMAX];
    __shared__ float s_birth_e[NBMAX];
    __shared__ uint16_t s_birth_g[NBMAX];
    __shared__ uint16_t s_lost[NC];
    __shared__ uint8_t s_lost_state[NC];
    __shared__ uint16_t s_risk[NC];
    __shared__ uint16_t s_dead[NC];
    __shared__ unsigned int s_bins[16];
    __shared__ unsigned int s_n[8];   // 0 claimants, 1 births listed, 2 lost, 3 risk, 4 dead, 5 n_old, 6 birth cells, 7 living deaths
    __shared__ unsigned int s_base[4];
    __shared__ unsigned int s_nct;   // contested cells
    uint16_t* s_ct = s_lost;         // their list; consumed before the first loser is listed
    __shared__ int s_dxy[16], s_off[8];  // dx[8] dy[8]; shared-memory offset of neighbour d
    const int tid = threadIdx.x;
    const uint32_t x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (tid < 8) s_n[tid] = 0;
    if (tid == 0) s_nct = 0;
    if (tid < 16) s_bins[tid] = 0;
    if (tid < 8) { s_dxy[tid] = DX(tid); s_dxy[8 + tid] = DY(tid); s_off[tid] = DY(tid) * SW + DX(tid); }
    // pull the tile's energies into L2 now: they are read two barriers later, and the kernel has no registers
    // to spare for holding them (6 CTAs/SM)
    for (int i = tid; i < NC / 32; i += NT) {
        const int ly = (i * 32) / TW, lx = (i * 32) - ly * TW;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.E1 + ((size_t)(y0 + ly) << p.log2W) + x0 + lx));
    }
    load_tile_vec<TW, TH, 2>(p, p.tfB, s_tf, x0, y0);
    for (int i = tid; i < SH * SW / 16; i += NT) reinterpret_cast<uint4*>(s_cl)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    scatter_claims<TW, TH, 4>(s_tf, s_cl, s_dxy);
    __syncthreads();
    // ---- dense scan, 4 cells per thread, branch-free per cell; the rare cells (claimants, birth cells, at-risk,
    // dying) are appended to their queues from bit masks
    unsigned int n_old = 0, n_dead = 0;
    for (int i = tid; i < NC / 4; i += NT) {
        const int ly = (i * 4) / TW, lx = (i * 4) - ly * TW;
        const int si = (ly + 2) * SW + lx + 16;
        const size_t cell0 = ((size_t)(y0 + ly) << p.log2W) + x0 + lx;
        const uint32_t t4 = *reinterpret_cast<const uint32_t*>(s_tf + si);
        const uint32_t c4 = *reinterpret_cast<const uint32_t*>(s_cl + si);
        const uint32_t occ4 = t4 & 0x80808080u;
        const uint32_t claim4 = (t4 << 1) & 0x80808080u;  // parents with a claim: decided in the queue below
        float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t st4 = 0;
        if (occ4) {
            e4 = *reinterpret_cast<const float4*>(p.E1 + cell0);
            if (p.want_counts) st4 = *reinterpret_cast<const uint32_t*>(p.strat + cell0);
        }
        float ev[4] = {e4.x, e4.y, e4.z, e4.w};
        const uint32_t pay4 = occ4 & ~claim4;               // everybody else pays the cost of living now (:1357-1368)
        const uint32_t born4 = nonzero_bytes(c4) & ~occ4;   // a child is born in an empty cell that received a claim
        uint32_t dead4 = 0, risk4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bit = 0x80u << (8 * k);
            float e = ev[k];
            const bool alive = punish(p, e);
            if (pay4 & bit) {
                // a death (:1364-1366) keeps its cell for the reproduction retry rounds (the punishment layer runs
                // AFTER the god submodel); k_finalize clears it
                if (!alive) dead4 |= bit;
                else if (e <= p.risk_thr) risk4 |= bit;
                ev[k] = alive ? e : 0.0f;
            } else if (!(claim4 & bit)) {
                ev[k] = 0.0f;
            }
        }
        // OCC | trait for the survivors, OCC | DYING for the dying, 0 for claimants (patched below) and empty cells
        static_assert(TF_OCC == 0x80u && TF_DYING == 0x08u && TF_TRAIT == 0x03u && TF_CLAIM == 0x40u, "flag layout");
        const uint32_t live4 = pay4 & ~dead4;
        const uint32_t out4 = (t4 & (live4 | (live4 >> 6) | (live4 >> 7))) | dead4 | (dead4 >> 4);
        n_old += __popc(occ4);
        n_dead += __popc(dead4);
        if (p.want_counts)
            for (uint32_t m = live4; m; m &= m - 1) {
                const int b = (__ffs(m) - 1) >> 3;
                atomicAdd(&s_bins[count_bin((st4 >> (8 * b)) & 0xFFu, (t4 >> (8 * b)) & TF_TRAIT)], 1u);
            }
        for (uint32_t m = claim4; m; m &= m - 1) s_work[satom_add(&s_n[0], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
        for (uint32_t m = born4; m; m &= m - 1) s_work[NC - 1 - satom_add(&s_n[6], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
        for (uint32_t m = risk4; m; m &= m - 1) s_risk[satom_add(&s_n[3], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
        for (uint32_t m = dead4; m; m &= m - 1) s_dead[satom_add(&s_n[4], 1u)] = (uint16_t)(i * 4 + ((__ffs(m) - 1) >> 3));
        append_contested(c4, si, s_ct, &s_nct);
        *reinterpret_cast<uint32_t*>(p.tfA + cell0) = out4;
        *reinterpret_cast<float4*>(p.E0 + cell0) = make_float4(ev[0], ev[1], ev[2], ev[3]);
    }
    append_contested_ring<TW, TH>(s_cl, s_ct, &s_nct);
    __syncthreads();  // the queues are complete; the dense stores above are ordered before the patches below
    for (unsigned int k = tid; k < s_nct; k += NT) winner_of<TW>(p, s_cl, s_ct[k], x0, y0, ReproKey());  // caches the winner
    __syncthreads();
    const bool owned_tile = row_owned(p, y0);
    // two homogeneous queues (no divergence between the two kinds of work inside a warp)
    const unsigned int n_claim = s_n[0], n_born = s_n[6];
    for (unsigned int k = tid; k < n_claim + n_born; k += NT) {
        // claimants first, then birth cells; the boundary falls inside at most one warp
        const bool is_claim = k < n_claim;
        const int cl = is_claim ? s_work[k] : s_work[NC - 1 - (k - n_claim)];
        const int ly = cl / TW, lx = cl - ly * TW;
        const int si = (ly + 2) * SW + lx + 16;
        const uint32_t x = x0 + lx, y = y0 + ly;
        const size_t cell = ((size_t)y << p.log2W) | x;
        const uint32_t t = s_tf[si];
        if (is_claim) {
            const uint32_t trait = t & TF_TRAIT;
            const int d = (int)((t & TF_DIR) >> TF_DIR_SHIFT);
            const int st = si + s_off[d];
            const float e_in = p.E1[cell];
            const bool won = winner_of<TW>(p, s_cl, st, x0, y0, ReproKey()) == 7 - d;
            // Winner: reproduction cost (:1211-1213), then the cost of living.  Loser (stays
            // ATTEMPTING_REPRODUCTION, :1203-1205, on the retry list): it is charged the cost of living right
            // away, as if it never reproduces -- true for all losers in the saturated state, where no retry round
            // runs; if it wins a later round, k_repro_retry recomputes its energy from E1 in the reference's layer
            // order.  Only a loser the living cost would KILL is deferred to k_finalize: it must stay a live
            // parent for the retry rounds.  One code path for both keeps the warp converged.
            float e = won ? __fsub_rn(e_in, p.reproduce_cost) : e_in;
            const bool alive = punish(p, e);
            if (alive) {
                p.tfA[cell] = (uint8_t)(TF_OCC | trait);
                p.E0[cell] = e;
                if (e <= p.risk_thr) s_risk[satom_add(&s_n[3], 1u)] = (uint16_t)cl;
                if (p.want_counts) atomicAdd(&s_bins[count_bin(p.strat[cell], trait)], 1u);
            } else if (won) {
                p.tfA[cell] = (uint8_t)(TF_OCC | TF_DYING);
                p.E0[cell] = 0.0f;
                s_dead[satom_add(&s_n[4], 1u)] = (uint16_t)cl;
                satom_add(&s_n[7], 1u);
            } else {
                p.tfA[cell] = (uint8_t)(TF_OCC | trait);
                p.E0[cell] = e_in;
                if (owned_tile) {
                    const unsigned int dslot = atomicAdd(&p.ctr->dl_n, 1u);
                    if (dslot < p.list_cap) p.dl_cell[dslot] = (uint32_t)cell;
                    else atomicExch(&p.ctr->overflow, 1u);
                }
            }
            if (!won) {
                const unsigned int slot = satom_add(&s_n[2], 1u);
                s_lost[slot] = (uint16_t)cl;
                s_lost_state[slot] = (uint8_t)(0xC0u | (uint32_t)d);
            }
        } else {  // a child is born here (:1216-1328): the winner's trait, its own energy and strategies
            const int j = winner_of<TW>(p, s_cl, si, x0, y0, ReproKey());
            const uint32_t tn = s_tf[si + s_off[j]];
            const uint32_t src = nb_local(p, x, y, j);
            const float pe = __fsub_rn(p.E1[src], p.reproduce_cost);
            float ce;
            uint32_t cs;
            make_child(p, nb_gcell(p, x, y, j), p.strat[src], tn & TF_TRAIT, pe, ce, cs);
            if (owned_tile) {
                // Not written into the planes yet: near the carrying capacity almost every newborn is culled
                // in the same step (k_finalize), so it only goes on the newborn list with its payload.  If
                // retry rounds run, k_repro_retry first marks the listed cells as occupied.
                const unsigned int slot = satom_add(&s_n[1], 1u);
                if (slot < (unsigned int)NBMAX) {
                    s_birth[slot] = (uint16_t)cl;
                    s_birth_e[slot] = ce;
                    s_birth_g[slot] = (uint16_t)(cs | ((tn & TF_TRAIT) << 8));
                } else atomicExch(&p.ctr->overflow, 1u);
            } else {  // ghost tile of a slab: the owner lists it; here it only has to block its cell
                p.strat[cell] = (uint8_t)cs;
                p.E0[cell] = ce;
                p.tfA[cell] = (uint8_t)(TF_OCC | TF_NEWBORN | (tn & TF_TRAIT));
            }
        }
    }
    if (n_old) satom_add(&s_n[5], n_old);
    if (n_dead) satom_add(&s_n[7], n_dead);
    __syncthreads();
    // one atomic per list per CTA, then plain copies.  A ghost tile of a slab only contributes its claim
    // losers (the retry rounds need them); its newborns, deaths and counts belong to the owning slab.
    const bool owned = row_owned(p, y0);
    const unsigned int n_lost = s_n[2];
    const unsigned int n_birth = owned ? (s_n[1] < (unsigned int)NBMAX ? s_n[1] : (unsigned int)NBMAX) : 0u;
    const unsigned int n_risk = owned ? s_n[3] : 0u, n_dying = owned ? s_n[4] : 0u;
    if (tid < 4) {
        const unsigned int n = tid == 0 ? n_lost : tid == 1 ? n_birth : tid == 2 ? n_risk : n_dying;
        unsigned int* gc = tid == 0 ? &p.ctr->rp_n : tid == 1 ? &p.ctr->nb_n
                         : tid == 2 ? &p.ctr->risk_next_n : &p.ctr->dead_n;
        s_base[tid] = n ? atomicAdd(gc, n) : 0u;
    }
    __syncthreads();
    copy_cells<TW>(p, s_lost, n_lost, s_base[0], p.rp_cell, s_lost_state, p.rp_state, x0, y0);
    copy_cells<TW>(p, s_birth, n_birth, s_base[1], p.nb_cell, nullptr, nullptr, x0, y0);
    if (n_birth && s_base[1] + n_birth <= p.list_cap)
        for (unsigned int k = tid; k < n_birth; k += NT) {
            p.nb_e[s_base[1] + k] = s_birth_e[k];
            p.nb_gene[s_base[1] + k] = s_birth_g[k];
        }
    copy_cells<TW>(p, s_risk, n_risk, s_base[2], p.risk_next, nullptr, nullptr, x0, y0);
    copy_cells<TW>(p, s_dead, n_dying, s_base[3], p.dead_cell, nullptr, nullptr, x0, y0);
    if (owned) {
        if (tid == 0) {
            if (s_n[5]) atomicAdd(&p.ctr->n_old, (unsigned long long)s_n[5]);
            if (n_birth) atomicAdd(&p.ctr->births, (unsigned long long)n_birth);
            if (s_n[7]) atomicAdd(&p.ctr->living_deaths, (unsigned long long)s_n[7]);
        }
        if (p.want_counts && tid < 16 && s_bins[tid]) atomicAdd(&p.ctr->bins[tid], (unsigned long long)s_bins[tid]);
    }
}

}  // namespace pd

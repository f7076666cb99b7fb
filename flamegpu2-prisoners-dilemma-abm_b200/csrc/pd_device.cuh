// Shared device-side definitions: plane bit layout, kernel parameter block, Philox
// stream ids, and the per-agent rules (game, probe, child) every kernel uses.
//
// Citations are /root/reference/src/model.py.  The rules implement the EFFECTIVE
// semantics written down in DESIGN.md ("Semantics"); the NumPy oracle (oracle/pd_oracle.py)
// restates the same reference functions independently on an agent list.
#pragma once
#include <stdint.h>
#include "philox.cuh"

namespace pd {

// ---- tf plane bits ---------------------------------------------------------------------
// Canonical (between steps): OCC | trait.  Transient meanings per phase:
//   tfA during play      : bits 3-6 = D code of an at-risk agent (0 none, c+1 = dies in sub-step c)
//   tfB during travel     : GHOST = cell was occupied at step start but is empty now (still blocks movers,
//                          quirk B-13)
//   tfB during reproduce  : NEWBORN (same bit as GHOST, but with OCC) = child of this step (only written when
//                          retry rounds must see its cell)
// Claims are not flags: a claimant posts its key with an atomic max into the claim plane (claim_word below).
constexpr uint32_t TF_TRAIT = 0x03u;
constexpr uint32_t TF_GHOST = 0x04u;
constexpr uint32_t TF_NEWBORN = 0x04u;
constexpr uint32_t TF_OCC = 0x80u;
constexpr uint32_t TF_D_SHIFT = 3u;
constexpr uint32_t TF_D = 0x78u;

// ---- Philox stream ids (counter word 3); must equal oracle/pd_oracle.py -----------------
constexpr uint32_t S_ROLL = 0;    // ONE call per four cells in a row: counter = cell >> 2, the cell's word = word cell & 3.
                                  // Top 16 bits: pair-ordering roll (:364; a tie is won by the neighbour in a direction
                                  // >= 4, which replaces the id tie-break of :413).  Bits 2c / 2c+1: the challenger's /
                                  // my RANDOM coin (:631 / :664) of the game I respond to in sub-step c -- a game
                                  // without noise needs no random call of its own
constexpr uint32_t S_NOISE0 = 1;  // +(c>>1); keyed by the RESPONDER's cell: even c: x my noise roll (:611) y the
                                  //     challenger's (:612); odd c: z mine, w the challenger's
constexpr uint32_t S_TRAVEL = 14; // x: travel priority (:806)  y>>29: travel start dir (:810)
constexpr uint32_t ROLL_SHIFT = 16;
constexpr uint32_t S_REPRO = 9;   // x: reproduction claim priority (:1076)  y>>29: start dir (:1079)
constexpr uint32_t S_MUT = 10;    // mutation rolls (:1263, :1278, :1291-1292), keyed by the parent's cell
constexpr uint32_t S_PICK = 11;   // replacement picks (:1265, :1281, :1303, :1314)
constexpr uint32_t S_CULL = 12;   // x: newborn cull priority, keyed by the child's cell (replaces :1343/:1354)
constexpr uint32_t S_ENERGY = 13; // child energy normal deviate (:1236), keyed by the parent's cell
constexpr uint32_t INIT_STEP = 0xFFFFFFFFu;
constexpr uint32_t I_PLACE = 0;   // x: placement priority (:1438-1450)  y>>30: trait (:1467)
constexpr uint32_t I_STRAT = 1;   // strategy draws (:1473-1497)
constexpr uint32_t I_ENERGY = 2;  // initial energy normal deviate (:1460-1465)

// 1 / (2^32 * sqrt(24 + 1/12)) -- scale of the integer-exact normal stand-in
#define PD_NORMAL_SCALE 0x1.a15286234fcddp-35

// Moore offsets (:302-307); the opposite of direction i is 7-i.
// x: {-1,-1,-1,0,0,1,1,1}  y: {-1,0,1,-1,1,-1,0,1} -- row-major 3x3 minus the centre, so with
// e = d + (d >= 4):  dx = e / 3 - 1,  dy = e % 3 - 1  (folds to constants in unrolled loops).
__host__ __device__ constexpr int DX(int d) { return (d + (d >= 4 ? 1 : 0)) / 3 - 1; }
__host__ __device__ constexpr int DY(int d) { return (d + (d >= 4 ? 1 : 0)) % 3 - 1; }
static_assert(DX(0) == -1 && DX(2) == -1 && DX(3) == 0 && DX(4) == 0 && DX(5) == 1 && DX(7) == 1, "dx");
// the same for a direction only known at run time: two bits per direction, packed
__host__ __device__ constexpr int dxr(int d) { return (int)((0xA940u >> (2 * d)) & 3u) - 1; }
__host__ __device__ constexpr int dyr(int d) { return (int)((0x9224u >> (2 * d)) & 3u) - 1; }
static_assert(dxr(0) == DX(0) && dxr(1) == DX(1) && dxr(2) == DX(2) && dxr(3) == DX(3) && dxr(4) == DX(4) &&
              dxr(5) == DX(5) && dxr(6) == DX(6) && dxr(7) == DX(7), "dxr");
static_assert(dyr(0) == DY(0) && dyr(1) == DY(1) && dyr(2) == DY(2) && dyr(3) == DY(3) && dyr(4) == DY(4) &&
              dyr(5) == DY(5) && dyr(6) == DY(6) && dyr(7) == DY(7), "dyr");
static_assert(DY(0) == -1 && DY(1) == 0 && DY(2) == 1 && DY(3) == -1 && DY(4) == 1 && DY(5) == -1 &&
              DY(6) == 0 && DY(7) == 1, "dy");

// ---- device-resident counters -----------------------------------------------------------
struct Counters {
    // zeroed at the start of every step (k_begin_step)
    unsigned int mv_n, rp_n, nb_n, risk_next_n, cand_n, nb_flagged, dep_n;
    unsigned int lc_n;  // reproduction claims k_claim_punish resolved inside its tiles: listed from the BACK of the claim list
    unsigned int rt_n[2], pad4_[2];  // retry list: [0] travel, [1] reproduction
    unsigned int mv_active[8], rp_active[8];
    unsigned long long n_old, admitted, living_deaths, pad2_;  // admitted = newborns that entered the planes
    unsigned long long st_play_deaths, st_movers, st_moved, st_travel_deaths;
    unsigned long long bins[16];
    unsigned int hist[8][256];
    unsigned long long tau;
    // persistent across steps
    unsigned int risk_n;
    unsigned int overflow;
    unsigned int barrier[2];
    unsigned int log_n, log_dropped;  // rows in the step log / rows that did not fit
    unsigned int step_no;  // model step being computed: the Philox key of every draw.  Device-resident so that
    unsigned int pad3_;    // a captured CUDA graph of one step can be replayed (k_set_step / fin_bookkeeping)
    unsigned long long n_agents;
    unsigned long long agents_start;
    unsigned long long agent_steps;  // running sum of the population at the start of every step
    unsigned long long last_bins[16];
    unsigned long long last_stats[12];
};

// ---- kernel parameter block (lives in the constant bank as a __grid_constant__ arg) ------
struct Params {
    uint32_t W, H, row0, log2W;
    uint32_t HW;  // height of the whole world (power of two; == W for the reference's square worlds)
    uint32_t k0, k1;
    float pay[4];  // index (i_coop << 1) | challenger_coop: {dd, dc, cd, cc} seen by the responder
    float max_energy, travel_cost, cost_of_living, reproduce_min_energy, reproduce_cost;
    float inherit, e_mu, e_sigma, e_min, risk_thr, mutation_rate;
    unsigned long long noise_thr, mut_thr, strat_thr[3];
    unsigned long long max_agents;
    uint32_t pure, per_trait, max_children, want_counts, stats;
    uint32_t list_cap, sparse_ctas;
    // row-slab decomposition (DESIGN.md "Multi-GPU"): the local planes hold `H` rows = owned rows plus
    // a ghost zone above and below; rows wrap only when the handle owns the whole world
    uint32_t wrap_y;          // 1: H == W, rows wrap (whole world); 0: slab with ghost rows, no wrap
    uint32_t own_y0, own_y1;  // owned local rows [own_y0, own_y1); counters/cull only see these
    uint32_t round0, round1;  // reproduction retry rounds to run in this launch (slab mode)
    uint32_t speculative;     // k_fin_cand / k_fin_apply: do nothing if the global gate says a retry round is due
    uint32_t world;           // number of slabs
    uint32_t cand_cap;        // capacity (keys) of one rank's candidate block
    unsigned long long* step_log;  // [log_cap][18] rows: step, agents, 16 counts; nullptr = off
    uint32_t log_cap;
    uint32_t select_cap;      // cull select: most keys resolved in shared memory (tests lower it to reach the fallback)
    uint32_t tie_shift;       // k_claim_punish: local claims whose priorities agree above this bit are a near-tie (3;
                              // tests raise it to reach the detour through the claim plane)
    unsigned long long* glb;        // [0] olds [1] births [2] active repro losers [3] agents [4..19] bins,
                                    // summed over all slabs by the caller between the step's parts
    unsigned int* ghist;            // 2048-bin cull histogram, summed over all slabs by the caller
    unsigned long long* cand_out;   // this slab's cull candidates: [0] = count, then keys
    const unsigned long long* cand_all;  // all slabs' candidate blocks, gathered by the caller
    const float* lut;  // 256-entry payoff table of the play sweep (lut_entry, pd_dense.cuh)
    // planes
    float* E0;       // caller-bound energy (state between steps)
    float* E1;       // scratch energy (state between play and the final sweep)
    uint8_t* strat;  // caller-bound
    uint8_t* tfA;    // caller-bound
    uint8_t* tfB;    // scratch
    // sparse lists
    uint32_t* risk_cur; uint32_t* risk_next; float* risk_e;
    unsigned long long* cand;  // cull select: the threshold bin's candidate keys (list_cap / 2 of them)
    // One 64-bit word per cell: the highest claim posted on it (claim_word; atomic max, never cleared -- newer
    // rounds carry a higher tag).  Replaces the move / reproduction request messages keyed by the target's bucket
    // (:864-868, :1126-1130) and the "highest roll wins" scan of their readers (:929, :1192).
    unsigned long long* claim;
    // Round 1 claimants (the travel phase's, then the reproduction phase's).  An entry carries what the commit
    // needs of the claimant, so that it costs the commit no random reads: cell, energy (E1 of the cell), and
    // aux = claimed direction | trait << 3 | strategies << 5 (strategies: travel only)
    uint32_t* cl_cell; float* cl_e; uint16_t* cl_aux;
    uint32_t* rt_cell; uint8_t* rt_state; uint8_t* rt_aux;  // the losers of round 1: rounds 2..8 (travel, then reproduction)
    // newborns: the child's cell and the direction the parent claimed (the parent is the child's neighbour 7 - dir).
    // The child itself -- energy, strategies, trait -- is made from the parent's (unchanged) E1 / strat / tfB only if
    // it survives the cull: near the carrying capacity 99 % do not
    uint32_t* nb_cell; uint8_t* nb_dir;
    uint32_t* nb_prio;  // the newborn's cull priority (S_CULL word of its cell): drawn once, when it is born, and
                        // counted into the cull histogram (Counters::hist, 2048 bins of its top 11 bits) on the spot
    Counters* ctr;
};

__device__ __forceinline__ Words rng(const Params& p, uint64_t gcell, uint32_t stream) {
    // the step number only changes between kernels (end of k_finalize / k_fin_apply): read-only here
    return philox4x32_10((uint32_t)gcell, (uint32_t)(gcell >> 32), __ldg(&p.ctr->step_no), stream, p.k0, p.k1);
}
__device__ __forceinline__ Words rng_init(const Params& p, uint64_t gcell, uint32_t stream) {
    return philox4x32_10((uint32_t)gcell, (uint32_t)(gcell >> 32), INIT_STEP, stream, p.k0, p.k1);
}

// global cell index (x + y*env_max, :339) of local cell (x, y) and of its neighbours
__device__ __forceinline__ uint64_t gcell_of(const Params& p, uint32_t x, uint32_t y) {
    return ((uint64_t)((p.row0 + y) & (p.HW - 1)) << p.log2W) | x;
}
// local row of the neighbour row y + dy: wraps for a whole world (:311-312); a slab never wraps --
// its outermost ghost rows are outside the zone whose results are used, so the index is just clamped
__device__ __forceinline__ uint32_t nb_row(const Params& p, uint32_t y, int dy) {
    if (p.wrap_y) return (y + (uint32_t)dy) & (p.H - 1);
    const int r = (int)y + dy;
    return (uint32_t)(r < 0 ? 0 : (r >= (int)p.H ? (int)p.H - 1 : r));
}
// local plane index of the neighbour of local (x,y) in direction d
__device__ __forceinline__ uint32_t nb_local(const Params& p, uint32_t x, uint32_t y, int d) {
    const uint32_t nx = (x + (uint32_t)DX(d)) & (p.W - 1);
    return (nb_row(p, y, DY(d)) << p.log2W) | nx;
}
// the parent of the newborn at `cell`: it claimed direction `dir`, so it is the child's neighbour 7 - dir
__device__ __forceinline__ uint32_t parent_of(const Params& p, uint32_t cell, uint32_t dir) {
    return nb_local(p, cell & (p.W - 1), cell >> p.log2W, 7 - (int)dir);
}
__device__ __forceinline__ bool row_owned(const Params& p, uint32_t y) {
    return y >= p.own_y0 && y < p.own_y1;
}
__device__ __forceinline__ uint64_t nb_gcell(const Params& p, uint32_t x, uint32_t y, int d) {
    const uint32_t nx = (x + (uint32_t)DX(d)) & (p.W - 1);
    const uint32_t gy = (p.row0 + y + (uint32_t)DY(d)) & (p.HW - 1);
    return ((uint64_t)gy << p.log2W) | nx;
}

// ---- claims (move_request :864-868 / move_response :929-935; god_go_forth :1126-1130 / god_multiply :1192-1198) --
// A claimant posts ONE word on its target cell with an atomic max and has won iff the cell still holds exactly
// its word when the round is resolved.  Layout, most significant first:
//     tag  (29 bits)  (step & 0x1FFFFFF) << 4 | phase << 3 | round   -- phase 0 travel, 1 reproduction; round 0..7
//     prio (32 bits)  the claimant's Philox priority (:806, :1076): the highest roll wins
//     j    (3 bits)   7 - claimed direction = where the claimant stands as seen from the cell.  Unique among the
//                     rivals of one cell; decides a tie of the priorities (it takes the place of the id comparison
//                     of :929 / :1192, like the direction class does for the play rolls)
// The tag grows with every round, so a word left over from an earlier round, phase or step always loses: the plane
// is never cleared (the host zeroes it when the step counter's 25 bits wrap, and whenever the step is set).
// (A plane of one BYTE per cell -- a bit per claimant position, OR-ed in, rivals decided by recomputed priorities,
// cleared by the winner -- is 8 x smaller and was 1 % faster, but row slabs at 2^28 cells then disagreed with the
// whole world in a few cells every other run, non-deterministically and for reasons not understood; this design
// has no state to clear and passed every repetition.  profiles/README.md "byte claim plane".)
constexpr uint32_t CLAIM_TRAVEL = 0u, CLAIM_REPRO = 8u;
__device__ __forceinline__ unsigned long long claim_word(uint32_t step, uint32_t phase_round, uint32_t prio, int dir) {
    const unsigned long long tag = ((unsigned long long)(step & 0x1FFFFFFu) << 4) | phase_round;
    return (tag << 35) | ((unsigned long long)prio << 3) | (unsigned long long)(7 - dir);
}
// A reduction on a word that misses in the L2 is served far more slowly than a load (measured: 16 M reductions
// per ms against > 100 M random sector loads): the claimant asks for the word's line ahead of posting
__device__ __forceinline__ void prefetch_claim(const unsigned long long* word) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(word));
}
__device__ __forceinline__ void post_claim(unsigned long long* word, unsigned long long v) {
    asm volatile("red.global.max.u64 [%0], %1;" ::"l"(word), "l"(v) : "memory");
}

// ---- the game as the RESPONDER sees it (play_response, :599-689) --------------------------
// strategies: 0 COOP, 1 DEFECT, 2 TFT (== COOP: its memory is submodel-local and every pair
// plays once per step, :621-629/:646-661), 3 RANDOM.  `coins` = the w word of the responder's
// play word.  (The dense play sweep has its own table-driven form of this, pd_dense.cuh.)
__device__ __forceinline__ float game_payoff(const Params& p, uint64_t my_gcell, int c, uint32_t coins,
                                             uint32_t my_strat, uint32_t my_trait,
                                             uint32_t ch_strat, uint32_t ch_trait) {
    const uint32_t mine = (my_strat >> (2 * ch_trait)) & 3u;    // :599
    const uint32_t theirs = (ch_strat >> (2 * my_trait)) & 3u;  // :600
    bool i_coop = mine == 3u ? ((coins >> (2 * c + 1)) & 1u) != 0u : mine != 1u;      // :663-668
    bool ch_coop = theirs == 3u ? ((coins >> (2 * c)) & 1u) != 0u : theirs != 1u;     // :630-635
    if (p.noise_thr != 0ull) {
        const Words w = rng(p, my_gcell, S_NOISE0 + ((uint32_t)c >> 1));  // :611-612
        const uint32_t my_roll = (c & 1) ? w.z : w.x, ch_roll = (c & 1) ? w.w : w.y;
        if ((unsigned long long)ch_roll < p.noise_thr) ch_coop = !ch_coop;  // :638-640
        if ((unsigned long long)my_roll < p.noise_thr) i_coop = !i_coop;    // :672-674
    }
    return p.pay[((uint32_t)i_coop << 1) | (uint32_t)ch_coop];  // :677-689 (my side only)
}
// the play word of a cell (S_ROLL)
__device__ __forceinline__ uint32_t play_word(const Params& p, uint64_t gcell) {
    const Words w = rng(p, gcell >> 2, S_ROLL);
    const uint32_t k = (uint32_t)gcell & 3u;
    return k == 0u ? w.x : k == 1u ? w.y : k == 2u ? w.z : w.w;
}
// "the neighbour in direction d (play word rn) is higher than me (play word r)": it challenges me (:413)
__device__ __forceinline__ bool roll_higher(uint32_t rn, uint32_t r, int d) {
    const uint32_t kn = rn >> ROLL_SHIFT, k = r >> ROLL_SHIFT;
    return kn > k || (kn == k && d >= 4);
}

// ---- the newborn cull (replaces the list-index rule :1343/:1354, SURVEY B-4) ----------------
// priority word of the newborn at local cell `cell`, and its 64-bit key (priority, global cell): higher survives
__device__ __forceinline__ uint32_t cull_prio(const Params& p, uint32_t cell) {
    return rng(p, gcell_of(p, cell & (p.W - 1), cell >> p.log2W), S_CULL).x;
}
__device__ __forceinline__ unsigned long long cull_key_of(const Params& p, uint32_t prio, uint32_t cell) {
    return ((unsigned long long)prio << 32) | (uint32_t)gcell_of(p, cell & (p.W - 1), cell >> p.log2W);
}

// ---- triangular probe (:826-839, :1088-1101) ---------------------------------------------
// blocked: bit d set = direction d not free.  l = start direction, seq = probes used so far.
// Returns the direction found or -1; on success the caller stores next start (l+1)&7 and seq.
__device__ __forceinline__ int probe(uint32_t blocked, int& l, int& seq) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ++seq;
        if (seq > 8) return -1;
        l = (l + i) & 7;
        if (!((blocked >> l) & 1u)) return l;
    }
    return -1;
}
// number of probes a FIRST attempt starting at s needed to reach direction d
__device__ __forceinline__ int probe_pos(int s, int d) {
    int l = s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        l = (l + i) & 7;
        if (l == d) return i + 1;
    }
    return 8;
}

// ---- integer-exact N(0,1) stand-in (see oracle.normal_from_words) -------------------------
__device__ __forceinline__ float normal_from_words(const Words& w) {
    const long long pc = (long long)(__popc(w.x) + __popc(w.y) + __popc(w.z)) - 48ll;
    const long long zi = pc * 4294967296ll + ((long long)w.w - 2147483648ll);
    return __double2float_rn(__dmul_rn((double)zi, PD_NORMAL_SCALE));
}

__device__ __forceinline__ uint32_t mutate(uint32_t s, uint32_t pick_word) {
    return (s + 1u + (uint32_t)(((unsigned long long)pick_word * 3ull) >> 32)) & 3u;
}

// ---- the child (god_multiply, :1224-1326) --------------------------------------------------
// parent_energy_after_cost is only used in inheritance mode (:1242).
__device__ __forceinline__ void make_child(const Params& p, uint64_t parent_gcell,
                                           uint32_t parent_strat, uint32_t parent_trait,
                                           float parent_energy_after_cost, float& child_energy,
                                           uint32_t& child_strat) {
    float ce;
    if (p.inherit <= 0.0f || p.inherit > 1.0f) {  // :1231-1238
        const float z = normal_from_words(rng(p, parent_gcell, S_ENERGY));
        ce = __fadd_rn(__fmul_rn(z, p.e_sigma), p.e_mu);
    } else {
        ce = __fmul_rn(p.inherit, parent_energy_after_cost);  // :1242
    }
    if (ce < p.e_min) ce = p.e_min;  // :1245-1249
    else if (ce > p.max_energy) ce = p.max_energy;
    child_energy = ce;
    uint32_t cs = parent_strat;
    if (p.mutation_rate > 0.0f) {
        const Words m = rng(p, parent_gcell, S_MUT);
        const Words k = rng(p, parent_gcell, S_PICK);
        if (p.pure) {  // :1259-1271 -- the roll is drawn but never compared (quirk B-7)
            const uint32_t s = mutate(parent_strat & 3u, k.x);
            cs = s * 0x55u;
        } else if (p.per_trait) {  // :1273-1286
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
            const uint32_t kw[4] = {k.x, k.y, k.z, k.w};
            cs = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t s = (parent_strat >> (2 * i)) & 3u;
                if ((unsigned long long)mw[i] < p.mut_thr) s = mutate(s, kw[i]);
                cs |= s << (2 * i);
            }
        } else {  // kin / other, :1287-1326
            const uint32_t first_other = parent_trait == 0u ? 1u : 0u;
            uint32_t s_my = (parent_strat >> (2 * parent_trait)) & 3u;
            uint32_t s_ot = (parent_strat >> (2 * first_other)) & 3u;
            if ((unsigned long long)m.x < p.mut_thr) s_my = mutate(s_my, k.x);
            if ((unsigned long long)m.y < p.mut_thr) s_ot = mutate(s_ot, k.y);
            cs = (s_ot * 0x55u) & ~(3u << (2 * parent_trait));
            cs |= s_my << (2 * parent_trait);
        }
    } else if (p.pure) {
        cs = (parent_strat & 3u) * 0x55u;  // all four slots = slot 0 (:1260, :1268-1270)
    } else if (!p.per_trait) {
        // kin/other without mutation still rewrites every "other" slot from the first one (:1296-1323)
        const uint32_t first_other = parent_trait == 0u ? 1u : 0u;
        const uint32_t s_my = (parent_strat >> (2 * parent_trait)) & 3u;
        const uint32_t s_ot = (parent_strat >> (2 * first_other)) & 3u;
        cs = ((s_ot * 0x55u) & ~(3u << (2 * parent_trait))) | (s_my << (2 * parent_trait));
    }
    child_strat = cs & 0xFFu;
}

// strategy-count bin of an agent: 4*my + other (:1413-1419, :1501-1511), derived (B-8)
__device__ __forceinline__ uint32_t count_bin(uint32_t strat, uint32_t trait) {
    const uint32_t my = (strat >> (2 * trait)) & 3u;
    const uint32_t other = (strat >> (2 * (trait == 0u ? 1u : 0u))) & 3u;
    return 4u * my + other;
}

// ---- software grid barrier for the persistent sparse kernels ------------------------------
// Co-residency is guaranteed by cudaLaunchCooperativeKernel (or by nblk == 1).
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int nblk) {
    __syncthreads();
    if (nblk > 1) {
        if (threadIdx.x == 0) {
            volatile unsigned int* vgen = bar + 1;
            const unsigned int gen = *vgen;
            __threadfence();
            if (atomicAdd(bar, 1u) == nblk - 1u) {
                atomicExch(bar, 0u);
                __threadfence();
                atomicAdd(bar + 1, 1u);
            } else {
                while (*vgen == gen) __nanosleep(64);
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// warp-aggregated append: every lane of the (converged) warp must call this
__device__ __forceinline__ uint32_t warp_append(unsigned int* counter, bool pred) {
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0u) return 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return pred ? base + (unsigned int)__popc(m & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}

}  // namespace pd

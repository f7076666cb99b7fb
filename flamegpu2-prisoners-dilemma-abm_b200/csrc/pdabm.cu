// libpdabm.so -- host side of the C ABI declared in include/pdabm.h.
//
// Builds the kernel parameter block from pd_config (the reference's environment properties,
// /root/reference/src/model.py:1678-1743), owns the scratch planes and sparse lists, and
// issues the kernels of one model step in the order of the reference's main layers
// (model.py:2140-2163).  No torch types, no exceptions across the boundary.
#include "../../include/pdabm.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "pd_kernels.cuh"

using namespace pd;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(PD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                  \
    } while (0)

unsigned long long prob_threshold(float p) {
    // P(event) = p  <=>  (u32 word) < threshold  (same arithmetic as oracle.prob_threshold)
    const double v = (double)p * 4294967296.0;
    if (!(v > 0.0)) return 0ull;
    const double f = floor(v);
    return f >= 4294967296.0 ? 4294967296ull : (unsigned long long)f;
}

struct HostOut {  // pinned staging for read-backs
    unsigned long long bins[16];
    unsigned long long stats[12];
    unsigned long long n_agents;
    unsigned int overflow;
    unsigned int hist[256];
};

}  // namespace

struct pd_sim {
    pd_config cfg;
    Params P;
    size_t cells;
    bool bound, ready;
    uint32_t step;
    int risk_cur;
    uint32_t* risk[2];
    int n_sms;
    int sparse_ctas, sparse_threads;
    bool sparse_auto;                       // CTA count per kernel from its occupancy
    std::map<const void*, int> sparse_occ;  // kernel -> CTAs of its cooperative launch
    bool big_tiles;
    bool slab;        // row slab of a larger world (ghost rows, pd_step_part)
    bool comm_bound;
    dim3 grid;
    HostOut* h;  // pinned
    bool lists_allocated;
    // optional per-kernel timing (pd_set_profile): events around every launch of a step
    bool profile;
    // One event BETWEEN consecutive profiled launches of a call (the end of one kernel is the start of the next):
    // 10 records per step, not 18 -- every record in the stream keeps the next kernel from starting back to back
    // (measured: 18 records cost 0.065 ms per step at 2^28 cells).
    std::vector<cudaEvent_t> ev;
    struct ProfSpan { uint32_t start, end; int slot; };
    std::vector<ProfSpan> spans;   // consumed by pd_read_profile
    size_t ev_used;                // events recorded since the last read
    bool ev_chain;                 // the last recorded event ended a span and nothing was launched since
    double kernel_ms[PDABM_NUM_KERNELS];
    uint32_t prof_steps;
    bool step_dirty;               // host step number not yet written to Counters::step_no
    bool claim_dirty;              // the claim plane may hold words with tags that are not in the past: zero it
    // CUDA graphs of one whole step (9 launches), one per (at-risk list parity, want_counts); used for worlds
    // small enough that launch latency matters, on explicit (capturable) streams, when profiling is off
    bool graphs;
    cudaGraphExec_t gexec[4];
    uint64_t graph_replays;
};

namespace {

void drop_graphs(pd_sim* s) {  // the captured launches hold plane / list pointers by value
    for (cudaGraphExec_t& g : s->gexec)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

void free_lists(pd_sim* s) {  // cudaFree(nullptr) is a no-op: safe on a partially allocated set
    cudaFree(s->risk[0]); cudaFree(s->risk[1]); cudaFree(s->P.risk_e); cudaFree(s->P.cand);
    cudaFree(s->P.cl_cell); cudaFree(s->P.cl_e); cudaFree(s->P.cl_aux);
    cudaFree(s->P.rt_cell); cudaFree(s->P.rt_state); cudaFree(s->P.rt_aux);
    cudaFree(s->P.nb_cell); cudaFree(s->P.nb_dir); cudaFree(s->P.nb_prio);
    s->risk[0] = s->risk[1] = nullptr; s->P.risk_e = nullptr; s->P.cand = nullptr;
    s->P.cl_cell = nullptr; s->P.cl_e = nullptr; s->P.cl_aux = nullptr;
    s->P.rt_cell = nullptr; s->P.rt_state = s->P.rt_aux = nullptr;
    s->P.nb_cell = nullptr; s->P.nb_dir = nullptr; s->P.nb_prio = nullptr;
    s->lists_allocated = false;
}

int alloc_lists_raw(pd_sim* s, uint64_t cap) {
    CK(cudaMalloc(&s->risk[0], cap * 4));
    CK(cudaMalloc(&s->risk[1], cap * 4));
    CK(cudaMalloc(&s->P.risk_e, cap * 4));
    CK(cudaMalloc(&s->P.cand, (cap / 2 + 1) * 8));
    CK(cudaMalloc(&s->P.cl_cell, cap * 4));
    CK(cudaMalloc(&s->P.cl_e, cap * 4));
    CK(cudaMalloc(&s->P.cl_aux, cap * 2));
    CK(cudaMalloc(&s->P.rt_cell, cap * 4));
    CK(cudaMalloc(&s->P.rt_state, cap));
    CK(cudaMalloc(&s->P.rt_aux, cap));
    CK(cudaMalloc(&s->P.nb_cell, cap * 4));
    CK(cudaMalloc(&s->P.nb_dir, cap));
    CK(cudaMalloc(&s->P.nb_prio, cap * 4));
    return PD_OK;
}

int alloc_lists(pd_sim* s, uint64_t n_agents_hint) {
    if (s->lists_allocated) {
        if (n_agents_hint + 1024 <= s->P.list_cap || s->P.list_cap == s->cells) return PD_OK;
        drop_graphs(s);  // grow: free and reallocate
        free_lists(s);
    }
    // this handle's share of the global cap (+25 % for density fluctuations between slabs; an overflow
    // is detected and reported by pd_read_counts)
    const double share = (double)s->cells / ((double)s->cfg.width * (double)s->P.HW);
    uint64_t local_max = share >= 1.0 ? s->cfg.max_agents : (uint64_t)((double)s->cfg.max_agents * share * 1.25);
    uint64_t cap = local_max > n_agents_hint ? local_max : n_agents_hint;
    cap += 1024;
    if (cap > s->cells) cap = s->cells;
    if (cap > 0xFFFFFFFFull) cap = 0xFFFFFFFFull;  // list indices are 32 bit
    s->P.list_cap = (uint32_t)cap;
    const int rc = alloc_lists_raw(s, cap);
    if (rc) { free_lists(s); return rc; }  // nothing of a partial set is kept
    s->lists_allocated = true;
    return PD_OK;
}

// Persistent sparse kernels: 1 CTA for small worlds; otherwise a cooperative launch that fills every SM
// with as many 512-thread CTAs as the kernel's registers allow (these kernels are latency-bound list
// walks, so more resident warps = more loads in flight).
template <typename K>
int launch_sparse_k(pd_sim* s, K kernel, cudaStream_t st) {
    void* args[] = {(void*)&s->P};
    if (s->sparse_ctas == 1) {
        CK(cudaLaunchKernel((const void*)kernel, dim3(1), dim3(s->sparse_threads), args, 0, st));
    } else {
        int ctas = s->sparse_ctas;
        if (s->sparse_auto) {
            auto it = s->sparse_occ.find((const void*)kernel);
            if (it == s->sparse_occ.end()) {
                int occ = 1;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, s->sparse_threads, 0));
                if (occ < 1) occ = 1;
                if (occ > 4) occ = 4;
                it = s->sparse_occ.emplace((const void*)kernel, occ * s->n_sms).first;
            }
            ctas = it->second;
        }
        CK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(ctas), dim3(s->sparse_threads), args, 0, st));
    }
    return PD_OK;
}

// one 1024-thread CTA, or a cooperative launch of 512-thread CTAs (a second instantiation: 128 registers each)
#define launch_sparse(s, kernel, st) \
    ((s)->sparse_threads == 1024 ? launch_sparse_k(s, kernel<1024>, st) : launch_sparse_k(s, kernel<512>, st))

// per-kernel timing: a (start, end) CUDA event pair around one or more launches, tagged with a kernel slot
int prof_record(pd_sim* s, cudaStream_t st) {
    if (s->ev_used == s->ev.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        s->ev.push_back(e);
    }
    CK(cudaEventRecord(s->ev[s->ev_used], st));
    s->ev_used += 1;
    return PD_OK;
}
int prof_begin(pd_sim* s, int slot, cudaStream_t st) {
    if (!s->profile) return PD_OK;
    if (!s->ev_chain) {
        const int rc = prof_record(s, st);
        if (rc) return rc;
    }
    s->spans.push_back({(uint32_t)(s->ev_used - 1), 0u, slot});
    return PD_OK;
}
int prof_end(pd_sim* s, cudaStream_t st) {
    if (!s->profile) return PD_OK;
    const int rc = prof_record(s, st);
    if (rc) return rc;
    s->spans.back().end = (uint32_t)(s->ev_used - 1);
    s->ev_chain = true;
    return PD_OK;
}
// PDABM_DEBUG_SYNC=1 in the environment: synchronise after every kernel and name the one that failed
static const bool g_debug_sync = getenv("PDABM_DEBUG_SYNC") != nullptr;
#define PROF(slot, ...)                                                                              \
    do {                                                                                             \
        if ((rc = prof_begin(s, slot, st))) return rc;                                               \
        __VA_ARGS__;                                                                                 \
        if ((rc = prof_end(s, st))) return rc;                                                       \
        if (g_debug_sync) {                                                                          \
            cudaError_t e_ = cudaStreamSynchronize(st);                                              \
            if (e_ == cudaSuccess) e_ = cudaGetLastError();                                          \
            if (e_ != cudaSuccess)                                                                   \
                return fail(PD_ERR_CUDA, "kernel slot %d of step %u failed: %s", slot, s->step,       \
                            cudaGetErrorString(e_));                                                 \
        }                                                                                            \
    } while (0)

// Claim words carry (step mod 2^25) in their tag and are never cleared, because a later round always has a higher
// tag.  That stops being true when the step is set from outside (pd_set_step, pd_init_population) or its 25 bits
// wrap: the step then starts from a clean plane.
int begin_claims(pd_sim* s, cudaStream_t st) {
    if (s->claim_dirty || (s->step & 0x1FFFFFFu) == 0u) {
        CK(cudaMemsetAsync(s->P.claim, 0, s->cells * sizeof(unsigned long long), st));
        s->claim_dirty = false;
    }
    return PD_OK;
}

int push_step(pd_sim* s, cudaStream_t st) {
    if (!s->step_dirty) return PD_OK;
    k_set_step<<<1, 1, 0, st>>>(s->P.ctr, s->step);
    CK(cudaGetLastError());
    s->step_dirty = false;
    return PD_OK;
}

// the sparse round-1 commits: independent threads over a device-side list length, grid-stride
int launch_commit(pd_sim* s, void (*kernel)(const Params), cudaStream_t st) {
    const size_t want = (s->cells / 16 + COMMIT_CHUNK - 1) / COMMIT_CHUNK;  // ~5 % of the cells claim
#ifndef PD_COMMIT_CAP
#define PD_COMMIT_CAP 16  // CTAs per SM of the commit kernels' grid (measured: 8 -> 16 saves 0.02 ms)
#endif
    const size_t cap = (size_t)s->n_sms * PD_COMMIT_CAP;
    kernel<<<(unsigned int)(want < cap ? want : cap), COMMIT_NT, 0, st>>>(s->P);
    CK(cudaGetLastError());
    return PD_OK;
}

// one dense sweep over the tiles
#define DENSE(kernel) do { if (s->big_tiles) kernel<128, 16><<<s->grid, NT, 0, st>>>(P); \
                           else kernel<16, 16><<<s->grid, NT, 0, st>>>(P); } while (0)

// the 9 launches of one model step, in the order of the reference's main layers (model.py:2140-2163)
int launch_step(pd_sim* s, cudaStream_t st) {
    Params& P = s->P;
    int rc = PD_OK;
    s->ev_chain = false;  // (whatever the caller launched since the last span is not the first kernel's time)
    PROF(0, (k_begin_step<<<1, 256, 0, st>>>(P.ctr, 1)));
    PROF(1, { if ((rc = launch_sparse(s, k_risk_resolve, st))) return rc; });
    PROF(2, DENSE(k_play_travel));
    PROF(3, { if ((rc = launch_commit(s, k_move_commit, st))) return rc; });
    PROF(4, { if ((rc = launch_sparse(s, k_move_retry, st))) return rc; });
    PROF(5, DENSE(k_claim_punish));
    PROF(6, { if ((rc = launch_commit(s, k_repro_commit, st))) return rc; });
    PROF(7, { if ((rc = launch_sparse(s, k_repro_retry, st))) return rc; });
    PROF(8, { if ((rc = launch_sparse(s, k_finalize, st))) return rc; });
    return PD_OK;
}

// Replay (capturing on first use) the graph of one step for the current Params.  Returns PD_OK with
// *done = false when graphs cannot be used on this stream -- the caller then launches directly.
int graph_step(pd_sim* s, cudaStream_t st, bool* done) {
    *done = false;
    if (!s->graphs || s->profile || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return PD_OK;
    const int idx = s->risk_cur * 2 + (s->P.want_counts ? 1 : 0);
    if (!s->gexec[idx]) {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        int rc = PD_OK;
        if (e == cudaSuccess) {
            rc = launch_step(s, st);
            e = cudaStreamEndCapture(st, &g);
        }
        if (e == cudaSuccess && rc == PD_OK) e = cudaGraphInstantiate(&s->gexec[idx], g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess || rc != PD_OK) {  // not capturable here: fall back to plain launches for good
            cudaGetLastError();
            s->gexec[idx] = nullptr;
            s->graphs = false;
            return PD_OK;
        }
    }
    CK(cudaGraphLaunch(s->gexec[idx], st));
    s->graph_replays += 1;
    *done = true;
    return PD_OK;
}

int run_scan(pd_sim* s, cudaStream_t st) {
    s->P.risk_cur = s->risk[s->risk_cur];
    s->P.risk_next = s->risk[s->risk_cur ^ 1];
    k_begin_step<<<1, 256, 0, st>>>(s->P.ctr, 0);
    // the overflow flag is sticky within a run only: a (re)initialised state starts clean (the first scan of
    // pd_sync_state may overflow the initial lists before they are grown and the scan is repeated)
    CK(cudaMemsetAsync(&s->P.ctr->overflow, 0, sizeof(unsigned int), st));
    const int blocks = (int)((s->cells + NT - 1) / NT < (size_t)(s->n_sms * 16) ? (s->cells + NT - 1) / NT
                                                                                  : (size_t)(s->n_sms * 16));
    k_scan_state<<<blocks, NT, 0, st>>>(s->P, s->cells);
    k_scan_finish<<<1, 32, 0, st>>>(s->P.ctr, s->P.list_cap);
    CK(cudaGetLastError());
    s->risk_cur ^= 1;
    return PD_OK;
}

}  // namespace

extern "C" {

const char* pd_last_error(void) { return g_err.c_str(); }
int pd_abi_version(void) { return PDABM_ABI_VERSION; }

int pd_default_config(pd_config* c, uint32_t width) {
    if (!c) return fail(PD_ERR_INVALID, "cfg is NULL");
    memset(c, 0, sizeof *c);
    c->struct_size = (uint32_t)sizeof *c;
    c->width = width;
    c->row0 = 0;
    c->rows = width;
    c->seed = 12345;
    const uint64_t cells = (uint64_t)width * width;
    c->max_agents = (uint64_t)((double)cells * 0.5);  // model.py:49
    c->payoff_cc = 3.0f; c->payoff_dc = 5.0f; c->payoff_cd = -1.0f; c->payoff_dd = 0.0f;
    c->env_noise = 0.0f;
    c->cost_of_living = 1.0f;
    c->travel_cost = 0.5f;
    c->reproduce_min_energy = 100.0f;
    c->reproduce_cost = 50.0f;
    c->reproduction_inheritence = 0.0f;
    c->init_energy_mu = 50.0f; c->init_energy_sigma = 10.0f; c->init_energy_min = 5.0f;
    c->max_energy = 150.0f;
    c->mutation_rate = 0.0f;
    for (int i = 0; i < 4; ++i) c->strategy_weights[i] = 0.25f;
    c->strategy_pure = 0; c->strategy_per_trait = 0; c->max_children_per_step = 1;
    c->collect_stats = 0;
    c->device = 0;
    c->sparse_ctas = 0;
    c->use_graphs = 0;
    return PD_OK;
}

int pd_create(const pd_config* cfg, pd_sim** out) {
    if (!cfg || !out) return fail(PD_ERR_INVALID, "NULL argument");
    if (cfg->struct_size != sizeof(pd_config))
        return fail(PD_ERR_INVALID, "pd_config.struct_size %u != %zu (ABI mismatch)", cfg->struct_size, sizeof(pd_config));
    const uint32_t W = cfg->width;
    if (W < 16 || (W & (W - 1))) return fail(PD_ERR_INVALID, "width must be a power of two >= 16 (got %u)", W);
    if (W > 65536) return fail(PD_ERR_INVALID, "width > 65536 not supported");
    const uint32_t HW = cfg->world_height ? cfg->world_height : W;  // non-square worlds: weak-scaling series
    if (HW < 16 || (HW & (HW - 1)) || HW > (1u << 20)) return fail(PD_ERR_INVALID, "world_height must be a power of two in [16, 2^20]");
    const bool slab = cfg->rows != HW;
    if (!slab && (cfg->row0 != 0 || cfg->ghost_rows != 0))
        return fail(PD_ERR_INVALID, "a whole world (rows == world height) needs row0 = 0 and ghost_rows = 0");
    if (slab) {
        if (cfg->rows == 0 || cfg->rows % 16 || cfg->ghost_rows % 16 || cfg->ghost_rows < 64)
            return fail(PD_ERR_INVALID, "slab: rows and ghost_rows must be multiples of 16 and ghost_rows >= 64 (got %u, %u)",
                        cfg->rows, cfg->ghost_rows);
        if (cfg->ghost_rows > cfg->rows || cfg->row0 + cfg->rows > HW || HW % cfg->rows)
            return fail(PD_ERR_INVALID, "slab: need ghost_rows <= rows, rows | world height and row0 + rows <= world height");
    }
    // cell indices in the lists, the claim entries and the low word of the cull / placement keys are 32 bit
    if ((uint64_t)W * HW > (1ull << 32))
        return fail(PD_ERR_INVALID, "world of %u x %u cells exceeds 2^32 cells", W, HW);
    if ((uint64_t)W * ((uint64_t)cfg->rows + 2ull * cfg->ghost_rows) > 0xFFFFFFFFull && slab)
        return fail(PD_ERR_INVALID, "slab of %u x %u local cells exceeds 2^32 - 1", W, cfg->rows + 2 * cfg->ghost_rows);
    double wsum = 0;
    for (int i = 0; i < 4; ++i) {
        if (!(cfg->strategy_weights[i] >= 0.0f)) return fail(PD_ERR_INVALID, "strategy_weights must be >= 0");
        wsum += (double)cfg->strategy_weights[i];
    }
    if (!(wsum > 0)) return fail(PD_ERR_INVALID, "strategy_weights sum to 0");
    if (cfg->debug_tie_shift > 31u) return fail(PD_ERR_INVALID, "debug_tie_shift must be in [0, 31]");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PD_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));
    pd_sim* s = new pd_sim();  // value-initialised: every pointer below starts as nullptr
    struct Guard {             // a failure part-way releases what was allocated so far
        pd_sim* s;
        ~Guard() { if (s) pd_destroy(s); }
    } guard{s};
    s->cfg = *cfg;
    const uint32_t Hl = cfg->rows + 2 * cfg->ghost_rows;  // local rows incl. the ghost zones
    s->cells = (size_t)Hl * W;
    s->slab = slab;
    Params& P = s->P;
    P.W = W; P.H = Hl; P.HW = HW;
    P.row0 = (cfg->row0 + HW - cfg->ghost_rows) & (HW - 1);  // global row of local row 0
    P.wrap_y = slab ? 0u : 1u;
    P.own_y0 = cfg->ghost_rows; P.own_y1 = cfg->ghost_rows + cfg->rows;
    P.round0 = 1; P.round1 = 8; P.world = 1; P.cand_cap = 0; P.speculative = 0;
    P.select_cap = cfg->debug_select_cap ? cfg->debug_select_cap : 0xFFFFFFFFu;
    P.tie_shift = cfg->debug_tie_shift ? cfg->debug_tie_shift : 3u;
    P.step_log = nullptr; P.log_cap = 0;
    P.glb = nullptr; P.ghist = nullptr; P.cand_out = nullptr; P.cand_all = nullptr;
    P.log2W = 0;
    while ((1u << P.log2W) < W) ++P.log2W;
    P.k0 = (uint32_t)cfg->seed; P.k1 = (uint32_t)(cfg->seed >> 32);
    P.pay[0] = cfg->payoff_dd; P.pay[1] = cfg->payoff_dc; P.pay[2] = cfg->payoff_cd; P.pay[3] = cfg->payoff_cc;
    P.max_energy = cfg->max_energy; P.travel_cost = cfg->travel_cost; P.cost_of_living = cfg->cost_of_living;
    P.reproduce_min_energy = cfg->reproduce_min_energy; P.reproduce_cost = cfg->reproduce_cost;
    P.inherit = cfg->reproduction_inheritence;
    P.e_mu = cfg->init_energy_mu; P.e_sigma = cfg->init_energy_sigma; P.e_min = cfg->init_energy_min;
    P.mutation_rate = cfg->mutation_rate;
    P.noise_thr = prob_threshold(cfg->env_noise);
    P.mut_thr = prob_threshold(cfg->mutation_rate);
    double cum = 0;
    for (int i = 0; i < 3; ++i) {
        cum += (double)cfg->strategy_weights[i];
        const double f = floor(cum / wsum * 4294967296.0);
        P.strat_thr[i] = f >= 4294967296.0 ? 4294967296ull : (unsigned long long)f;
    }
    float m = 0.0f;
    for (int i = 0; i < 4; ++i) m = fminf(m, P.pay[i]);
    // an agent above this energy cannot die during the (at most 8) games of one step
    P.risk_thr = m < 0.0f ? (float)(8.0 * fabs((double)m) * 1.0001 + 1e-3) : -1.0f;
    P.max_agents = cfg->max_agents;
    P.pure = cfg->strategy_pure ? 1u : 0u;
    P.per_trait = cfg->strategy_per_trait ? 1u : 0u;
    P.max_children = cfg->max_children_per_step;
    P.stats = cfg->collect_stats ? 1u : 0u;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    s->n_sms = prop.multiProcessorCount;
    s->big_tiles = W >= 128;
    s->grid = s->big_tiles ? dim3(W / 128, Hl / 16) : dim3(W / 16, Hl / 16);
    // measured on B200 (profiles/README.md): one 1024-thread CTA wins up to 2^19 cells (no grid barriers),
    // 16 CTAs up to 2^22 (0.283 -> 0.235 ms/step at 1024^2 and still co-schedulable with other runs'
    // kernels in an ensemble), every SM beyond that
    s->sparse_auto = cfg->sparse_ctas <= 0 && s->cells > ((size_t)1 << 22);
    if (cfg->sparse_ctas > 0) s->sparse_ctas = cfg->sparse_ctas;
    else s->sparse_ctas = s->cells <= ((size_t)1 << 19) ? 1 : s->cells <= ((size_t)1 << 22) ? 16 : s->n_sms;
    if (s->sparse_ctas > 1 && !prop.cooperativeLaunch) s->sparse_ctas = 1;
    s->sparse_threads = s->sparse_ctas == 1 ? 1024 : 512;
    if (s->sparse_ctas > s->n_sms) s->sparse_ctas = s->n_sms;
    P.sparse_ctas = (uint32_t)s->sparse_ctas;
    s->graphs = cfg->use_graphs > 0 || (cfg->use_graphs == 0 && !slab && s->cells <= ((size_t)1 << 22));
    if (slab) s->graphs = false;  // a slab is stepped in parts with host decisions in between
    CK(cudaMalloc(&P.claim, s->cells * sizeof(unsigned long long)));
    s->claim_dirty = true;  // zeroed on the stream of the first step
    {
        float h_lut[256], *d_lut = nullptr;
        for (uint32_t i = 0; i < 256; ++i) h_lut[i] = lut_entry(P.pay, i);
        CK(cudaMalloc(&d_lut, sizeof h_lut));
        P.lut = d_lut;
        CK(cudaMemcpy(d_lut, h_lut, sizeof h_lut, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&P.E1, s->cells * sizeof(float)));
    CK(cudaMalloc(&P.tfB, s->cells));
    CK(cudaMalloc(&P.ctr, sizeof(Counters)));
    CK(cudaMemset(P.ctr, 0, sizeof(Counters)));
    CK(cudaMallocHost(&s->h, sizeof(HostOut)));
    memset(s->h, 0, sizeof(HostOut));
    guard.s = nullptr;
    *out = s;
    return PD_OK;
}

int pd_destroy(pd_sim* s) {
    if (!s) return PD_OK;
    cudaSetDevice(s->cfg.device);
    // (no cudaDeviceSynchronize: it is an error while another thread captures a graph; cudaFree below waits)
    drop_graphs(s);
    if (s->P.step_log) cudaFree(s->P.step_log);
    cudaFree(s->P.E1); cudaFree(s->P.tfB); cudaFree(s->P.ctr); cudaFree(const_cast<float*>(s->P.lut));
    cudaFree(s->P.claim);
    free_lists(s);
    if (s->h) cudaFreeHost(s->h);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    delete s;
    return PD_OK;
}

int pd_bind_planes(pd_sim* s, float* energy, uint8_t* strat, uint8_t* tf) {
    if (!s || !energy || !strat || !tf) return fail(PD_ERR_INVALID, "NULL argument");
    cudaPointerAttributes a;
    const void* ptrs[3] = {energy, strat, tf};
    for (int i = 0; i < 3; ++i) {
        CK(cudaPointerGetAttributes(&a, ptrs[i]));
        if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
            return fail(PD_ERR_INVALID, "plane %d is not device memory", i);
    }
    for (int i = 0; i < 3; ++i)
        if (((uintptr_t)ptrs[i]) & 15u) return fail(PD_ERR_INVALID, "plane %d is not 16-byte aligned", i);
    s->P.E0 = energy; s->P.strat = strat; s->P.tfA = tf;
    drop_graphs(s);
    s->bound = true;
    s->ready = false;
    return PD_OK;
}

int pd_sync_state(pd_sim* s, void* stream) {
    if (!s || !s->bound) return fail(PD_ERR_STATE, "planes not bound");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    if (!s->lists_allocated) {
        int rc = alloc_lists(s, 0);
        if (rc) return rc;
    }
    int rc = run_scan(s, st);
    if (rc) return rc;
    // make sure the lists are large enough for the population actually found
    CK(cudaMemcpyAsync(&s->h->n_agents, &s->P.ctr->n_agents, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (s->h->n_agents + 1024 > s->P.list_cap && s->P.list_cap < s->cells) {
        rc = alloc_lists(s, s->h->n_agents);
        if (rc) return rc;
        rc = run_scan(s, st);
        if (rc) return rc;
    }
    s->ready = true;
    return PD_OK;
}

int pd_init_population(pd_sim* s, uint64_t n_agents, void* stream) {
    if (!s || !s->bound) return fail(PD_ERR_STATE, "planes not bound");
    const size_t world_cells = (size_t)s->cfg.width * s->P.HW;
    if (n_agents > world_cells) return fail(PD_ERR_INVALID, "n_agents %llu > cells %zu", (unsigned long long)n_agents, world_cells);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    int rc = alloc_lists(s, n_agents);
    if (rc) return rc;
    const int blocks = (int)((s->cells + NT - 1) / NT < (size_t)(s->n_sms * 16) ? (s->cells + NT - 1) / NT
                                                                                  : (size_t)(s->n_sms * 16));
    unsigned long long tau = 0;
    const int none = n_agents == 0;
    const int gblocks = (int)((world_cells + NT - 1) / NT < (size_t)(s->n_sms * 16) ? (world_cells + NT - 1) / NT
                                                                                      : (size_t)(s->n_sms * 16));
    if (n_agents > 0 && n_agents < world_cells) {
        // n-th SMALLEST (priority, cell) key by an 8-pass radix select (model.py:1438-1450 restated)
        k_begin_step<<<1, 256, 0, st>>>(s->P.ctr, 0);
        unsigned long long prefix = 0, krem = n_agents;
        for (int pass = 0; pass < 8; ++pass) {
            k_init_hist<<<gblocks, NT, 0, st>>>(s->P, world_cells, pass, prefix);
            CK(cudaMemcpyAsync(s->h->hist, s->P.ctr->hist[pass], sizeof s->h->hist, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            unsigned long long cum = 0;
            int digit = 255;
            for (int b = 0; b < 256; ++b) {
                if (cum + s->h->hist[b] >= krem) { digit = b; break; }
                cum += s->h->hist[b];
            }
            krem -= cum;
            prefix = (prefix << 8) | (unsigned long long)digit;
        }
        tau = prefix;
    } else if (n_agents == world_cells) {
        tau = ~0ull;
    }
    k_init_write<<<blocks, NT, 0, st>>>(s->P, s->cells, tau, none);
    CK(cudaGetLastError());
    s->step = 0;
    s->step_dirty = true;
    s->claim_dirty = true;
    rc = run_scan(s, st);
    if (rc) return rc;
    s->ready = true;
    return PD_OK;
}

int pd_step(pd_sim* s, uint32_t n_steps, void* stream, int want_counts) {
    if (!s || !s->ready) return fail(PD_ERR_STATE, "pd_step before pd_init_population / pd_sync_state");
    if (s->slab) return fail(PD_ERR_STATE, "a row slab is stepped with pd_step_part (the caller exchanges ghost rows and sums in between)");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    int rc = push_step(s, st);
    if (rc) return rc;
    for (uint32_t k = 0; k < n_steps; ++k) {
        Params& P = s->P;
        P.want_counts = (want_counts == 2 || (want_counts && k + 1 == n_steps)) ? 1u : 0u;
        P.risk_cur = s->risk[s->risk_cur];
        P.risk_next = s->risk[s->risk_cur ^ 1];
        if ((rc = begin_claims(s, st))) return rc;
        bool replayed = false;
        if ((rc = graph_step(s, st, &replayed))) return rc;
        if (!replayed && (rc = launch_step(s, st))) return rc;
        s->risk_cur ^= 1;
        s->step += 1;
    }
    CK(cudaGetLastError());
    return PD_OK;
}

static int read_back(pd_sim* s, cudaStream_t st) {
    CK(cudaSetDevice(s->cfg.device));
    Counters* c = s->P.ctr;
    CK(cudaMemcpyAsync(s->h->bins, c->last_bins, sizeof s->h->bins, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s->h->stats, c->last_stats, sizeof s->h->stats, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&s->h->n_agents, &c->n_agents, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&s->h->overflow, &c->overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (s->h->overflow)
        return fail(PD_ERR_OVERFLOW, "a sparse list overflowed its capacity (%u entries); results are invalid", s->P.list_cap);
    return PD_OK;
}

int pd_read_counts(pd_sim* s, uint64_t counts[16], uint64_t* n_agents, void* stream) {
    if (!s || !s->ready) return fail(PD_ERR_STATE, "no state");
    int rc = read_back(s, (cudaStream_t)stream);
    if (rc) return rc;
    if (counts) for (int k = 0; k < 16; ++k) counts[k] = s->h->bins[k];
    if (n_agents) *n_agents = s->h->n_agents;
    return PD_OK;
}

int pd_read_stats(pd_sim* s, pd_step_stats* o, void* stream) {
    if (!s || !s->ready || !o) return fail(PD_ERR_STATE, "no state");
    int rc = read_back(s, (cudaStream_t)stream);
    if (rc) return rc;
    const unsigned long long* v = s->h->stats;
    o->agents_start = v[0]; o->agents_end = v[1]; o->play_deaths = v[2]; o->movers = v[3];
    o->moved = v[4]; o->travel_deaths = v[5]; o->births = v[6]; o->culled = v[7];
    o->living_deaths = v[8]; o->risk_list = v[9]; o->move_losers = v[10]; o->repro_losers = v[11];
    return PD_OK;
}

int pd_bind_comm(pd_sim* s, uint64_t* glb, uint32_t* ghist, uint64_t* cand_out, const uint64_t* cand_all,
                 uint32_t cand_cap, uint32_t world) {
    if (!s || !glb || !ghist || !cand_out || !cand_all || !cand_cap || !world) return fail(PD_ERR_INVALID, "NULL argument");
    if (!s->slab) return fail(PD_ERR_STATE, "pd_bind_comm is for row slabs (rows != width)");
    if (world > PD_MAX_WORLD) return fail(PD_ERR_INVALID, "world %u exceeds PD_MAX_WORLD %d", world, PD_MAX_WORLD);
    s->P.glb = (unsigned long long*)glb;
    s->P.ghist = ghist;
    s->P.cand_out = (unsigned long long*)cand_out;
    s->P.cand_all = (const unsigned long long*)cand_all;
    s->P.cand_cap = cand_cap;
    s->P.world = world;
    s->comm_bound = true;
    return PD_OK;
}

int pd_step_part(pd_sim* s, int part, int arg, void* stream, int want_counts) {
    if (!s || !s->ready) return fail(PD_ERR_STATE, "pd_step_part before pd_init_population / pd_sync_state");
    if (!s->slab || !s->comm_bound) return fail(PD_ERR_STATE, "pd_step_part needs a row slab with pd_bind_comm done");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    Params& P = s->P;
    int rc = push_step(s, st);
    if (rc) return rc;
    P.want_counts = want_counts ? 1u : 0u;
    P.risk_cur = s->risk[s->risk_cur];
    P.risk_next = s->risk[s->risk_cur ^ 1];
    s->ev_chain = false;  // (the caller's exchange / collectives ran since the last span)
    switch (part) {
    case PD_PART_PRE: {
        if ((rc = begin_claims(s, st))) return rc;
        // ghost rows were just refreshed by the caller: their at-risk agents join the list
        const size_t ghost_cells = (size_t)2 * s->cfg.ghost_rows * P.W;
        const int blocks = (int)((ghost_cells + 255) / 256 < (size_t)(s->n_sms * 8) ? (ghost_cells + 255) / 256 : (size_t)(s->n_sms * 8));
        PROF(0, { k_ghost_risk_scan<<<blocks, 256, 0, st>>>(P); k_begin_step<<<1, 256, 0, st>>>(P.ctr, 1); });
        PROF(1, { if ((rc = launch_sparse(s, k_risk_resolve, st))) return rc; });
        PROF(2, DENSE(k_play_travel));
        PROF(3, { if ((rc = launch_commit(s, k_move_commit, st))) return rc; });
        PROF(4, { if ((rc = launch_sparse(s, k_move_retry, st))) return rc; });
        PROF(5, DENSE(k_claim_punish));
        PROF(6, { if ((rc = launch_commit(s, k_repro_commit, st))) return rc; });
        k_publish_gate<<<1, 32, 0, st>>>(P, 0);
        // the cull histogram of round 1's newborns rides on the same all-reduce as the gate words
        PROF(8, (k_fin_hist<<<1, 256, 0, st>>>(P)));
        break;
    }
    case PD_PART_RETRY:
        if (arg < 1 || arg > 7) return fail(PD_ERR_INVALID, "retry round %d out of range 1..7", arg);
        P.round0 = (uint32_t)arg; P.round1 = (uint32_t)arg + 1;
        PROF(7, { if ((rc = launch_sparse(s, k_repro_retry, st))) return rc; });
        k_publish_gate<<<1, 32, 0, st>>>(P, arg);
        P.round0 = 1; P.round1 = 8;
        break;
    case PD_PART_HIST:  // retry rounds added newborns: this slab's cull histogram again
        PROF(8, (k_fin_hist<<<1, 256, 0, st>>>(P)));
        break;
    case PD_PART_CAND:  // the reproduction rounds are over: this slab's cull candidates
        P.speculative = arg ? 1u : 0u;
        PROF(8, { if ((rc = launch_sparse(s, k_fin_cand, st))) return rc; });
        P.speculative = 0;
        break;
    case PD_PART_APPLY:
        P.speculative = arg ? 1u : 0u;
        PROF(8, { if ((rc = launch_sparse(s, k_fin_apply, st))) return rc; });
        P.speculative = 0;
        if (arg) break;  // speculative: the caller confirms with PD_PART_DONE once it has seen the gate
        s->risk_cur ^= 1;
        s->step += 1;
        break;
    case PD_PART_DONE:  // the speculative CAND / APPLY stood: the step is complete (no launch)
        s->risk_cur ^= 1;
        s->step += 1;
        break;
    default:
        return fail(PD_ERR_INVALID, "unknown part %d", part);
    }
    CK(cudaGetLastError());
    return PD_OK;
}

int pd_halo_bytes(pd_sim* s, uint64_t* strip_bytes) {
    if (!s || !strip_bytes) return fail(PD_ERR_INVALID, "NULL argument");
    if (!s->slab) return fail(PD_ERR_STATE, "halo strips exist for row slabs only");
    *strip_bytes = (uint64_t)s->cfg.ghost_rows * s->cfg.width * 6u;
    return PD_OK;
}

static int halo_copy(pd_sim* s, void* a, void* b, uint32_t a_row, uint32_t b_row, int unpack, void* stream) {
    if (!s || !a || !b) return fail(PD_ERR_INVALID, "NULL argument");
    if (!s->slab || !s->bound) return fail(PD_ERR_STATE, "halo strips exist for bound row slabs only");
    if (((uintptr_t)a | (uintptr_t)b) & 15u) return fail(PD_ERR_INVALID, "strips must be 16-byte aligned");
    CK(cudaSetDevice(s->cfg.device));
    const size_t chunks = (size_t)s->cfg.ghost_rows * s->cfg.width * 6u / 16u * 2u;
    const size_t want = (chunks + 255) / 256, cap = (size_t)s->n_sms * 8;
    k_halo_copy<<<(unsigned int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        s->P, (uint4*)a, (uint4*)b, a_row, b_row, s->cfg.ghost_rows, unpack);
    CK(cudaGetLastError());
    return PD_OK;
}
int pd_halo_pack(pd_sim* s, void* first_rows, void* last_rows, void* stream) {
    if (!s) return fail(PD_ERR_INVALID, "NULL argument");
    return halo_copy(s, first_rows, last_rows, s->P.own_y0, s->P.own_y1 - s->cfg.ghost_rows, 0, stream);
}
int pd_halo_unpack(pd_sim* s, const void* top_ghost, const void* bottom_ghost, void* stream) {
    if (!s) return fail(PD_ERR_INVALID, "NULL argument");
    return halo_copy(s, const_cast<void*>(top_ghost), const_cast<void*>(bottom_ghost), 0, s->P.own_y1, 1, stream);
}

int pd_set_profile(pd_sim* s, int on) {
    if (!s) return fail(PD_ERR_INVALID, "NULL argument");
    s->profile = on != 0;
    s->ev_used = 0;
    s->spans.clear();
    s->ev_chain = false;
    s->prof_steps = 0;
    for (int i = 0; i < PDABM_NUM_KERNELS; ++i) s->kernel_ms[i] = 0.0;
    return PD_OK;
}

int pd_read_profile(pd_sim* s, double kernel_ms[PDABM_NUM_KERNELS], uint32_t* steps, void* stream) {
    if (!s || !kernel_ms) return fail(PD_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    for (const pd_sim::ProfSpan& sp : s->spans) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ev[sp.start], s->ev[sp.end]));
        s->kernel_ms[sp.slot] += (double)ms;
        if (sp.slot == 2) s->prof_steps += 1;  // one play sweep per step
    }
    s->spans.clear();
    s->ev_used = 0;
    s->ev_chain = false;
    for (int i = 0; i < PDABM_NUM_KERNELS; ++i) kernel_ms[i] = s->kernel_ms[i];
    if (steps) *steps = s->prof_steps;
    return PD_OK;
}

int pd_read_agent_steps(pd_sim* s, uint64_t* agent_steps, void* stream) {
    if (!s || !agent_steps) return fail(PD_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(s->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(&s->h->n_agents, &s->P.ctr->agent_steps, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *agent_steps = s->h->n_agents;
    return PD_OK;
}

int pd_set_step_log(pd_sim* s, uint32_t capacity_rows, void* stream) {
    if (!s) return fail(PD_ERR_INVALID, "NULL argument");
    if (s->slab) return fail(PD_ERR_STATE, "the step log holds whole-world counts; a slab's caller sums them itself");
    CK(cudaSetDevice(s->cfg.device));
    // No device-wide synchronisation here: another handle's thread may be capturing a graph (an ensemble),
    // during which cudaDeviceSynchronize is an error.  cudaFree waits for the device by itself; the caller
    // must not have steps of THIS handle in flight (enable the log before stepping, resize after a read).
    drop_graphs(s);  // the captured launches hold the log pointer by value
    if (s->P.step_log) { cudaFree(s->P.step_log); s->P.step_log = nullptr; }
    s->P.log_cap = 0;
    if (capacity_rows) {
        CK(cudaMalloc(&s->P.step_log, (size_t)capacity_rows * PD_LOG_ROW_WORDS * sizeof(unsigned long long)));
        s->P.log_cap = capacity_rows;
    }
    CK(cudaMemsetAsync(&s->P.ctr->log_n, 0, 2 * sizeof(unsigned int), (cudaStream_t)stream));
    return PD_OK;
}

int pd_read_step_log(pd_sim* s, uint64_t* rows, uint32_t max_rows, uint32_t* n_rows, void* stream) {
    if (!s || !rows || !n_rows) return fail(PD_ERR_INVALID, "NULL argument");
    if (!s->P.step_log) return fail(PD_ERR_STATE, "step log not enabled (pd_set_step_log)");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    unsigned int head[2] = {0, 0};  // log_n, log_dropped
    CK(cudaMemcpyAsync(head, &s->P.ctr->log_n, sizeof head, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n = head[0] < max_rows ? head[0] : max_rows;
    if (n) CK(cudaMemcpyAsync(rows, s->P.step_log, (size_t)n * PD_LOG_ROW_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(&s->P.ctr->log_n, 0, 2 * sizeof(unsigned int), st));
    CK(cudaStreamSynchronize(st));
    *n_rows = n;
    if (head[1] || head[0] > max_rows)
        return fail(PD_ERR_OVERFLOW, "step log: %u rows did not fit (capacity %u, read %u of %u)", head[1], s->P.log_cap, n, head[0]);
    return PD_OK;
}

int pd_graph_replays(pd_sim* s, uint64_t* steps) {
    if (!s || !steps) return fail(PD_ERR_INVALID, "NULL argument");
    *steps = s->graph_replays;
    return PD_OK;
}

int pd_get_step(pd_sim* s, uint32_t* step) {
    if (!s || !step) return fail(PD_ERR_INVALID, "NULL argument");
    *step = s->step;
    return PD_OK;
}
int pd_set_step(pd_sim* s, uint32_t step) {
    if (!s) return fail(PD_ERR_INVALID, "NULL argument");
    s->step = step;
    s->step_dirty = true;  // written to the device at the start of the next step call, on its stream
    s->claim_dirty = true;
    return PD_OK;
}

int pd_step_host(pd_sim* s, const float* h_e, const uint8_t* h_s, const uint8_t* h_t, uint32_t n_steps,
                 float* o_e, uint8_t* o_s, uint8_t* o_t, uint64_t counts[16], uint64_t* n_agents, void* stream) {
    if (!s || !s->bound) return fail(PD_ERR_STATE, "planes not bound");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(s->cfg.device));
    if (h_e && h_s && h_t) {
        CK(cudaMemcpyAsync(s->P.E0, h_e, s->cells * sizeof(float), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->P.strat, h_s, s->cells, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->P.tfA, h_t, s->cells, cudaMemcpyHostToDevice, st));
        int rc = pd_sync_state(s, stream);
        if (rc) return rc;
    } else if (h_e || h_s || h_t) {
        return fail(PD_ERR_INVALID, "pass all three input planes or none");
    }
    int rc = pd_step(s, n_steps, stream, counts != NULL);
    if (rc) return rc;
    if (o_e) CK(cudaMemcpyAsync(o_e, s->P.E0, s->cells * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (o_s) CK(cudaMemcpyAsync(o_s, s->P.strat, s->cells, cudaMemcpyDeviceToHost, st));
    if (o_t) CK(cudaMemcpyAsync(o_t, s->P.tfA, s->cells, cudaMemcpyDeviceToHost, st));
    return pd_read_counts(s, counts, n_agents, stream);
}

// test hook: Philox known-answer vectors through the device implementation
int pd_debug_philox(const uint32_t* ctr_key6, uint32_t* out4, int n) {
    uint32_t *d_in = nullptr, *d_out = nullptr;
    struct Guard {
        uint32_t *&a, *&b;
        ~Guard() { cudaFree(a); cudaFree(b); }
    } guard{d_in, d_out};
    CK(cudaMalloc(&d_in, (size_t)n * 6 * 4));
    CK(cudaMalloc(&d_out, (size_t)n * 4 * 4));
    CK(cudaMemcpy(d_in, ctr_key6, (size_t)n * 6 * 4, cudaMemcpyHostToDevice));
    k_philox_kat<<<(n + 127) / 128, 128>>>(d_in, d_out, n);
    CK(cudaMemcpy(out4, d_out, (size_t)n * 4 * 4, cudaMemcpyDeviceToHost));
    return PD_OK;
}

}  // extern "C"

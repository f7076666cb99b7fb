/*
 * pdabm.h -- C ABI of libpdabm.so: the B200-native per-step agent pipeline of the
 * Prisoner's-Dilemma ABM (play -> travel -> reproduce -> cost of living/death ->
 * per-strategy counts) on dense structure-of-arrays grid planes.
 *
 * This is the drop-in boundary for the path the reference runs through pyflamegpu
 * (all citations are /root/reference/src/model.py):
 *   - the 12 RTC agent functions + 9 conditions        :352-1371
 *   - the host conditions that drive the sub-loops     :1527-1627
 *   - step_fn's 16 strategy counts                     :1405-1419
 *   - init_fn's population set-up                      :1423-1524
 *   - CUDASimulation::simulate()'s step loop           :2176
 * A reference-side maintainer replaces `simulation.simulate()` by a Python loop over
 * pd_step()/pd_read_counts() (INTEGRATION.md shows the ctypes stub).
 *
 * Conventions: every function returns 0 on success and a negative pd_status on error;
 * pd_last_error() returns a thread-local message.  No C++ exception crosses the ABI.
 * `stream` arguments are cudaStream_t values passed as void* (NULL = default stream).
 * A pd_sim handle is single-threaded; distinct handles are independent (one per CUDA
 * stream for ensembles, one per device for a row-decomposed world).
 */
#ifndef PDABM_H_
#define PDABM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDABM_ABI_VERSION 1

typedef enum pd_status {
    PD_OK = 0,
    PD_ERR_INVALID = -1,   /* bad argument / unsupported configuration            */
    PD_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has the detail)  */
    PD_ERR_STATE = -3,     /* call order violated (e.g. step before bind/init)     */
    PD_ERR_OVERFLOW = -4   /* an internal sparse list overflowed its capacity      */
} pd_status;

/* Bits of the `tf` plane as seen by the caller BETWEEN steps (canonical form).
 * Other bits are used transiently inside a step and are always 0 between steps. */
#define PD_TF_TRAIT_MASK 0x03u  /* agent_trait, model.py:1638                      */
#define PD_TF_OCCUPIED   0x80u  /* a prisoner stands on this cell                  */

/* Mirrors the configuration block (model.py:39-169) and the environment properties
 * (model.py:1678-1743, :1855-1862).  Floats are float like the reference's properties. */
typedef struct pd_config {
    uint32_t struct_size;           /* = sizeof(pd_config), checked                 */
    uint32_t width;                 /* env_max (:211, :1679); power of two, >= 16   */
    uint32_t row0;                  /* first global row owned by this handle        */
    uint32_t rows;                  /* rows owned (== width for a whole world)      */
    uint64_t seed;                  /* RANDOM_SEED (:39) -> Philox key              */
    uint64_t max_agents;            /* AGENT_HARD_LIMIT (:49, :1743)                */
    float payoff_cc;                /* :93,  :1700 */
    float payoff_cd;                /* :97,  :1701 */
    float payoff_dc;                /* :95,  :1702 */
    float payoff_dd;                /* :99,  :1703 */
    float env_noise;                /* :117, :1683 */
    float travel_cost;              /* :106, :1716 */
    float cost_of_living;           /* :78,  :1734 */
    float reproduce_min_energy;     /* :81,  :1727 */
    float reproduce_cost;           /* :83,  :1731 */
    float reproduction_inheritence; /* :88,  :1740 */
    float init_energy_mu;           /* :111, :1735 */
    float init_energy_sigma;        /* :112, :1736 */
    float init_energy_min;          /* :115, :1737 */
    float max_energy;               /* :109, :1705 */
    float mutation_rate;            /* :164, :1738 */
    float strategy_weights[4];      /* AGENT_WEIGHTS (:218-220), COOP/DEFECT/TFT/RANDOM */
    uint8_t strategy_pure;          /* :158, :1687 */
    uint8_t strategy_per_trait;     /* :161, :1685 */
    uint8_t max_children_per_step;  /* :90,  :1742 */
    uint8_t collect_stats;          /* 1: maintain the pd_step_stats counters        */
    int32_t device;                 /* CUDA device ordinal                           */
    int32_t sparse_ctas;            /* 0 = auto; CTAs of the persistent sparse kernels */
    uint32_t ghost_rows;            /* row slab only: ghost rows above AND below (multiple of 16, >= 64) */
    uint32_t world_height;          /* 0 = width (square, like the reference); else a power of two: the world is
                                       width x world_height cells (used for weak-scaling series of row slabs) */
    int32_t use_graphs;             /* 0 = auto (whole worlds <= 2^22 cells), 1 = always, -1 = never: replay one
                                       captured CUDA graph per step instead of 9 launches.  Only on explicit
                                       (capturable) streams and while per-kernel profiling is off. */
    uint32_t debug_select_cap;      /* 0 = default; tests: cap of the cull select's shared-memory stage, to force its
                                       global-memory fallback */
    uint32_t debug_tie_shift;       /* 0 = default (3); tests: two reproduction claims settled inside a tile count as a
                                       near-tie -- the tile's claims then take the exact route through the claim plane --
                                       when their priorities agree above this bit: 24..31 forces that route often */
} pd_config;

/* Per-step event counters of the LAST executed step (debug / parity aid; the
 * reference's only equivalent is VERBOSE_OUTPUT, model.py:1539-1548). */
typedef struct pd_step_stats {
    uint64_t agents_start, agents_end;
    uint64_t play_deaths, movers, moved, travel_deaths;
    uint64_t births, culled, living_deaths;
    uint64_t risk_list, move_losers, repro_losers;
} pd_step_stats;

typedef struct pd_sim pd_sim;

const char* pd_last_error(void);
int pd_abi_version(void);

/* Fill *cfg with the reference defaults (model.py:39-169) for a width x width world. */
int pd_default_config(pd_config* cfg, uint32_t width);

int pd_create(const pd_config* cfg, pd_sim** out);
int pd_destroy(pd_sim* sim);

/* Caller-owned DEVICE planes, row-major [rows][width], borrowed for the handle's lifetime:
 *   energy  float32  agent energy (model.py:1636); 0 on empty cells
 *   strat   uint8    4 x 2-bit agent_strategies[t] at bits 2t..2t+1 (model.py:1641)
 *   tf      uint8    PD_TF_OCCUPIED | agent_trait
 * Replaces the FLAMEGPU2 agent state list (make_core_agent, model.py:1630-1664). */
int pd_bind_planes(pd_sim* sim, float* energy, uint8_t* strat, uint8_t* tf);

/* init_fn (model.py:1423-1524) on the GPU: exactly n_agents on distinct uniformly random
 * cells, energy clamp(N(mu,sigma),min,max), uniform trait, weighted strategies by mode. */
int pd_init_population(pd_sim* sim, uint64_t n_agents, void* stream);

/* Call after writing the bound planes yourself: rebuilds internal lists and counts. */
int pd_sync_state(pd_sim* sim, void* stream);

/* Run n_steps model steps (main layers 1-7, model.py:2140-2163) asynchronously on `stream`.
 * want_counts = 1 also produces the 16 strategy counts of step_fn (:1405-1419) for the LAST of
 * the n_steps, want_counts = 2 for every step (the population total is always maintained). */
int pd_step(pd_sim* sim, uint32_t n_steps, void* stream, int want_counts);

/* population_strat_count[16] (bin = 4*my + other, model.py:1413-1419) and the agent count
 * (logCount, model.py:185).  Synchronises `stream` only.  Either pointer may be NULL. */
int pd_read_counts(pd_sim* sim, uint64_t counts[16], uint64_t* n_agents, void* stream);

int pd_read_stats(pd_sim* sim, pd_step_stats* out, void* stream);

/* ---- One world in row slabs over several GPUs (one handle per slab) ---------------------------
 * A slab owns rows [row0, row0 + rows) of the width x width world; its planes hold
 * rows + 2*ghost_rows rows: ghost_rows rows of the upper neighbour slab, the owned rows, ghost_rows
 * rows of the lower neighbour slab (periodic ring).  One model step is
 *     caller: refresh the ghost rows of energy/strat/tf from the neighbour slabs (NVLink copy)
 *     PD_PART_PRE            play, all travel rounds, reproduction claims (with the cost of living) and round 1;
 *                            publishes glb[0..2] and the cull histogram of round 1's newborns in ghist
 *     caller: sum glb and ghist over the slabs (ONE all-reduce if the two are views of one buffer)
 *     PD_PART_RETRY r=1..7   one further reproduction round, gated by the global population; caller sums glb.
 *                            Only needed while glb[0] + glb[1] <= max_agents and glb[2] > 0 (exit_god_fn).
 *     PD_PART_HIST           only if retry rounds ran: the histogram again;  caller sums ghist
 *     PD_PART_CAND           this slab's cull candidates into cand_out;  caller gathers all blocks into cand_all
 *     PD_PART_APPLY          cull, bookkeeping; publishes glb[3..19] (agents, 16 counts)
 * CAND and APPLY with arg = 1 are SPECULATIVE: queued before the caller has looked at the gate words, they do
 * nothing on the device if a retry round is due (the caller then runs RETRY ... CAND, APPLY with arg = 0), and
 * if none is due the caller confirms the step with PD_PART_DONE (no launch).  The host's read of the gate is
 * then off the critical path: the device never waits for it in the saturated state.
 * The ghost zone is wide enough that every other phase needs no communication (DESIGN.md "Multi-GPU");
 * results are bit-identical to the whole-world run because all draws are keyed by the global cell. */
#define PD_PART_PRE 0
#define PD_PART_RETRY 1
#define PD_PART_HIST 2
#define PD_PART_CAND 3
#define PD_PART_APPLY 4
#define PD_PART_DONE 5
#define PD_GLB_WORDS 20
#define PD_HIST_WORDS 2048
#define PD_MAX_WORLD 256   /* most slabs one world can be cut into */
/* glb: PD_GLB_WORDS u64; ghist: PD_HIST_WORDS u32; cand_out: 1 + cand_cap u64; cand_all: world * (1 + cand_cap) u64;
 * all DEVICE buffers owned by the caller (torch tensors) so that it can run collectives on them. */
int pd_bind_comm(pd_sim* sim, uint64_t* glb, uint32_t* ghist, uint64_t* cand_out, const uint64_t* cand_all,
                 uint32_t cand_cap, uint32_t world);
int pd_step_part(pd_sim* sim, int part, int arg, void* stream, int want_counts);

/* Halo strips for an exchange through peer memory (NVLink loads from a buffer the neighbour slab exposes, e.g.
 * torch symmetric memory): a strip holds ghost_rows rows of the three planes, packed energy | strat | tf
 * (pd_halo_bytes bytes, 16-byte aligned).  pd_halo_pack writes this slab's first / last owned rows into two
 * strips; pd_halo_unpack fills the top / bottom ghost rows from the strips of the upper / lower neighbour --
 * these may be PEER device pointers.  The caller orders pack -> (all slabs packed) -> unpack. */
int pd_halo_bytes(pd_sim* sim, uint64_t* strip_bytes);
int pd_halo_pack(pd_sim* sim, void* first_rows, void* last_rows, void* stream);
int pd_halo_unpack(pd_sim* sim, const void* top_ghost, const void* bottom_ghost, void* stream);

/* Sum over all executed steps of the population at the START of the step (the numerator of
 * the agent-steps/s metric).  Synchronises `stream`. */
int pd_read_agent_steps(pd_sim* sim, uint64_t* agent_steps, void* stream);

/* Device-resident step log (the reference logs population_strat_count and the agent count every
 * OUTPUT_EVERY_N_STEPS steps, model.py:180-187, reading them back on the host each time).  With the log
 * enabled, every step that computes counts (pd_step's want_counts: 1 = the call's last step, 2 = every
 * step) appends one row of PD_LOG_ROW_WORDS u64 -- step index after the step, agent count, 16 counts --
 * to a device buffer, so a whole run can be queued without a host synchronisation per logged step.
 * pd_read_step_log synchronises `stream`, copies up to max_rows rows (oldest first) and empties the log;
 * rows beyond the capacity are dropped and make the call return PD_ERR_OVERFLOW.  Whole worlds only.
 * Call pd_set_step_log while no step of this handle is in flight (before stepping, or after a read). */
#define PD_LOG_ROW_WORDS 18
int pd_set_step_log(pd_sim* sim, uint32_t capacity_rows, void* stream);
int pd_read_step_log(pd_sim* sim, uint64_t* rows, uint32_t max_rows, uint32_t* n_rows, void* stream);

/* Per-kernel device timing: with profiling on, pd_step records a CUDA event between every
 * two kernel launches; pd_read_profile synchronises `stream` and returns the accumulated
 * milliseconds per kernel, in launch order: begin_step, risk_resolve, play_travel, move_commit,
 * move_retry, claim_punish, repro_commit, repro_retry, finalize. */
#define PDABM_NUM_KERNELS 9
int pd_set_profile(pd_sim* sim, int on);
int pd_read_profile(pd_sim* sim, double kernel_ms[PDABM_NUM_KERNELS], uint32_t* steps, void* stream);

/* Step index of the next step to run (0-based; Philox counter word 2). */
int pd_get_step(pd_sim* sim, uint32_t* step);
int pd_set_step(pd_sim* sim, uint32_t step);

/* Number of model steps this handle ran by replaying a captured CUDA graph (pd_config.use_graphs)
 * rather than by launching the 9 kernels one by one. */
int pd_graph_replays(pd_sim* sim, uint64_t* steps);

/* HOST-buffer entry point (what a host-side caller without device buffers binds):
 * uploads the three planes from host memory, runs n_steps, downloads the planes back
 * (if the out pointers are non-NULL) and the counts.  Synchronous. */
int pd_step_host(pd_sim* sim, const float* h_energy_in, const uint8_t* h_strat_in,
                 const uint8_t* h_tf_in, uint32_t n_steps, float* h_energy_out,
                 uint8_t* h_strat_out, uint8_t* h_tf_out, uint64_t counts[16],
                 uint64_t* n_agents, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDABM_H_ */

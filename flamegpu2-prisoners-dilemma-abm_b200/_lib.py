"""ctypes binding of libpdabm.so (include/pdabm.h).  Fails loudly if the library is missing:
there is no CPU fallback for the pipeline."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PDABM_LIB: developer hook for A/B timing of two builds of the library in one GPU session
LIB_PATH = os.environ.get("PDABM_LIB") or os.path.join(_HERE, "libpdabm.so")

PD_OK = 0
PD_TF_TRAIT_MASK = 0x03
PD_TF_OCCUPIED = 0x80
PD_PART_PRE, PD_PART_RETRY, PD_PART_HIST, PD_PART_CAND, PD_PART_APPLY, PD_PART_DONE = 0, 1, 2, 3, 4, 5
PD_GLB_WORDS, PD_HIST_WORDS = 20, 2048


class pd_config(C.Structure):
    """struct pd_config, include/pdabm.h (field order and types must match)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("width", C.c_uint32),
        ("row0", C.c_uint32),
        ("rows", C.c_uint32),
        ("seed", C.c_uint64),
        ("max_agents", C.c_uint64),
        ("payoff_cc", C.c_float),
        ("payoff_cd", C.c_float),
        ("payoff_dc", C.c_float),
        ("payoff_dd", C.c_float),
        ("env_noise", C.c_float),
        ("travel_cost", C.c_float),
        ("cost_of_living", C.c_float),
        ("reproduce_min_energy", C.c_float),
        ("reproduce_cost", C.c_float),
        ("reproduction_inheritence", C.c_float),
        ("init_energy_mu", C.c_float),
        ("init_energy_sigma", C.c_float),
        ("init_energy_min", C.c_float),
        ("max_energy", C.c_float),
        ("mutation_rate", C.c_float),
        ("strategy_weights", C.c_float * 4),
        ("strategy_pure", C.c_uint8),
        ("strategy_per_trait", C.c_uint8),
        ("max_children_per_step", C.c_uint8),
        ("collect_stats", C.c_uint8),
        ("device", C.c_int32),
        ("sparse_ctas", C.c_int32),
        ("ghost_rows", C.c_uint32),
        ("world_height", C.c_uint32),
        ("use_graphs", C.c_int32),
        ("debug_select_cap", C.c_uint32),
        ("debug_tie_shift", C.c_uint32),
    ]


class pd_step_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "agents_start", "agents_end", "play_deaths", "movers", "moved", "travel_deaths",
        "births", "culled", "living_deaths", "risk_list", "move_losers", "repro_losers")]


# every symbol include/pdabm.h declares (tests check the library exports all of them)
EXPORTS = (
    "pd_last_error", "pd_abi_version", "pd_default_config", "pd_create", "pd_destroy",
    "pd_bind_planes", "pd_init_population", "pd_sync_state", "pd_step", "pd_read_counts",
    "pd_read_stats", "pd_get_step", "pd_set_step", "pd_step_host", "pd_read_agent_steps",
    "pd_set_profile", "pd_read_profile", "pd_bind_comm", "pd_step_part", "pd_graph_replays",
    "pd_set_step_log", "pd_read_step_log", "pd_halo_bytes", "pd_halo_pack", "pd_halo_unpack",
)

_lib = None


def load() -> C.CDLL:
    """Load libpdabm.so and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the agent pipeline.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    lib.pd_last_error.restype = C.c_char_p
    lib.pd_last_error.argtypes = []
    lib.pd_abi_version.restype = i32
    lib.pd_abi_version.argtypes = []
    lib.pd_default_config.restype = i32
    lib.pd_default_config.argtypes = [C.POINTER(pd_config), u32]
    lib.pd_create.restype = i32
    lib.pd_create.argtypes = [C.POINTER(pd_config), C.POINTER(vp)]
    lib.pd_destroy.restype = i32
    lib.pd_destroy.argtypes = [vp]
    lib.pd_bind_planes.restype = i32
    lib.pd_bind_planes.argtypes = [vp, vp, vp, vp]
    lib.pd_init_population.restype = i32
    lib.pd_init_population.argtypes = [vp, u64, vp]
    lib.pd_sync_state.restype = i32
    lib.pd_sync_state.argtypes = [vp, vp]
    lib.pd_step.restype = i32
    lib.pd_step.argtypes = [vp, u32, vp, i32]
    lib.pd_read_counts.restype = i32
    lib.pd_read_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), vp]
    lib.pd_read_stats.restype = i32
    lib.pd_read_stats.argtypes = [vp, C.POINTER(pd_step_stats), vp]
    lib.pd_get_step.restype = i32
    lib.pd_get_step.argtypes = [vp, C.POINTER(u32)]
    lib.pd_set_step.restype = i32
    lib.pd_set_step.argtypes = [vp, u32]
    lib.pd_step_host.restype = i32
    lib.pd_step_host.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp, C.POINTER(u64), C.POINTER(u64), vp]
    lib.pd_read_agent_steps.restype = i32
    lib.pd_read_agent_steps.argtypes = [vp, C.POINTER(u64), vp]
    lib.pd_set_profile.restype = i32
    lib.pd_set_profile.argtypes = [vp, i32]
    lib.pd_read_profile.restype = i32
    lib.pd_read_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u32), vp]
    lib.pd_bind_comm.restype = i32
    lib.pd_bind_comm.argtypes = [vp, vp, vp, vp, vp, u32, u32]
    lib.pd_step_part.restype = i32
    lib.pd_step_part.argtypes = [vp, i32, i32, vp, i32]
    lib.pd_halo_bytes.restype = i32
    lib.pd_halo_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.pd_halo_pack.restype = i32
    lib.pd_halo_pack.argtypes = [vp, vp, vp, vp]
    lib.pd_halo_unpack.restype = i32
    lib.pd_halo_unpack.argtypes = [vp, vp, vp, vp]
    lib.pd_set_step_log.restype = i32
    lib.pd_set_step_log.argtypes = [vp, u32, vp]
    lib.pd_read_step_log.restype = i32
    lib.pd_read_step_log.argtypes = [vp, C.POINTER(u64), u32, C.POINTER(u32), vp]
    lib.pd_graph_replays.restype = i32
    lib.pd_graph_replays.argtypes = [vp, C.POINTER(u64)]
    lib.pd_debug_philox.restype = i32
    lib.pd_debug_philox.argtypes = [vp, vp, i32]
    _lib = lib
    return lib

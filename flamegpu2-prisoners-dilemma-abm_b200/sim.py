"""Host-side handle over libpdabm.so: the Python object that takes the place of
``pyflamegpu.CUDASimulation`` for this model (/root/reference/src/model.py:1780-1798, :2176).

torch is used for device buffers and streams only; every computation of the step happens in
the hand-written CUDA kernels behind the C ABI (include/pdabm.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from ._lib import pd_config, pd_step_stats, PD_TF_OCCUPIED, PD_TF_TRAIT_MASK


class PdabmError(RuntimeError):
    pass


@dataclass
class SimConfig:
    """The reference's tunables under their environment-property names
    (model.py:1678-1743; module constants model.py:39-169)."""
    width: int = 512                       # ENV_MAX (:211)
    height: int = -1                       # world height; -1 = width (the reference's worlds are square)
    seed: int = 12345                      # RANDOM_SEED (:39)
    init_agent_count: int = -1             # INIT_AGENT_COUNT (:45); -1 = int(cells * 0.16)
    max_agents: int = -1                   # AGENT_HARD_LIMIT (:49); -1 = int(cells * 0.5)
    payoff_cc: float = 3.0
    payoff_dc: float = 5.0
    payoff_cd: float = -1.0
    payoff_dd: float = 0.0
    env_noise: float = 0.0
    cost_of_living: float = 1.0
    travel_cost: float = 0.5
    reproduce_min_energy: float = 100.0
    reproduce_cost: float = 50.0
    reproduction_inheritence: float = 0.0
    max_children_per_step: int = 1
    max_energy: float = 150.0
    init_energy_mu: float = 50.0
    init_energy_sigma: float = 10.0
    init_energy_min: float = 5.0
    mutation_rate: float = 0.0
    strategy_pure: int = 0
    strategy_per_trait: int = 0
    strategy_weights: List[float] = field(default_factory=lambda: [0.25, 0.25, 0.25, 0.25])
    device: int = 0
    collect_stats: bool = False
    sparse_ctas: int = 0
    use_graphs: int = 0                    # 0 auto, 1 always, -1 never (one CUDA graph per step; needs a stream)
    debug_select_cap: int = 0              # tests only: force the cull select's global-memory fallback
    debug_tie_shift: int = 0               # tests only: force the claim-plane route of tile-local reproduction claims

    def __post_init__(self):
        if self.height < 0:
            self.height = self.width
        cells = self.width * self.height
        if self.init_agent_count < 0:
            self.init_agent_count = int(cells * 0.16)
        if self.max_agents < 0:
            self.max_agents = int(cells * 0.5)

    def to_c(self, row0: int = 0, rows: Optional[int] = None, ghost_rows: int = 0) -> pd_config:
        c = pd_config()
        c.struct_size = C.sizeof(pd_config)
        c.width = self.width
        c.row0 = row0
        c.rows = self.height if rows is None else rows
        c.ghost_rows = ghost_rows
        c.world_height = 0 if self.height == self.width else self.height
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.max_agents = self.max_agents
        for k in ("payoff_cc", "payoff_cd", "payoff_dc", "payoff_dd", "env_noise", "travel_cost",
                  "cost_of_living", "reproduce_min_energy", "reproduce_cost",
                  "reproduction_inheritence", "init_energy_mu", "init_energy_sigma",
                  "init_energy_min", "max_energy", "mutation_rate"):
            setattr(c, k, float(getattr(self, k)))
        for i in range(4):
            c.strategy_weights[i] = float(self.strategy_weights[i])
        c.strategy_pure = int(bool(self.strategy_pure))
        c.strategy_per_trait = int(bool(self.strategy_per_trait))
        c.max_children_per_step = int(self.max_children_per_step)
        c.collect_stats = int(bool(self.collect_stats))
        c.device = int(self.device)
        c.sparse_ctas = int(self.sparse_ctas)
        c.use_graphs = int(self.use_graphs)
        c.debug_select_cap = int(self.debug_select_cap)
        c.debug_tie_shift = int(self.debug_tie_shift)
        return c


def pack_strategies(strat4: np.ndarray) -> np.ndarray:
    """[..., 4] strategies (0..3) -> packed u8 plane (slot t at bits 2t..2t+1)."""
    s = np.asarray(strat4, dtype=np.uint8)
    return (s[..., 0] | (s[..., 1] << 2) | (s[..., 2] << 4) | (s[..., 3] << 6)).astype(np.uint8)


def unpack_strategies(packed: np.ndarray) -> np.ndarray:
    p = np.asarray(packed, dtype=np.uint8)
    return np.stack([(p >> (2 * t)) & 3 for t in range(4)], axis=-1).astype(np.uint8)


class Simulation:
    """One world on one GPU.  ``energy``, ``strat`` and ``tf`` are torch CUDA tensors of
    shape [width, width] ([y, x]) that the library borrows (pd_bind_planes)."""

    def __init__(self, config: SimConfig, stream=None, slab=None):
        """slab = (row0, rows, ghost_rows) makes this handle one row slab of the world (see decomp.py);
        its planes then have rows + 2*ghost_rows rows."""
        import torch
        if not torch.cuda.is_available():
            raise PdabmError("CUDA device required: the agent pipeline has no CPU path")
        self.cfg = config
        self.lib = _lib.load()
        self.torch = torch
        self.device = torch.device("cuda", config.device)
        W = config.width
        self.slab = slab
        row0, rows, ghost = (0, config.height, 0) if slab is None else slab
        self.row0, self.rows, self.ghost_rows = row0, rows, ghost
        H = rows + 2 * ghost
        with torch.cuda.device(self.device):
            self.energy = torch.zeros((H, W), dtype=torch.float32, device=self.device)
            self.strat = torch.zeros((H, W), dtype=torch.uint8, device=self.device)
            self.tf = torch.zeros((H, W), dtype=torch.uint8, device=self.device)
            torch.cuda.synchronize(self.device)
        self._stream = stream
        self._log_cap = 0
        self._h = C.c_void_p()
        ccfg = config.to_c(row0, rows, ghost)
        self._check(self.lib.pd_create(C.byref(ccfg), C.byref(self._h)))
        self._check(self.lib.pd_bind_planes(self._h, self.energy.data_ptr(), self.strat.data_ptr(),
                                            self.tf.data_ptr()))

    # -- plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise PdabmError(f"libpdabm error {rc}: {self.lib.pd_last_error().decode()}")

    @property
    def stream_ptr(self) -> int:
        return 0 if self._stream is None else int(self._stream.cuda_stream)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.pd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference-facing surface
    def init_population(self, n_agents: Optional[int] = None):
        """init_fn (model.py:1423-1524)."""
        n = self.cfg.init_agent_count if n_agents is None else n_agents
        self._check(self.lib.pd_init_population(self._h, n, self.stream_ptr))

    def set_state(self, occ, energy, trait, strat4):
        """Upload a world given as NumPy planes ([y,x]; strat4 is [y,x,4])."""
        torch = self.torch
        tf = (np.asarray(occ, np.uint8) * np.uint8(PD_TF_OCCUPIED)) | (np.asarray(trait, np.uint8) & 3)
        tf = np.where(np.asarray(occ) != 0, tf, 0).astype(np.uint8)
        if self._stream is not None:  # the copies below run on torch's current stream: no step may still be in flight
            self._stream.synchronize()
        self.tf.copy_(torch.from_numpy(tf))
        self.energy.copy_(torch.from_numpy(np.asarray(energy, np.float32)))
        self.strat.copy_(torch.from_numpy(pack_strategies(strat4)))
        torch.cuda.synchronize(self.device)
        self._check(self.lib.pd_sync_state(self._h, self.stream_ptr))

    def step(self, n_steps: int = 1, want_counts=True):
        """Main layers 1-7 (model.py:2140-2163) n_steps times; asynchronous.  want_counts: False, True (the
        strategy counts of the last step) or "every" (of every step, for the device-resident step log)."""
        wc = 2 if want_counts == "every" else int(bool(want_counts))
        self._check(self.lib.pd_step(self._h, int(n_steps), self.stream_ptr, wc))

    def counts(self):
        """(population_strat_count[16], agent count) -- step_fn (model.py:1405-1419) and logCount (:185)."""
        arr = (C.c_uint64 * 16)()
        n = C.c_uint64()
        self._check(self.lib.pd_read_counts(self._h, arr, C.byref(n), self.stream_ptr))
        return np.array(list(arr), dtype=np.int64), int(n.value)

    def stats(self) -> Dict[str, int]:
        st = pd_step_stats()
        self._check(self.lib.pd_read_stats(self._h, C.byref(st), self.stream_ptr))
        return {n: int(getattr(st, n)) for n, _ in pd_step_stats._fields_}

    KERNELS = ("begin_step", "risk_resolve", "play_travel", "move_commit", "move_retry", "claim_punish",
               "repro_commit", "repro_retry", "finalize")

    def agent_steps(self) -> int:
        """Sum over executed steps of the population at the start of the step."""
        v = C.c_uint64()
        self._check(self.lib.pd_read_agent_steps(self._h, C.byref(v), self.stream_ptr))
        return int(v.value)

    @property
    def step_index(self) -> int:
        """Index of the next step to run (a Philox key word: a run is resumed by restoring it with the planes)."""
        v = C.c_uint32()
        self._check(self.lib.pd_get_step(self._h, C.byref(v)))
        return int(v.value)

    @step_index.setter
    def step_index(self, step: int):
        self._check(self.lib.pd_set_step(self._h, int(step)))

    @staticmethod
    def _ckpt_path(path: str) -> str:
        return path if path.endswith(".npz") else path + ".npz"  # (np.savez appends the suffix by itself)

    def save_checkpoint(self, path: str):
        """World + step index as one .npz: everything a run needs to continue bit-identically.  (A row slab is
        checkpointed through its DecomposedWorld: a slab alone does not hold a world.)"""
        if self.slab is not None:
            raise PdabmError("save_checkpoint on a row slab: checkpoint the DecomposedWorld instead")
        pl = self.planes()
        np.savez_compressed(self._ckpt_path(path), occ=pl["occ"], energy=pl["energy"], trait=pl["trait"], strat=pl["strat"],
                            step=np.int64(self.step_index), seed=np.uint64(self.cfg.seed & 0xFFFFFFFFFFFFFFFF))

    def load_checkpoint(self, path: str):
        if self.slab is not None:
            raise PdabmError("load_checkpoint on a row slab: restore the DecomposedWorld instead")
        with np.load(self._ckpt_path(path)) as z:
            if int(z["seed"]) != (self.cfg.seed & 0xFFFFFFFFFFFFFFFF):
                raise PdabmError("checkpoint was written with a different seed: the run would not continue identically")
            self.set_state(z["occ"], z["energy"], z["trait"], z["strat"])
            self.step_index = int(z["step"])

    def enable_step_log(self, capacity_rows: int):
        """Keep (step, agent count, 16 counts) of every counted step on the device (0 = off)."""
        self._check(self.lib.pd_set_step_log(self._h, int(capacity_rows), self.stream_ptr))
        self._log_cap = int(capacity_rows)

    def read_step_log(self) -> np.ndarray:
        """Rows logged since the last call, oldest first: int64 [n, 18] = step, agents, counts[16]."""
        buf = (C.c_uint64 * (18 * max(1, self._log_cap)))()
        n = C.c_uint32()
        self._check(self.lib.pd_read_step_log(self._h, buf, self._log_cap, C.byref(n), self.stream_ptr))
        return np.frombuffer(buf, dtype=np.uint64, count=18 * n.value).reshape(n.value, 18).astype(np.int64)

    def graph_replays(self) -> int:
        """Steps run by replaying a captured CUDA graph (small worlds on an explicit stream)."""
        v = C.c_uint64()
        self._check(self.lib.pd_graph_replays(self._h, C.byref(v)))
        return int(v.value)

    def set_profile(self, on: bool):
        self._check(self.lib.pd_set_profile(self._h, int(bool(on))))

    def read_profile(self):
        """({kernel: total ms}, profiled steps)."""
        ms = (C.c_double * len(self.KERNELS))()
        n = C.c_uint32()
        self._check(self.lib.pd_read_profile(self._h, ms, C.byref(n), self.stream_ptr))
        return {k: float(ms[i]) for i, k in enumerate(self.KERNELS)}, int(n.value)

    def planes(self) -> Dict[str, np.ndarray]:
        """Download the world as NumPy planes (same keys as the oracle's ``planes()``)."""
        self.torch.cuda.synchronize(self.device)
        g, r = self.ghost_rows, self.rows
        tf = self.tf[g:g + r].cpu().numpy()  # owned rows only
        occ = (tf & PD_TF_OCCUPIED) != 0
        return {
            "occ": occ.astype(np.uint8),
            "energy": np.where(occ, self.energy[g:g + r].cpu().numpy(), np.float32(0)).astype(np.float32),
            "trait": np.where(occ, tf & PD_TF_TRAIT_MASK, 0).astype(np.uint8),
            "strat": np.where(occ[..., None], unpack_strategies(self.strat[g:g + r].cpu().numpy()), 0).astype(np.uint8),
            "tf_raw": tf,
        }

    def export_state(self, path: str):
        """Dump the world for visual inspection (takes the place of the OpenGL visualiser,
        configure_visualisation, model.py:1753-1777): a .npz with occ / trait / strat / energy planes, or,
        for a .ppm path, an image coloured by trait like the reference (agent_color = trait, :1470)."""
        pl = self.planes()
        if path.endswith(".ppm"):
            palette = np.array([[228, 26, 28], [55, 126, 184], [77, 175, 74], [152, 78, 163]], np.uint8)  # SET1
            img = np.full(pl["occ"].shape + (3,), 26, np.uint8)  # VISUALISATION_BG_RGB 0.1
            occ = pl["occ"].astype(bool)
            img[occ] = palette[pl["trait"][occ]]
            with open(path, "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
                f.write(img.tobytes())
        else:
            np.savez_compressed(path, occ=pl["occ"], trait=pl["trait"], strat=pl["strat"], energy=pl["energy"],
                                step=self.step_index)

    def step_host(self, h_energy, h_strat, h_tf, n_steps=1, out_energy=None, out_strat=None, out_tf=None):
        """pd_step_host: host buffers in, host buffers out (the e2e path of bench.py)."""
        arr = (C.c_uint64 * 16)()
        n = C.c_uint64()
        ptr = lambda t: 0 if t is None else t.data_ptr()
        self._check(self.lib.pd_step_host(self._h, ptr(h_energy), ptr(h_strat), ptr(h_tf), int(n_steps),
                                          ptr(out_energy), ptr(out_strat), ptr(out_tf), arr, C.byref(n),
                                          self.stream_ptr))
        return np.array(list(arr), dtype=np.int64), int(n.value)

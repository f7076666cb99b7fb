"""One world row-decomposed over several GPUs (BASELINE config 5).

The reference cannot do this at all (FLAMEGPU2 simulations are single-GPU and its bucket ids are
32 bit, /root/reference/src/model.py:214, :338-339).  Here every rank owns a slab of rows plus a
ghost zone of `ghost_rows` rows on each side.  Per step there is ONE halo exchange (the three state
planes' boundary rows over NVLink) and a handful of tiny collectives for the two global decisions
of the model -- the carrying-capacity gate of the reproduction rounds (exit_god_fn, :1607-1627) and
the cull (:1339-1356):

    exchange ghost rows -> PRE -> sum(glb + ghist, one all-reduce) -> [7 x [RETRY r -> sum(glb)] -> HIST -> sum(ghist)]
        -> CAND -> gather(candidates) -> APPLY [-> sum(counts)]

(the bracketed retry branch only runs while the global population is within the cap, i.e. never in the
saturated state: a step then costs one halo exchange, one 8 KB all-reduce and one all-gather).  CAND / APPLY
are queued SPECULATIVELY, before the host has read the gate words: on the device they do nothing if a retry
round turns out to be due, so the host's read never stalls the device in the saturated state.

Everything else runs on the slab alone: the ghost zone (64 rows) is wider than the deepest
dependency cone of one step (8 play sub-steps + 2 sweeps + 7 retry rounds of travel and of
reproduction, <= 56 rows), so the owned rows come out bit-identical to a single-GPU run -- all
random draws are keyed by the GLOBAL cell index.

Three transports: `PeerComm` (one slab per process/GPU; the halo goes through peer memory -- every rank packs
its boundary rows into a symmetric-memory outbox and its neighbours LOAD them over NVLink inside one kernel;
the small collectives stay on NCCL), `DistComm` (the same with NCCL send/recv for the halo; also the CPU/gloo
transport of the tests) and `LocalComm` (all slabs in one process on one GPU -- used by the tests to check
decomposition invariance without a multi-GPU box).
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from ._lib import (PD_PART_PRE, PD_PART_RETRY, PD_PART_HIST, PD_PART_CAND, PD_PART_APPLY, PD_PART_DONE,
                   PD_GLB_WORDS, PD_HIST_WORDS)
from .sim import SimConfig, Simulation, PdabmError
from .dist import slab_rows, ring_neighbours

GHOST_ROWS = 64
CAND_CAP = 1 << 14   # keys per slab in the cull's threshold bin: newborns / 2048 (+ Poisson), 9 k at 2^29 cells per slab


class _Slab:
    """A Simulation handle for one slab plus its communication buffers."""

    def __init__(self, cfg: SimConfig, rank: int, world: int, ghost_rows: int, cand_cap: int, stream=None):
        import torch
        row0, rows = slab_rows(cfg.height, rank, world)
        if ghost_rows > rows:
            raise PdabmError(f"ghost zone ({ghost_rows} rows) larger than the slab ({rows} rows): use fewer slabs")
        self.rank, self.world = rank, world
        self.sim = Simulation(cfg, stream=stream, slab=(row0, rows, ghost_rows))
        dev = self.sim.device
        # the gate words and the cull histogram are views of ONE buffer, so that one all-reduce sums both: the
        # 2048 u32 bins travel as 1024 i64 words (a bin's sum over the slabs is < 2^32, so no carry crosses a word)
        self.wire = torch.zeros(PD_GLB_WORDS + PD_HIST_WORDS // 2, dtype=torch.int64, device=dev)
        self.glb = self.wire[:PD_GLB_WORDS]
        self.ghist = self.wire[PD_GLB_WORDS:].view(torch.int32)
        self.gate_host = torch.zeros(3, dtype=torch.int64).pin_memory()
        self.cand_out = torch.zeros(1 + cand_cap, dtype=torch.int64, device=dev)
        self.cand_all = torch.zeros(world * (1 + cand_cap), dtype=torch.int64, device=dev)
        self.sim._check(self.sim.lib.pd_bind_comm(self.sim._h, self.glb.data_ptr(), self.ghist.data_ptr(),
                                                  self.cand_out.data_ptr(), self.cand_all.data_ptr(),
                                                  cand_cap, world))

    def part(self, part: int, arg: int = 0, want_counts: bool = False):
        s = self.sim
        s._check(s.lib.pd_step_part(s._h, part, arg, s.stream_ptr, int(want_counts)))

    def planes3(self):
        return (self.sim.energy, self.sim.strat, self.sim.tf)


class LocalComm:
    """All slabs in this process (one GPU): collectives are plain tensor ops."""

    def __init__(self, world: int):
        self.world = world
        self.local_ranks = list(range(world))

    def exchange(self, slabs: List[_Slab]):
        g = slabs[0].sim.ghost_rows
        r = slabs[0].sim.rows
        for s in slabs:
            up, down = ring_neighbours(s.rank, self.world)
            for mine, theirs_up, theirs_down in zip(s.planes3(), slabs[up].planes3(), slabs[down].planes3()):
                mine[0:g].copy_(theirs_up[r:r + g])              # upper neighbour's last owned rows
                mine[g + r:g + r + g].copy_(theirs_down[g:2 * g])  # lower neighbour's first owned rows

    def allreduce(self, tensors):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t
        for t in tensors:
            t.copy_(total)

    def allgather(self, outs, ins):
        import torch
        cat = torch.cat(ins)
        for o in outs:
            o.copy_(cat)


class DistComm:
    """One slab per process: torch.distributed (NCCL over NVLink on GPUs, gloo on CPU in the tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.local_ranks = [self.rank]

    def exchange(self, slabs: List[_Slab]):
        dist = self.dist
        s = slabs[0]
        g, r = s.sim.ghost_rows, s.sim.rows
        up, down = ring_neighbours(self.rank, self.world)
        ops = []
        for t in s.planes3():
            # Per peer the messages match in FIFO order, and with 2 slabs `up` and `down` are the same peer:
            # it sends (its first rows, its last rows), so I must receive (bottom ghosts, top ghosts) in that order.
            ops.append(dist.P2POp(dist.isend, t[g:2 * g], up))              # my first owned rows -> upper slab's bottom ghosts
            ops.append(dist.P2POp(dist.isend, t[r:r + g], down))            # my last owned rows  -> lower slab's top ghosts
            ops.append(dist.P2POp(dist.irecv, t[g + r:g + r + g], down))    # lower slab's first owned rows
            ops.append(dist.P2POp(dist.irecv, t[0:g], up))                  # upper slab's last owned rows
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def allreduce(self, tensors):
        self.dist.all_reduce(tensors[0], op=self.dist.ReduceOp.SUM)

    def allgather(self, outs, ins):
        self.dist.all_gather_into_tensor(outs[0], ins[0])


class PeerComm(DistComm):
    """DistComm with the halo exchange through peer memory (torch symmetric memory: cuMem allocations mapped into
    every rank of the node, NVLink / NVSwitch underneath).  Per step and rank: ONE kernel packs the boundary rows
    of the three planes into this rank's outbox, one device-side barrier, ONE kernel loads the two neighbours'
    strips straight into the ghost rows.  The outbox is double-buffered by step parity, so a rank that runs ahead
    into the next step's pack never overwrites a strip a neighbour is still reading (it cannot pass the NEXT
    barrier before that neighbour has arrived there, i.e. has finished reading).
    NCCL send/recv needs ~0.28 ms for the same 12.6 MB (12 point-to-point operations, profiles/README.md)."""

    def __init__(self):
        super().__init__()
        self.hdl = None
        self.parity = 0

    def _setup(self, s: _Slab):
        import ctypes as C
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        sim = s.sim
        nb = C.c_uint64()
        sim._check(sim.lib.pd_halo_bytes(sim._h, C.byref(nb)))
        self.strip = int(nb.value)
        total = 4 * self.strip                       # [parity][first rows | last rows]
        self.box = symm_mem.empty((total,), dtype=torch.uint8, device=sim.device)
        self.hdl = symm_mem.rendezvous(self.box, self.dist.group.WORLD)
        up, down = ring_neighbours(self.rank, self.world)
        self.up_box = self.hdl.get_buffer(up, (total,), torch.uint8)
        self.down_box = self.hdl.get_buffer(down, (total,), torch.uint8)

    def exchange(self, slabs: List[_Slab]):
        s = slabs[0]
        if self.hdl is None:
            self._setup(s)
        sim, n = s.sim, self.strip
        off = self.parity * 2 * n
        self.parity ^= 1
        mine = self.box.data_ptr() + off
        sim._check(sim.lib.pd_halo_pack(sim._h, mine, mine + n, sim.stream_ptr))
        self.hdl.barrier(channel=0)
        # top ghost rows = the upper slab's LAST owned rows, bottom ghost rows = the lower slab's FIRST owned rows
        sim._check(sim.lib.pd_halo_unpack(sim._h, self.up_box.data_ptr() + off + n, self.down_box.data_ptr() + off,
                                          sim.stream_ptr))


def best_comm():
    """PeerComm where torch symmetric memory works for this process group (NCCL ranks of one node), else
    DistComm -- a slower transport for the same data, never a different compute path."""
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
            return PeerComm()
        except Exception:
            pass
    return DistComm()


class DecomposedWorld:
    """`world` row slabs of one width x width world.  Same surface as Simulation where it matters:
    init_population(), step(), counts(), planes() (rank-local rows for DistComm, all rows for LocalComm)."""

    def __init__(self, cfg: SimConfig, comm, ghost_rows: int = GHOST_ROWS, cand_cap: int = CAND_CAP,
                 retry_rounds: int = 7, stream=None):
        import torch
        self.torch = torch
        self.cfg, self.comm = cfg, comm
        self.stream = stream
        self.retry_rounds = retry_rounds
        self.slabs = [_Slab(cfg, r, comm.world, ghost_rows, cand_cap, stream) for r in comm.local_ranks]
        self.retry_rounds_run = 0  # reproduction retry rounds executed so far (2 launches + 1 all-reduce each)
        self._counts = None
        self._glb_summed = True    # glb[3:] (agents + 16 counts) holds the sum over the slabs

    def close(self):
        for s in self.slabs:
            s.sim.close()

    def init_population(self, n_agents: Optional[int] = None):
        for s in self.slabs:
            s.sim.init_population(n_agents)
        self._publish_local_counts()

    def _publish_local_counts(self):
        import torch
        for s in self.slabs:
            c, n = s.sim.counts()
            s.glb[3:] = torch.tensor([n] + [int(v) for v in c], dtype=torch.int64, device=s.glb.device)
        self.comm.allreduce([s.glb[3:] for s in self.slabs])
        self._glb_summed = True

    def set_state_from(self, planes: Dict[str, np.ndarray]):
        """Upload a whole-world state (oracle-style planes) -- each slab takes its rows; test helper."""
        for s in self.slabs:
            sim = s.sim
            g, r, r0 = sim.ghost_rows, sim.rows, sim.row0
            HW = self.cfg.height
            rows = [(r0 - g + k) % HW for k in range(r + 2 * g)]
            sim.set_state(planes["occ"][rows], planes["energy"][rows], planes["trait"][rows], planes["strat"][rows])
        self._publish_local_counts()

    def _gate_request(self):
        """exit_god_fn (model.py:1614-1627) is decided on the host, like in the reference: (olds, births, parents
        still trying), summed over the slabs.  The read is asynchronous: a small pinned copy and an event."""
        s = self.slabs[0]
        s.gate_host.copy_(s.glb[:3], non_blocking=True)
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream(s.sim.device))
        return ev

    def _gate_open(self, ev=None):
        s = self.slabs[0]
        if ev is None:
            ev = self._gate_request()
        ev.synchronize()
        olds, births, active = (int(v) for v in s.gate_host.tolist())
        return olds + births <= self.cfg.max_agents and active > 0

    def step(self, n_steps: int = 1, want_counts: bool = True):
        comm, slabs = self.comm, self.slabs
        # everything -- kernels, halo copies, collectives -- is ordered on ONE stream: the handle's, if it has one
        ctx = self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
        with ctx:
            for k in range(n_steps):
                wc = want_counts and k + 1 == n_steps
                comm.exchange(slabs)
                for s in slabs:
                    s.part(PD_PART_PRE, 0, wc)
                comm.allreduce([s.wire for s in slabs])   # gate words + cull histogram of round 1's newborns
                # The rest of the step is queued before the host knows the gate: CAND / APPLY with arg = 1 do nothing
                # on the device if a retry round is due.  In the saturated state (gate closed) the host's read is
                # therefore never waited for by the device.
                ev = self._gate_request()
                self._finish(1, wc)
                if not self._gate_open(ev):
                    for s in slabs:
                        s.part(PD_PART_DONE)
                    continue
                for rnd in range(1, self.retry_rounds + 1):
                    # another round only while the GLOBAL population is within the cap and somebody is still trying
                    if rnd > 1 and not self._gate_open():
                        break
                    for s in slabs:
                        s.part(PD_PART_RETRY, rnd, wc)
                    comm.allreduce([s.glb for s in slabs])
                    self.retry_rounds_run += 1
                for s in slabs:   # the retry rounds added newborns: the histogram again
                    s.part(PD_PART_HIST, 0, wc)
                comm.allreduce([s.ghist for s in slabs])
                self._finish(0, wc)
            # agents + 16 counts of the last step: summed right away when the caller logs them, else on demand
            # (counts()) -- one small collective less per step, and less host work between a step's gate read and
            # the next step's first launches
            self._glb_summed = False
            if want_counts:
                self._sum_counts()

    def _sum_counts(self):
        if not self._glb_summed:
            ctx = self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
            with ctx:
                self.comm.allreduce([s.glb[3:] for s in self.slabs])
            self._glb_summed = True

    def _finish(self, speculative: int, wc: bool):
        comm, slabs = self.comm, self.slabs
        for s in slabs:
            s.part(PD_PART_CAND, speculative, wc)
        comm.allgather([s.cand_all for s in slabs], [s.cand_out for s in slabs])
        for s in slabs:
            s.part(PD_PART_APPLY, speculative, wc)

    def counts(self):
        """Global (population_strat_count[16], agent count) after the last step().  COLLECTIVE after a step() with
        want_counts=False: every rank must call it."""
        self._sum_counts()
        v = self.slabs[0].glb.cpu().numpy()
        for s in self.slabs:
            s.sim.counts()  # raises on a list overflow in any slab
        return v[4:20].astype(np.int64), int(v[3])

    def planes(self) -> Dict[str, np.ndarray]:
        """Owned rows of the local slabs, stacked in rank order."""
        parts = [s.sim.planes() for s in self.slabs]
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}

    def agent_steps(self) -> int:
        return sum(s.sim.agent_steps() for s in self.slabs)

"""world_size-2 gloo tests of the multi-GPU host logic (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdabm_b200 import dist as pdist
from pdabm_b200 import model
from pdabm_b200.sim import SimConfig


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        runs = model.configure_runplan(SimConfig(width=64, seed=10), multi_run_count=1, multi_run_steps=5)
        mine = pdist.shard_runs(runs, rank, world)
        # whole-job throughput: sum of units, max of time
        total, secs = pdist.reduce_throughput(agent_steps=1000.0 * (rank + 1), seconds=1.0 + rank)
        paths = pdist.gather_logs([f"rank{rank}/{r['subdir']}" for r in mine])
        q.put((rank, len(mine), total, secs, pdist.rank_world(), len(paths), pdist.slab_rows(64, rank, world),
               pdist.ring_neighbours(rank, world)))
    finally:
        dist.destroy_process_group()


def test_two_rank_bookkeeping():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n_mine, total, secs, rw, n_paths, slab, ring in res:
        assert n_mine == 6                       # 12 runs over 2 ranks
        assert total == 3000.0 and secs == 2.0   # SUM of units, MAX of time
        assert rw == (rank, 2)
        assert n_paths == 12
        assert slab == (rank * 32, 32)
        assert ring == ((rank - 1) % 2, (rank + 1) % 2)


def test_shard_and_slab_single_process():
    assert pdist.shard_runs(list(range(10)), 1, 4) == [1, 5, 9]
    assert pdist.slab_rows(65536, 7, 8) == (57344, 8192)
    with pytest.raises(ValueError):
        pdist.slab_rows(100, 0, 3)
    assert pdist.reduce_throughput(5.0, 2.0) == (5.0, 2.0)
    assert pdist.rank_world() == (0, 1)


class _FakeSim:
    def __init__(self, rank, rows, ghost, width):
        self.rows, self.ghost_rows = rows, ghost
        h = rows + 2 * ghost
        base = torch.arange(h * width, dtype=torch.float32).reshape(h, width)
        self.energy = base + 1000.0 * rank
        self.strat = torch.full((h, width), 10 + rank, dtype=torch.uint8)
        self.tf = torch.full((h, width), 20 + rank, dtype=torch.uint8)


class _FakeSlab:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.sim = _FakeSim(rank, rows=8, ghost=4, width=16)

    def planes3(self):
        return (self.sim.energy, self.sim.strat, self.sim.tf)


def _comm_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pdabm_b200.decomp import DistComm
        comm = DistComm()
        slab = _FakeSlab(rank, world)
        own_before = slab.sim.energy[4:12].clone()
        comm.exchange([slab])
        up, down = (rank - 1) % world, (rank + 1) % world
        h, w = 16, 16
        base = torch.arange(h * w, dtype=torch.float32).reshape(h, w)
        ok = torch.equal(slab.sim.energy[0:4], (base + 1000.0 * up)[8:12])       # upper slab's last owned rows
        ok &= torch.equal(slab.sim.energy[12:16], (base + 1000.0 * down)[4:8])   # lower slab's first owned rows
        ok &= torch.equal(slab.sim.energy[4:12], own_before)                      # owned rows untouched
        ok &= bool((slab.sim.strat[0:4] == 10 + up).all() and (slab.sim.tf[12:16] == 20 + down).all())
        g = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
        comm.allreduce([g])
        ok &= g.tolist() == [sum(range(1, world + 1)), 10 * sum(range(1, world + 1))]
        out = torch.zeros(world * 3, dtype=torch.int64)
        comm.allgather([out], [torch.tensor([rank, rank, rank], dtype=torch.int64)])
        ok &= out.tolist() == [r for r in range(world) for _ in range(3)]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distcomm_transport_gloo(world):
    """Halo exchange, sum and gather of the decomposed world's transport on CPU tensors (gloo)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res

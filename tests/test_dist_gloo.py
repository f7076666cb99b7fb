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
        counts, n = pdist.allreduce_counts([rank + 1] * 16, 100 * (rank + 1))
        paths = pdist.gather_logs([f"rank{rank}/{r['subdir']}" for r in mine])
        q.put((rank, len(mine), total, secs, counts, n, len(paths), pdist.slab_rows(64, rank, world),
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
    for rank, n_mine, total, secs, counts, n, n_paths, slab, ring in res:
        assert n_mine == 6                       # 12 runs over 2 ranks
        assert total == 3000.0 and secs == 2.0   # SUM of units, MAX of time
        assert counts == [3] * 16 and n == 300
        assert n_paths == 12
        assert slab == (rank * 32, 32)
        assert ring == ((rank - 1) % 2, (rank + 1) % 2)


def test_shard_and_slab_single_process():
    assert pdist.shard_runs(list(range(10)), 1, 4) == [1, 5, 9]
    assert pdist.slab_rows(65536, 7, 8) == (57344, 8192)
    with pytest.raises(ValueError):
        pdist.slab_rows(100, 0, 3)
    assert pdist.reduce_throughput(5.0, 2.0) == (5.0, 2.0)

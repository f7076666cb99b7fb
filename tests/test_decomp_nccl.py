"""The torch.distributed transport of the decomposed world on real GPUs: 2 ranks, NCCL, one slab per
GPU; every rank checks its rows against a whole-world run it performs itself.  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, comm_name):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, DistComm, PeerComm
        kw = dict(width=256, seed=91, env_noise=0.05, mutation_rate=0.01, device=rank)
        ref = Simulation(SimConfig(**kw))
        dec = DecomposedWorld(SimConfig(**kw), PeerComm() if comm_name == "peer" else DistComm())
        ref.init_population()
        dec.init_population()
        rows = slice(rank * (256 // world), (rank + 1) * (256 // world))
        ok = True
        for s in range(70):
            ref.step(1)
            dec.step(1)
            a, b = ref.planes(), dec.planes()
            for k in ("occ", "trait", "strat", "energy"):
                ok &= bool(np.array_equal(a[k][rows], b[k]))
            ca, na = ref.counts()
            cb, nb = dec.counts()
            ok &= (na == nb) and bool(np.array_equal(ca, cb))
            if not ok:
                q.put((rank, False, s))
                return
        q.put((rank, True, 70))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("comm_name", ["nccl", "peer"])
def test_two_gpus_nccl_equal_whole_world(comm_name):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, comm_name)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), res

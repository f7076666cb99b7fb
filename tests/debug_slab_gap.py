"""Developer tool (torchrun, N GPUs): ms per slab step of a continuous run under different host-side settings --
per-kernel profiling on / off, one step() call per step or one call for all steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pdabm_b200 import SimConfig, DecomposedWorld, best_comm

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W = 16384
dw = DecomposedWorld(SimConfig(width=W, height=W * world, seed=12345, device=lr, strategy_per_trait=1), best_comm())
dw.init_population()
dw.step(85, want_counts=False)
one = dw.slabs[0].sim
K = 60


def timed(label, fn):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); fn(); e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / K   # host time to ENQUEUE (the gate read blocks once per step)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label}: {t.item():.4f} ms per step (host loop {host_ms:.4f})", flush=True)


def per_step():
    for _ in range(K):
        dw.step(1, want_counts=False)


for prof in (True, False, True, False):
    one.set_profile(prof)
    timed(f"profile={prof} step(1) x K", per_step)
    timed(f"profile={prof} step(K)     ", lambda: dw.step(K, want_counts=False))
    if prof:
        one.read_profile()
one.set_profile(False)
dist.barrier(); dist.destroy_process_group()

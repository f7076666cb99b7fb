"""Developer tool (torchrun, N GPUs): where one slab step's time goes -- CUDA events between the phases of
DecomposedWorld.step, mean over the timed steps, rank 0 prints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pdabm_b200 import SimConfig, DecomposedWorld, DistComm, best_comm
from pdabm_b200._lib import PD_PART_PRE, PD_PART_CAND, PD_PART_APPLY

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W = 16384
dw = DecomposedWorld(SimConfig(width=W, height=W * world, seed=12345, device=lr, strategy_per_trait=1),
                     DistComm() if os.environ.get("PD_COMM") == "nccl" else best_comm())
dw.init_population()
dw.step(85, want_counts=False)
torch.cuda.synchronize(); dist.barrier()
names = ["exchange", "PRE", "allreduce", "gate(request)", "CAND", "allgather", "APPLY"]
acc = [0.0] * len(names)
K = 30
comm, slabs = dw.comm, dw.slabs
for it in range(K):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    comm.exchange(slabs); ev[1].record()
    slabs[0].part(PD_PART_PRE, 0, False); ev[2].record()
    comm.allreduce([s.wire for s in slabs]); ev[3].record()
    gev = dw._gate_request(); ev[4].record()
    slabs[0].part(PD_PART_CAND, 0, False); ev[5].record()
    comm.allgather([s.cand_all for s in slabs], [s.cand_out for s in slabs]); ev[6].record()
    slabs[0].part(PD_PART_APPLY, 0, False); ev[7].record()
    g = dw._gate_open(gev)
    torch.cuda.synchronize()
    for k in range(len(names)):
        acc[k] += ev[k].elapsed_time(ev[k + 1]) / K
if rank == 0:
    print("gate open:", g, type(comm).__name__)
    print({n: round(a, 4) for n, a in zip(names, acc)}, "sum", round(sum(acc), 4))
dist.barrier(); dist.destroy_process_group()

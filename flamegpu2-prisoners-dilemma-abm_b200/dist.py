"""Multi-GPU plumbing (torch.distributed; one process per GPU).

The path shards in two ways:
  * ensemble: independent runs, run i -> rank i mod world; NO data-path collective, the host only
    gathers per-run logs (CUDAEnsemble, /root/reference/src/model.py:1801-1810, :2188);
  * one world in row slabs: rows [row0, row0+rows) per rank, periodic ring of neighbours.
Only the bookkeeping lives here (model.run_ensemble, decomp.py and bench.py call it); it runs on CPU tensors
with the gloo backend in the tests and on CUDA tensors with NCCL on the GPUs.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def rank_world() -> Tuple[int, int]:
    """(rank, world size) of this process; (0, 1) outside torch.distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_runs(runs: Sequence, rank: int, world: int) -> List:
    """Ensemble partitioning: run i goes to rank i mod world."""
    return [r for i, r in enumerate(runs) if i % world == rank]


def slab_rows(width: int, rank: int, world: int) -> Tuple[int, int]:
    """(row0, rows) of this rank's slab of a width x width world; world must divide width."""
    if width % world:
        raise ValueError(f"world size {world} does not divide the world height {width}")
    rows = width // world
    return rank * rows, rows


def ring_neighbours(rank: int, world: int) -> Tuple[int, int]:
    """(upper, lower) slab owners on the periodic ring: slab 0's upper neighbour is slab world-1."""
    return (rank - 1) % world, (rank + 1) % world


def reduce_throughput(agent_steps: float, seconds: float, device=None) -> Tuple[float, float]:
    """Whole-job numbers: SUM of the units all ranks processed, MAX of their times."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(agent_steps), float(seconds)
    a = torch.tensor([float(agent_steps)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(a.item()), float(t.item())


def gather_logs(paths: List[str]) -> List[str]:
    """Collect the per-run log paths written by every rank on rank 0."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(paths)
    out: List = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(paths))
    return [p for sub in out for p in sub]

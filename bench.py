#!/usr/bin/env python
"""bench.py -- agent-steps/s of the per-step agent pipeline (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload config3|config2|config1] [--presteps P]

A "step" is ONE model step (main layers 1-7, /root/reference/src/model.py:2140-2163) of the
workload's world.  Default workload = BASELINE.json configs[2], the HBM-roofline config:
16384 x 16384 cells (2^28), per-trait strategies, measured in the saturated steady state
(P untimed pre-steps from init, then W warm-up steps, then K timed steps).  With --gpus N
(torchrun) the default is ONE world of 16384 x (16384*N) cells in N row slabs, one per GPU
(2^28 cells per GPU -> weak scaling): one NVLink halo exchange of the state planes' boundary
rows plus a handful of tiny collectives per step (decomp.py).  --partition ensemble instead
gives every rank its own independent 2^28-cell world (no communication on the data path).
value = all ranks' agent-steps / max time.  --workload config5 is BASELINE configs[4]
(65536 x 65536 in 8 slabs of 8192 rows; needs --gpus 8).

Keys of the JSON line beyond the base contract:
  roofline      dominant kernel (k_play_travel): algorithmic bytes per launch / mean device
                time of that launch (CUDA events recorded inside pd_step on the launch stream)
  kernels       the same for every kernel of the step (ms per step, share, GB/s)
  cpu_baseline  the NumPy oracle (a port of the reference's agent functions -- the reference
                itself is CUDA-only) on the host cores, on a bounded sample of the workload
  e2e           the same metric through the host-buffer C-ABI call pd_step_host: every step
                uploads the three state planes from pinned host memory, runs the step and
                downloads planes + counts
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG5 = (65536, dict(strategy_per_trait=1),
           "BASELINE configs[4]: 65536x65536 (2^32 cells) in row slabs, per-trait strategies, steady state")
WORKLOADS = {
    # name: (width, SimConfig overrides, description)
    "config3": (16384, dict(strategy_per_trait=1),
                "BASELINE configs[2]: 16384x16384 (2^28 cells), per-trait strategies, steady state"),
    "config2": (1024, dict(env_noise=0.05, mutation_rate=0.01),
                "BASELINE configs[1]: 1024x1024 (2^20 cells), noise 0.05, mutation 0.01, kin/other, counts every step"),
    "config1": (512, dict(),
                "BASELINE configs[0]: 512x512 (2^18 cells), model.py defaults (kin/other), counts every step"),
}
# algorithmic bytes per cell per launch of each dense sweep (DESIGN.md "Kernels"):
#   play_travel : read tfA 1 + strat 1 + E0 4, write tfB 1 + E1 4
#   claim_punish: read tfB 1 + E1 4, write tfA 1 + E0 4   (+1 strat when counts are wanted)
# The claim commits (move_commit, repro_commit) are sparse: one thread per claimant (~4 % / ~6 % of the cells),
# a handful of random 32-byte sectors each -- their cost is DRAM transactions, not bytes per cell, so they carry
# ncu's measured traffic (profiles/traffic.json) instead of an algorithmic figure.
ALG_BYTES = {"play_travel": 11, "claim_punish": 10}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", f"clocks_{os.getpid()}.csv")

    def start(self):
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU side
def _oracle_worker(args):
    """One oracle world on one host core: returns (agent_steps, seconds) for the timed steps."""
    width, overrides, seed, presteps, steps = args
    from oracle.pd_oracle import Oracle, Params
    kw = {k: v for k, v in overrides.items()}
    orc = Oracle(Params(width=width, seed=seed, **kw))
    orc.init_population()
    for _ in range(presteps):
        orc.step()
    t0 = time.perf_counter()
    agent_steps = 0
    for _ in range(steps):
        agent_steps += orc.n
        orc.step()
    return agent_steps, time.perf_counter() - t0


def cpu_oracle_throughput(overrides, sample_width, presteps, steps, procs):
    """agent-steps/s of the NumPy oracle: `procs` independent sample worlds, one per core."""
    import multiprocessing as mp
    jobs = [(sample_width, overrides, 12345 + i, presteps, steps) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, total, slowest, wall


def run_reference(args, rank, world):
    """--impl reference: the reference has no CPU implementation of this path (FLAMEGPU2 is
    CUDA-only and pyflamegpu is not installable here), so the arm times the oracle PORT of its
    agent functions on all host cores, on a bounded sample of the same workload."""
    if rank != 0:
        return
    width, overrides, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample_w = min(width, 256)
    presteps = min(args.presteps, 70)
    val, total, secs, wall = cpu_oracle_throughput(overrides, sample_w, presteps, max(1, args.steps), cores)
    sample = (f"{cores} independent {sample_w}x{sample_w} worlds (one per host core) with the workload's "
              f"parameters, {presteps} untimed pre-steps then {args.steps} timed steps each")
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": val, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU side
def decomp_parity(world, rank, local_rank, steps=70):
    """Before a slab run is timed: the same decomposition over the SAME transport (torch.distributed / NCCL) on a
    small world, checked bit for bit against one handle that steps the whole world (every rank runs its own).
    70 steps of a 256 x (256 N) world: growth with reproduction retry rounds (global gate), then the cap and the
    cull (global histogram + candidate gather).  Returns the line's `decomp_parity` string; a mismatch is fatal."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, best_comm
    kw = dict(width=256, height=256 * world, seed=77, strategy_per_trait=1, env_noise=0.05, mutation_rate=0.01,
              device=local_rank)
    ref = Simulation(SimConfig(**kw))
    comm = best_comm()
    dec = DecomposedWorld(SimConfig(**kw), comm)
    ref.init_population()
    dec.init_population()
    ok = True
    for s in range(steps):
        ref.step(1)
        dec.step(1)
        if s % 10 == 9 or s == steps - 1:
            a, b = ref.planes(), dec.planes()
            rows = slice(rank * 256, (rank + 1) * 256)
            ok &= all(np.array_equal(a[k][rows], b[k]) for k in ("occ", "trait", "strat", "energy", "tf_raw"))
            ca, na = ref.counts()
            cb, nb = dec.counts()
            ok &= na == nb and bool(np.array_equal(ca, cb))
    retries = dec.retry_rounds_run
    n_end = ref.counts()[1]
    ref.close()
    dec.close()
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t[0]) != 1:
        if rank == 0:
            print(json.dumps({"error": f"decomp_parity FAILED: {world} ranks differ from the single-handle run"}),
                  flush=True)
        sys.exit(3)
    via = "peer-memory halo (symmetric memory over NVLink) + NCCL collectives" if type(comm).__name__ == "PeerComm" else "NCCL"
    return (f"bit-identical to the single-handle run: {world} ranks, {via}, {steps} steps of a 256x{256 * world} "
            f"per-trait world with noise and mutation ({retries} reproduction retry rounds, cap reached: "
            f"{n_end} agents), planes and counts compared every 10 steps")


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pdabm_b200 import Simulation, SimConfig

    from pdabm_b200 import DecomposedWorld, best_comm
    from pdabm_b200.dist import reduce_throughput
    width, overrides, desc = CONFIG5 if args.workload == "config5" else WORKLOADS[args.workload]
    if args.width:
        width = args.width
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    slabs = world > 1 and args.partition == "slabs"
    parity = decomp_parity(world, rank, local_rank) if slabs else None
    if slabs:
        # one world, one row slab per GPU; square for config5, else width x (width * N) so that every
        # GPU owns the same 2^28 cells as the single-GPU run (weak scaling)
        height = args.height or (width if args.workload == "config5" else width * world)
        cfg = SimConfig(width=width, height=height, seed=12345, device=local_rank, **overrides)
        sim = DecomposedWorld(cfg, best_comm())
        one = sim.slabs[0].sim
        cells = width * (height // world)              # owned cells per GPU
        desc += f"; one {width}x{height} world in {world} row slabs"
    else:
        height = args.height or width
        cfg = SimConfig(width=width, height=height, seed=12345 + rank, device=local_rank, sparse_ctas=args.sparse_ctas,
                        **overrides)
        # an explicit stream: worlds of <= 2^22 cells then replay one CUDA graph per step (pd_config.use_graphs)
        run_stream = torch.cuda.Stream()
        torch.cuda.set_stream(run_stream)
        sim = Simulation(cfg, stream=run_stream)
        one = sim
        cells = width * height
        if height != width:
            desc += f"; one {width}x{height} world"
    want_counts = args.workload not in ("config3", "config5")  # these log first/last only

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    t_init0 = time.perf_counter()
    sim.init_population()
    sim.step(args.presteps, want_counts=False)   # untimed: reach the saturated steady state
    sim.step(args.warmup, want_counts=want_counts)
    _, n_before = sim.counts()
    as0 = sim.agent_steps()
    init_s = time.perf_counter() - t_init0

    # ---- timed region: K steps, per-kernel CUDA events recorded inside pd_step
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # Large worlds: the per-kernel events sit inside the timed region (18 event records against ~10 ms of
    # kernels).  Small worlds (graph replay) are timed without them and profiled in a second pass of K steps.
    graphed = (not slabs) and cells <= (1 << 22)
    one.set_profile(not graphed)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    rr0 = sim.retry_rounds_run if slabs else 0
    gr0 = 0 if slabs else one.graph_replays()
    for _ in range(args.steps):
        sim.step(1, want_counts=want_counts)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    retry_rounds = (sim.retry_rounds_run - rr0) if slabs else 0
    replays = 0 if slabs else one.graph_replays() - gr0
    clocks = sampler.stop() if rank == 0 else None
    as1 = sim.agent_steps()
    if graphed:
        one.set_profile(True)
        for _ in range(args.steps):
            sim.step(1, want_counts=want_counts)
    kernel_ms, prof_steps = one.read_profile()   # rank 0's slab (all slabs do the same work)
    one.set_profile(False)
    counts, n_after = sim.counts()
    agent_steps = as1 - as0

    # ---- row slabs: where a step's time goes on this rank -- CUDA events between the phases of DecomposedWorld.step,
    # a few extra (untimed) steps in the saturated state; max over ranks per phase
    slab_phases = None
    if slabs:
        slab_phases = _slab_phase_times(sim, 10)
        tph = torch.tensor(list(slab_phases.values()), dtype=torch.float64, device="cuda")
        dist.all_reduce(tph, op=dist.ReduceOp.MAX)
        slab_phases = {k: float(v) for k, v in zip(slab_phases, tph.tolist())}

    # ---- e2e: the host-buffer C-ABI call, state round trip through pinned host memory each step
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    rows_local = one.energy.shape[0]
    h_e = torch.empty((rows_local, width), dtype=torch.float32).pin_memory()
    h_s = torch.empty((rows_local, width), dtype=torch.uint8).pin_memory()
    h_t = torch.empty((rows_local, width), dtype=torch.uint8).pin_memory()
    h_e.copy_(one.energy)
    h_s.copy_(one.strat)
    h_t.copy_(one.tf)
    torch.cuda.synchronize()

    def host_step():
        if slabs:  # same round trip, spelled out: upload the slab's planes, resync, step, download
            one.energy.copy_(h_e, non_blocking=True)
            one.strat.copy_(h_s, non_blocking=True)
            one.tf.copy_(h_t, non_blocking=True)
            one._check(one.lib.pd_sync_state(one._h, one.stream_ptr))
            sim.step(1, want_counts=True)
            h_e.copy_(one.energy, non_blocking=True)
            h_s.copy_(one.strat, non_blocking=True)
            h_t.copy_(one.tf, non_blocking=True)
            sim.counts()
        else:
            sim.step_host(h_e, h_s, h_t, 1, h_e, h_s, h_t)

    host_step()  # warm-up of the host path
    barrier()
    t0 = time.perf_counter()
    as_a = sim.agent_steps()
    for _ in range(e2e_steps):
        host_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_agents = sim.agent_steps() - as_a
    h2d = rows_local * width * 6
    d2h = rows_local * width * 6 + 17 * 8

    # ---- where the end-to-end time goes: the same round trip spelled out, CUDA events between its three legs
    cur = torch.cuda.current_stream()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    legs = [0.0, 0.0, 0.0]
    nbd = 3
    for _ in range(nbd):
        evs[0].record(cur)
        one.energy.copy_(h_e, non_blocking=True)
        one.strat.copy_(h_s, non_blocking=True)
        one.tf.copy_(h_t, non_blocking=True)
        evs[1].record(cur)
        one._check(one.lib.pd_sync_state(one._h, one.stream_ptr))
        sim.step(1, want_counts=True)
        evs[2].record(cur)
        h_e.copy_(one.energy, non_blocking=True)
        h_s.copy_(one.strat, non_blocking=True)
        h_t.copy_(one.tf, non_blocking=True)
        evs[3].record(cur)
        torch.cuda.synchronize()
        for k in range(3):
            legs[k] += evs[k].elapsed_time(evs[k + 1]) / nbd
    barrier()

    # ---- also (N = 1, default workload): the other BASELINE configurations and the from-init phase, in short
    also = None
    if world == 1 and args.workload == "config3" and not args.no_also:
        also = {}
        # config 3 from init: the growth phase (16 % -> 50 % density in ~60 steps), counts off
        sim.init_population()
        torch.cuda.synchronize()
        as_i = sim.agent_steps()
        ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ei0.record()
        sim.step(60, want_counts=False)
        ei1.record()
        torch.cuda.synchronize()
        msi = ei0.elapsed_time(ei1)
        also["config3_from_init"] = {"workload": desc + " -- the first 60 steps from init (growth phase)", "steps": 60,
                                     "ms_per_step": msi / 60, "value": (sim.agent_steps() - as_i) / (msi / 1e3),
                                     "unit": "agent-steps/s", "agents_at_end": sim.counts()[1]}
        also["config1"] = _small_config_line("config1", 100)
        also["config2"] = _small_config_line("config2", 500)
        also["config4_16runs"] = _ensemble_line(16, 100)

    # ---- also (N = 8, default workload): BASELINE configs[4] itself -- 65536 x 65536 (2^32 cells) in 8 slabs of 8192 rows
    if slabs and world == 8 and args.workload == "config3" and not args.no_also:
        sim.close()
        w5, ov5, d5 = CONFIG5
        big = DecomposedWorld(SimConfig(width=w5, height=w5, seed=12345, device=local_rank, **ov5), best_comm())
        big.init_population()
        big.step(args.presteps, want_counts=False)
        big.step(3, want_counts=False)
        barrier()
        a5 = big.agent_steps()
        e50, e51 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e50.record()
        for _ in range(20):
            big.step(1, want_counts=False)
        e51.record()
        barrier()
        g5, t5 = reduce_throughput(big.agent_steps() - a5, e50.elapsed_time(e51), device="cuda")
        also = {"config5": {"workload": d5 + f"; one {w5}x{w5} world in 8 row slabs (2^29 cells per GPU)", "steps": 20,
                            "ms_per_step": t5 / 20, "value": g5 / (t5 / 1e3),
                            "unit": "agent-steps/s", "cell_steps_per_s": w5 * w5 * 20 / (t5 / 1e3),
                            "agents_at_end": big.counts()[1]}}
        sim = big

    # ---- reduce over ranks: total agent-steps, max time
    if world > 1:
        t = torch.tensor([ms, e2e_s] + legs, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a = torch.tensor([agent_steps, e2e_agents], dtype=torch.float64, device="cuda")
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(t[0]), float(t[1])
        legs = [float(t[2]), float(t[3]), float(t[4])]
        agent_steps, e2e_agents = float(a[0]), float(a[1])
    if rank != 0:
        sim.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    secs = ms / 1e3
    kern = {}
    step_ms_sum = sum(kernel_ms.values()) / max(1, prof_steps)
    for k, v in kernel_ms.items():
        per = v / max(1, prof_steps)
        ent = {"ms_per_step": per, "share": per / step_ms_sum if step_ms_sum else None}
        if k in ALG_BYTES:
            b = ALG_BYTES[k] + (1 if (k == "claim_punish" and want_counts) else 0)
            ent["alg_bytes_per_cell"] = b
            ent["gbs"] = b * cells / (per * 1e-3) / 1e9 if per > 0 else None
            ent["frac_of_hbm_peak"] = ent["gbs"] / peak if per > 0 else None
        kern[k] = ent
    dom = max(ALG_BYTES, key=lambda k: kern[k]["ms_per_step"])
    total_alg = sum(ALG_BYTES.values()) + (1 if want_counts else 0)
    step_gbs = total_alg * cells * args.steps / secs / 1e9

    # ---- CPU baseline (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sw = min(width, 256)
        csteps = 250  # ~10 s of CPU work at 256^2 (the contract asks for a 10-30 s sample)
        cval, ctotal, csecs, _ = cpu_oracle_throughput(overrides, sw, 70, csteps, 1)
        cpu = {"value": cval, "unit": "agent-steps/s", "cores": 1, "kind": "port",
               "sample": f"NumPy oracle (oracle/pd_oracle.py), one {sw}x{sw} world with the workload's parameters, "
                         f"70 untimed pre-steps then {csteps} timed steps ({csecs:.1f} s), single process"}

    line = {
        "metric": "agent-steps/sec", "value": agent_steps / secs, "unit": "agent-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "strong" if (args.height or (slabs and args.workload == "config5")) else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "width": width, "cells": cells * world if slabs else cells, "cells_per_gpu": cells,
                   "partitioning": ("row slabs of one world, one per GPU: halo of 64 boundary rows x 3 planes through "
                                    f"{type(sim.comm).__name__} (PeerComm = packed strips loaded from the neighbours' "
                                    "symmetric memory over NVLink, one pack kernel + one device barrier + one unpack "
                                    "kernel; DistComm = NCCL send/recv), one 8 KB all-reduce (gate words + cull "
                                    "histogram) and one all-gather of cull candidates per step; reproduction retry "
                                    "rounds add one launch + one all-reduce each while the global gate is open" if slabs else
                                    "one world per GPU (ensemble), no data-path collective" if world > 1 else "single world"),
                   "presteps": args.presteps, "agents_at_start": n_before, "agents_at_end": n_after,
                   "density": n_after / (cells * world if slabs else cells), "counts_every_step": want_counts,
                   "l2": "state (11 B/cell of planes + 8 B/cell claim words resident) larger than the 126 MB L2" if cells * 11 > 126e6
                         else "state fits in L2: HBM roofline figures are not meaningful at this size"},
        "cell_steps_per_s": cells * world * args.steps / secs,
        "step_alg_bytes_per_cell": total_alg, "step_alg_gbs": step_gbs, "step_frac_of_hbm_peak": step_gbs / peak,
        "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": kern[dom]["gbs"], "peak": peak,
                     "unit": "GB/s", "frac": kern[dom]["frac_of_hbm_peak"], "traffic": None,
                     "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, STORED in "
                                       "profiles/traffic.json (not measured by this run)",
                     "peak_source": peak_src, "alg_bytes_per_launch": ALG_BYTES[dom] * cells,
                     "launch_ms": kern[dom]["ms_per_step"],
                     # HBM is the roofline this path is held against; what limits the kernel TODAY is named here
                     "limited_by": "instruction issue (ncu, profiles/: issue slots ~75 % busy, the integer ALU pipe "
                                   "~73 % of its cycles, 48 registers -> 5 CTAs per SM, DRAM ~20 % of its peak)"
                                   if (kern[dom]["frac_of_hbm_peak"] or 0) < 0.5 else "HBM bandwidth"},
        "kernels": kern,
        "also": also,
        "decomp_parity": parity,
        "slab_phases_ms": slab_phases,   # (row slabs) ms per phase of a step, max over ranks; their sum is the step
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_agents / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "call": "pd_step_host (host planes in, host planes + counts out)",
                # the legs of one round trip (max over ranks, CUDA events): the PCIe copies are the floor
                "h2d_ms": legs[0], "step_ms": legs[1], "d2h_ms": legs[2],
                "h2d_gbs_per_gpu": h2d / (legs[0] * 1e-3) / 1e9 if legs[0] > 0 else None,
                "d2h_gbs_per_gpu": d2h / (legs[2] * 1e-3) / 1e9 if legs[2] > 0 else None,
                "note": ("the state crosses PCIe twice per step: the two copies are > 90 % of the time at N = 1; with "
                         "N ranks copying at once the per-GPU rates (h2d_gbs_per_gpu / d2h_gbs_per_gpu) drop -- shared "
                         "host memory / PCIe uplinks of the box, a platform limit -- and step_ms then includes waiting "
                         "for the slowest rank's upload at the step's first collective")},
        # per step: 9 kernels (whole world; replayed as one CUDA graph on small worlds), or 12 + 2 per
        # reproduction retry round that actually ran (row slab, this rank)
        "gpu_launches": (12 * args.steps + 2 * retry_rounds) if slabs else 9 * args.steps,
        "graph_replays": replays,
        "clocks": clocks,
        "init_and_presteps_s": init_s,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            line["roofline"]["traffic"] = tr.get("k_" + dom, {}).get(args.workload)
            step_traffic = 0.0
            for k, ent in kern.items():  # stored ncu bytes per launch of every kernel, and their ratio to the
                t_k = tr.get("k_" + k, {}).get(args.workload)  # algorithmic bytes where a per-cell figure exists
                if t_k is not None:
                    step_traffic += t_k
                    ent["traffic_bytes_per_launch_ncu_stored"] = t_k
                    if ent["ms_per_step"] > 0:
                        ent["traffic_gbs"] = t_k / (ent["ms_per_step"] * 1e-3) / 1e9
                        ent["traffic_frac_of_hbm_peak"] = ent["traffic_gbs"] / peak
                    if "alg_bytes_per_cell" in ent:
                        ent["traffic_ratio"] = t_k / (ent["alg_bytes_per_cell"] * cells)
            if step_traffic and args.workload == "config3" and not args.width and not args.height and world == 1:
                # "bytes moved per cell-step over peak bandwidth": the DRAM bytes ncu saw for ALL kernels of one step
                # (stored, same workload) over this run's time per step.  The dense sweeps' algorithmic bytes
                # (step_alg_*) are the floor of this layout; the rest is the claim plane and the sparse commits.
                line["step_traffic_bytes_per_cell_ncu_stored"] = step_traffic / cells
                line["step_traffic_gbs"] = step_traffic / (ms / args.steps * 1e-3) / 1e9
                line["step_traffic_frac_of_hbm_peak"] = line["step_traffic_gbs"] / peak
        except Exception:
            pass
    print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _slab_phase_times(dw, n_steps):
    """ms per phase of one slab step (this rank), mean over n_steps steps with the gate closed: halo exchange, the
    local PRE part, the 8 KB all-reduce, the speculative CAND / all-gather / APPLY.  Events add no synchronisation."""
    import torch
    from pdabm_b200._lib import PD_PART_PRE, PD_PART_CAND, PD_PART_APPLY, PD_PART_DONE
    names = ["halo_exchange", "PRE", "allreduce_gate_hist", "gate_request", "CAND", "allgather_candidates", "APPLY"]
    acc = dict.fromkeys(names, 0.0)
    comm, slabs = dw.comm, dw.slabs
    done = 0
    dw.step(2, want_counts=False)   # the ranks fall into step with each other again (they were read back one by one)
    for _ in range(n_steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        ev[0].record()
        comm.exchange(slabs); ev[1].record()
        for s in slabs:
            s.part(PD_PART_PRE, 0, False)
        ev[2].record()
        comm.allreduce([s.wire for s in slabs]); ev[3].record()
        gev = dw._gate_request(); ev[4].record()
        for s in slabs:
            s.part(PD_PART_CAND, 1, False)
        ev[5].record()
        comm.allgather([s.cand_all for s in slabs], [s.cand_out for s in slabs]); ev[6].record()
        for s in slabs:
            s.part(PD_PART_APPLY, 1, False)
        ev[7].record()
        if dw._gate_open(gev):       # not saturated: finish this step the regular way and stop measuring
            for rnd in range(1, dw.retry_rounds + 1):
                if rnd > 1 and not dw._gate_open():
                    break
                for s in slabs:
                    s.part(1, rnd, False)   # PD_PART_RETRY
                comm.allreduce([s.glb for s in slabs])
            for s in slabs:
                s.part(2, 0, False)         # PD_PART_HIST
            comm.allreduce([s.ghist for s in slabs])
            dw._finish(0, False)
            break
        for s in slabs:
            s.part(PD_PART_DONE)
        torch.cuda.synchronize()
        for k, n in enumerate(names):
            acc[n] += ev[k].elapsed_time(ev[k + 1])
        done += 1
    return {n: (v / done if done else None) for n, v in acc.items()}


def _small_config_line(name, steps, from_init=True):
    """One of BASELINE's small configurations (configs[0] / configs[1]) from init, counts every step, on an
    explicit stream (one CUDA graph replay per step): ms per step and agent-steps/s over `steps` steps."""
    import torch
    from pdabm_b200 import Simulation, SimConfig
    width, overrides, desc = WORKLOADS[name]
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        sim = Simulation(SimConfig(width=width, seed=12345, **overrides), stream=st)
        sim.init_population()
        sim.step(3, want_counts=True)            # graph capture + warm-up, then start again from init
        sim.init_population()
        st.synchronize()
        as0, gr0 = sim.agent_steps(), sim.graph_replays()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        sim.step(steps, want_counts="every")
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1)
        out = {"workload": desc, "steps": steps, "from": "init (16 % density)", "ms_per_step": ms / steps,
               "value": (sim.agent_steps() - as0) / (ms / 1e3), "unit": "agent-steps/s",
               "graph_replays": sim.graph_replays() - gr0, "agents_at_end": sim.counts()[1]}
        sim.close()
    return out


def _ensemble_line(runs, steps):
    """BASELINE configs[3] in small: `runs` concurrent 1024x1024 worlds of the noise x mutation x payoff grid on
    one GPU, one per CUDA stream, from init."""
    import threading
    import torch
    from pdabm_b200 import Simulation, SimConfig
    grid = [(n, m, cc) for n in (0.0, 0.01, 0.05, 0.1) for m in (0.0, 0.001, 0.01, 0.05) for cc in (2.0, 3.0, 4.0, 6.0)]
    sims, streams = [], []
    for i in range(runs):
        noise, mut, cc = grid[(i * 5) % len(grid)]
        stt = torch.cuda.Stream()
        sims.append(Simulation(SimConfig(width=1024, seed=12345 + i, env_noise=noise, mutation_rate=mut, payoff_cc=cc),
                               stream=stt))
        streams.append(stt)

    def all_streams(k):
        th = [threading.Thread(target=lambda s=s: s.step(k, want_counts="every")) for s in sims]
        for t in th:
            t.start()
        for t in th:
            t.join()

    for s in sims:
        s.init_population()
    all_streams(3)
    for s in sims:
        s.init_population()
    torch.cuda.synchronize()
    as0 = sum(s.agent_steps() for s in sims)
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    for stt in streams:
        stt.wait_event(start)
    all_streams(steps)
    ends = []
    for stt in streams:
        e = torch.cuda.Event(enable_timing=True)
        e.record(stt)
        ends.append(e)
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    val = (sum(s.agent_steps() for s in sims) - as0) / (ms / 1e3)
    for s in sims:
        s.close()
    return {"workload": f"BASELINE configs[3] in small: {runs} concurrent 1024x1024 runs (noise x mutation x payoff_cc), "
                        "one per CUDA stream, from init, counts every step",
            "runs": runs, "steps": steps, "ms_per_step_all_runs": ms / steps, "value": val, "unit": "agent-steps/s"}


def run_config4(args, rank, world, local_rank):
    """BASELINE configs[3]: an ensemble sweep of 1024x1024 worlds over a noise x mutation x payoff grid,
    one run per CUDA stream, runs spread over the GPUs (CUDAEnsemble, model.py:1801-1838).  No
    communication between runs; value = all runs' agent-steps / device time of the slowest GPU."""
    import threading
    import torch
    import torch.distributed as dist
    from pdabm_b200 import Simulation, SimConfig
    from pdabm_b200.dist import reduce_throughput
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    grid = [(n, m, cc) for n in (0.0, 0.01, 0.05, 0.1) for m in (0.0, 0.001, 0.01, 0.05) for cc in (2.0, 3.0, 4.0, 6.0)]
    per_gpu = args.runs
    mine = [grid[(rank * per_gpu + i) % len(grid)] for i in range(per_gpu)]
    sims, streams = [], []
    for i, (noise, mut, cc) in enumerate(mine):
        st = torch.cuda.Stream()
        cfg = SimConfig(width=1024, seed=12345 + rank * per_gpu + i, env_noise=noise, mutation_rate=mut,
                        payoff_cc=cc, device=local_rank)
        sims.append(Simulation(cfg, stream=st))
        streams.append(st)

    def drive(sim, k):
        sim.step(k, want_counts=True)

    def all_streams(k):
        th = [threading.Thread(target=drive, args=(s, k)) for s in sims]
        for t in th:
            t.start()
        for t in th:
            t.join()

    for s in sims:
        s.init_population()
    all_streams(args.presteps + args.warmup)
    as0 = sum(s.agent_steps() for s in sims)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    for st in streams:
        st.wait_event(start)
    all_streams(args.steps)
    ends = []
    for st in streams:
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        ends.append(e)
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    agent_steps = sum(s.agent_steps() for s in sims) - as0
    agent_steps, ms = reduce_throughput(agent_steps, ms, device="cuda")  # SUM of the units, MAX of the times
    if rank == 0:
        line = {"metric": "agent-steps/sec", "value": agent_steps / (ms / 1e3), "unit": "agent-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE configs[3]: ensemble sweep of 1024x1024 worlds over noise x mutation x payoff_cc, "
                                       f"{per_gpu} concurrent runs per GPU, one per CUDA stream, counts every step",
                           "runs": per_gpu * world, "presteps": args.presteps,
                           "l2": "each world's state fits in L2: no HBM roofline claim"},
                "gpu_launches": 9 * args.steps * per_gpu * world}
        print(json.dumps(line), flush=True)
    for s in sims:
        s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # ~1 s of timed region at the default workload
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS) + ["config4", "config5"])
    ap.add_argument("--runs", type=int, default=16, help="config4: concurrent runs per GPU")
    ap.add_argument("--partition", default="slabs", choices=["slabs", "ensemble"],
                    help="how N > 1 GPUs are used: row slabs of one world (default) or independent worlds")
    ap.add_argument("--width", type=int, default=0, help="override the workload's width (debug)")
    ap.add_argument("--height", type=int, default=0,
                    help="slabs: fixed world height in rows (strong scaling); default width * N (weak scaling)")
    ap.add_argument("--presteps", type=int, default=80)
    ap.add_argument("--sparse-ctas", type=int, default=0, help="CTAs of the sparse list kernels (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload in ("config4", "config5"):
            args.workload = "config2" if args.workload == "config4" else "config3"
        run_reference(args, rank, world)
    elif args.workload == "config4":
        run_config4(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

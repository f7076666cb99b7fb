#!/usr/bin/env python
"""Statistical comparison harness (north-star criterion 2): per-strategy population
trajectories of a 32-seed ensemble as mean and 95 % confidence band per step, and -- when logs of the
reference model are available (FLAMEGPU2 JSON step logs, one document per line, the format
src/simulation_analysis.R reads) -- the fraction of reference points that fall inside the band.

The reference cannot be run in this environment (pyflamegpu is not installable), so by default the
script only produces the band of THIS implementation; pass --reference-logs DIR to compare.
Test infrastructure (it lives under tests/ because its `--engine oracle` mode imports oracle/).

    python tests/stat_compare.py --width 512 --steps 100 --seeds 32 [--engine cuda|oracle]
                                 [--reference-logs data/] [--out bands.json]
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_ensemble(engine: str, width: int, steps: int, seeds, **kw) -> np.ndarray:
    """-> counts[seed, step, 16]"""
    out = np.zeros((len(seeds), steps + 1, 16), np.int64)
    for si, seed in enumerate(seeds):
        if engine == "cuda":
            from pdabm_b200 import Simulation, SimConfig
            sim = Simulation(SimConfig(width=width, seed=seed, **kw))
            sim.init_population()
            out[si, 0] = sim.counts()[0]
            for t in range(1, steps + 1):
                sim.step(1)
                out[si, t] = sim.counts()[0]
            sim.close()
        else:
            from oracle.pd_oracle import Oracle, Params
            o = Oracle(Params(width=width, seed=seed, **kw))
            o.init_population()
            out[si, 0] = o.population_strat_count
            for t in range(1, steps + 1):
                o.step()
                out[si, t] = o.population_strat_count
    return out


def bands(counts: np.ndarray):
    """mean and 95 % CI of the mean-free spread: [step, bin] arrays (normal approximation of the
    seed-to-seed distribution: mean +- 1.96 sd)."""
    mean = counts.mean(axis=0)
    sd = counts.std(axis=0, ddof=1) if counts.shape[0] > 1 else np.zeros_like(mean)
    return mean, mean - 1.96 * sd, mean + 1.96 * sd


def load_reference_logs(directory: str):
    """FLAMEGPU2 step logs -> list of [step, 16] arrays."""
    runs = []
    for path in sorted(glob.glob(os.path.join(directory, "**", "*.json"), recursive=True)):
        for line in open(path):
            line = line.strip()
            if not line:
                continue
            doc = json.loads(line)
            runs.append(np.array([s["environment"]["population_strat_count"] for s in doc["steps"]], np.int64))
    return runs


def coverage(ref_runs, lo, hi) -> float:
    inside = total = 0
    for r in ref_runs:
        n = min(r.shape[0], lo.shape[0])
        inside += int(((r[:n] >= lo[:n]) & (r[:n] <= hi[:n])).sum())
        total += n * 16
    return inside / max(1, total)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--seeds", type=int, default=32)
    ap.add_argument("--seed0", type=int, default=12345)
    ap.add_argument("--engine", default="cuda", choices=["cuda", "oracle"])
    ap.add_argument("--reference-logs", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    counts = run_ensemble(args.engine, args.width, args.steps, [args.seed0 + i for i in range(args.seeds)])
    mean, lo, hi = bands(counts)
    res = {"engine": args.engine, "width": args.width, "steps": args.steps, "seeds": args.seeds,
           "final_mean": mean[-1].round(1).tolist(), "final_lo": lo[-1].round(1).tolist(),
           "final_hi": hi[-1].round(1).tolist(), "population_final_mean": float(counts[:, -1].sum(axis=1).mean())}
    if args.reference_logs:
        ref = load_reference_logs(args.reference_logs)
        res["reference_runs"] = len(ref)
        res["fraction_inside_95ci"] = coverage(ref, lo, hi) if ref else None
    else:
        res["reference_runs"] = 0
        res["note"] = "reference pyflamegpu logs not available in this environment; band of this implementation only"
    if args.out:
        json.dump({"summary": res, "mean": mean.tolist(), "lo": lo.tolist(), "hi": hi.tolist()}, open(args.out, "w"))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

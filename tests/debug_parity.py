"""Developer tool (not a test): run CUDA and oracle side by side, stop at the first
difference and print the neighbourhood.  Usage: python tests/debug_parity.py '<json kwargs>' steps"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdabm_b200 import Simulation, SimConfig  # noqa: E402
from oracle.pd_oracle import Oracle, Params  # noqa: E402


def main():
    kw = json.loads(sys.argv[1])
    steps = int(sys.argv[2])
    sim = Simulation(SimConfig(collect_stats=True, **kw))
    orc = Oracle(Params(**{k: v for k, v in kw.items() if k != "sparse_ctas"}))
    sim.init_population()
    orc.init_population()
    W = kw["width"]
    for s in range(steps):
        before = orc.planes()
        sim.step(1)
        st = orc.step()
        got, ref = sim.planes(), orc.planes()
        gs = sim.stats()
        bad = np.argwhere((got["occ"] != ref["occ"]) | (got["trait"] != ref["trait"])
                          | (got["energy"] != ref["energy"]) | np.any(got["strat"] != ref["strat"], axis=-1))
        keys = ("play_deaths", "movers", "moved", "travel_deaths", "births", "culled", "living_deaths",
                "agents_end")
        diffstat = {k: (gs[k], st[k]) for k in keys if gs[k] != st[k]}
        if len(bad) or diffstat:
            print(f"step {s}: {len(bad)} cells differ; stat diffs (gpu, oracle): {diffstat}")
            print("gpu stats", gs)
            print("oracle stats", st)
            for (y, x) in bad[:6]:
                print(f"--- cell y={y} x={x}: gpu occ={got['occ'][y, x]} e={got['energy'][y, x]!r} "
                      f"oracle occ={ref['occ'][y, x]} e={ref['energy'][y, x]!r}")
                ys = [(y + d) % W for d in range(-3, 4)]
                xs = [(x + d) % W for d in range(-3, 4)]
                print("before-step energy (0 = empty), rows y-3..y+3:")
                for yy in ys:
                    print("  " + " ".join(f"{before['energy'][yy, xx]:7.2f}" for xx in xs))
                print("after-step gpu energy:")
                for yy in ys:
                    print("  " + " ".join(f"{got['energy'][yy, xx]:7.2f}" for xx in xs))
                print("after-step oracle energy:")
                for yy in ys:
                    print("  " + " ".join(f"{ref['energy'][yy, xx]:7.2f}" for xx in xs))
            return 1
    print("no difference in", steps, "steps")
    return 0


if __name__ == "__main__":
    sys.exit(main())

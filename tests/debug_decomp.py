"""Developer tool: whole world vs LocalComm slabs, print the first difference."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm

kw = json.loads(sys.argv[1]); world = int(sys.argv[2]); steps = int(sys.argv[3]); ghost = int(sys.argv[4]) if len(sys.argv) > 4 else 64
ref = Simulation(SimConfig(collect_stats=True, **kw)); dec = DecomposedWorld(SimConfig(collect_stats=True, **kw), LocalComm(world), ghost_rows=ghost)
ref.init_population(); dec.init_population()
W = kw["width"]
for s in range(1, steps + 1):
    before = ref.planes()
    ref.step(1); dec.step(1)
    a, b = ref.planes(), dec.planes()
    bad = np.argwhere((a["occ"] != b["occ"]) | (a["energy"] != b["energy"]) | (a["trait"] != b["trait"]) | np.any(a["strat"] != b["strat"], -1))
    if len(bad):
        print("step", s, "diff cells", bad[:10].tolist())
        print("ref stats", ref.stats())
        for sl in dec.slabs:
            print("slab", sl.rank, sl.sim.stats())
        for (y, x) in bad[:3]:
            ys = [(y + d) % W for d in range(-3, 4)]; xs = [(x + d) % W for d in range(-3, 4)]
            print(f"cell y={y} x={x} ref e={a['energy'][y,x]!r} occ={a['occ'][y,x]} dec e={b['energy'][y,x]!r} occ={b['occ'][y,x]}")
            for nm, pl in (("before", before), ("ref", a), ("dec", b)):
                print(nm)
                for yy in ys:
                    print("  " + " ".join(f"{pl['energy'][yy, xx]:7.2f}" for xx in xs))
        break
else:
    print("no difference")

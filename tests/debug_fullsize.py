"""Developer tool (1 GPU): whole world vs 4 slabs at 2^28 cells, compared on the device after every step; prints
where the first difference is."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm

W = int(os.environ.get("PD_W", 16384)); NS = int(os.environ.get("PD_SLABS", 4)); STEPS = int(os.environ.get("PD_STEPS", 66))
kw = dict(width=W, seed=12345, strategy_per_trait=1, collect_stats=True)
ref = Simulation(SimConfig(**kw))
# PD_MODE=ww: the "decomposed" side is ONE slab... of the whole height is not a slab; instead compare two slab runs (ss)
dec = DecomposedWorld(SimConfig(**kw), LocalComm(NS))
if os.environ.get("PD_MODE") == "ss":   # slabs against slabs: is the decomposed run deterministic by itself?
    class _Wrap:   # a second decomposed world presented with the whole-world interface used below
        def __init__(self, d): self.d = d
        def init_population(self): self.d.init_population()
        def step(self, n): self.d.step(n)
        def stats(self): return {}
        def counts(self): return self.d.counts()
        @property
        def energy(self): return torch.cat([s.sim.energy[s.sim.ghost_rows:s.sim.ghost_rows + s.sim.rows] for s in self.d.slabs])
        @property
        def strat(self): return torch.cat([s.sim.strat[s.sim.ghost_rows:s.sim.ghost_rows + s.sim.rows] for s in self.d.slabs])
        @property
        def tf(self): return torch.cat([s.sim.tf[s.sim.ghost_rows:s.sim.ghost_rows + s.sim.rows] for s in self.d.slabs])
    ref.close()
    ref = _Wrap(DecomposedWorld(SimConfig(**kw), LocalComm(NS)))
ref.init_population(); dec.init_population()
rows = W // NS
for s in range(STEPS):
    ref.step(1); dec.step(1)
    torch.cuda.synchronize()
    bad = False
    for k, sl in enumerate(dec.slabs):
        g = sl.sim.ghost_rows
        for name in ("energy", "strat", "tf"):
            a = getattr(ref, name)[k * rows:(k + 1) * rows]
            b = getattr(sl.sim, name)[g:g + rows]
            if name == "energy":
                occ = (ref.tf[k * rows:(k + 1) * rows] & 0x80) != 0
                ne = (a != b) & occ
            else:
                ne = a != b
            if bool(ne.any()):
                idx = ne.nonzero()[:8]
                print(f"step {s}: slab {k} plane {name}: {int(ne.sum())} cells differ; first (row in slab, col):", idx.tolist())
                for r, c in idx.tolist():
                    print("   ref", float(ref.energy[k * rows + r, c]), int(ref.tf[k * rows + r, c]), int(ref.strat[k * rows + r, c]),
                          "| slab", float(sl.sim.energy[g + r, c]), int(sl.sim.tf[g + r, c]), int(sl.sim.strat[g + r, c]))
                bad = True
    if bad:
        print("stats ref", ref.stats()); break
else:
    print("no difference in", STEPS, "steps;", ref.counts()[1], "agents")

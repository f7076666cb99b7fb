"""Developer tool (1 GPU): two whole-world runs with the same seed, compared on the device after every step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdabm_b200 import Simulation, SimConfig
W = int(os.environ.get("PD_W", 16384)); STEPS = int(os.environ.get("PD_STEPS", 66))
kw = dict(width=W, seed=12345, strategy_per_trait=1)
a = Simulation(SimConfig(**kw)); b = Simulation(SimConfig(**kw))
a.init_population(); b.init_population()
for s in range(STEPS):
    a.step(1); b.step(1)
    torch.cuda.synchronize()
    occ = (a.tf & 0x80) != 0
    ne = ((a.energy != b.energy) & occ) | (a.tf != b.tf)
    if bool(ne.any()):
        print(f"step {s}: {int(ne.sum())} cells differ between two identical whole-world runs:", ne.nonzero()[:6].tolist()); break
else:
    print("whole world deterministic over", STEPS, "steps")

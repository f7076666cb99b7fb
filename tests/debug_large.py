import sys
sys.path.insert(0, "/root/repo")
from pdabm_b200 import Simulation, SimConfig
for kw in (dict(width=4096, seed=1, collect_stats=True, env_noise=0.05, mutation_rate=0.01),
           dict(width=4096, seed=1, collect_stats=True)):
    sim = Simulation(SimConfig(**kw))
    sim.init_population()
    try:
        for s in range(75):
            sim.step(1)
            c, n = sim.counts()
            if s % 10 == 0: print(kw.get("env_noise"), s, n, flush=True)
        print("ok", kw)
    except Exception as e:
        print("FAIL at step", s, e)
        break
    sim.close()

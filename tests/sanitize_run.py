"""Developer tool: a short run that touches every kernel, meant to be wrapped in compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdabm_b200 import Simulation, SimConfig, DecomposedWorld, LocalComm

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for kw in (dict(width=64, seed=3, init_agent_count=1800, max_agents=2000, reproduce_min_energy=40.0, reproduce_cost=10.0,
                payoff_cc=10.0, payoff_cd=-8.0, env_noise=0.05, mutation_rate=0.05),
           dict(width=128, seed=4, strategy_per_trait=1, env_noise=0.05, init_agent_count=7000, max_agents=8000,
                reproduce_min_energy=45.0, reproduce_cost=10.0, payoff_cc=8.0, payoff_cd=-6.0),
           dict(width=128, seed=5, sparse_ctas=4, init_agent_count=7000, max_agents=8000, reproduce_min_energy=45.0,
                reproduce_cost=10.0, payoff_cc=8.0)):
    sim = Simulation(SimConfig(collect_stats=True, **kw))
    sim.init_population()
    sim.step(steps)
    print(kw["width"], sim.counts()[1], sim.stats())
    sim.close()
dec = DecomposedWorld(SimConfig(width=128, height=256, seed=6, init_agent_count=14000, max_agents=16000,
                                reproduce_min_energy=45.0, reproduce_cost=10.0, payoff_cc=8.0), LocalComm(2))
dec.init_population()
dec.step(max(2, steps // 4))
print("slabs", dec.counts()[1])
dec.close()

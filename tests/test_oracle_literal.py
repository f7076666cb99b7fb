"""The vectorised NumPy oracle against the literal per-agent transcription (oracle/literal.py)
on small worlds: same planes after every step, bit for bit."""
import numpy as np
import pytest

from oracle.pd_oracle import Oracle, Params
from oracle.literal import LiteralModel

CASES = {
    "defaults": dict(width=16, seed=1),
    "defaults32": dict(width=32, seed=2),
    "noise_mut_kin": dict(width=16, seed=3, env_noise=0.1, mutation_rate=0.1),
    "pure": dict(width=16, seed=4, strategy_pure=1, mutation_rate=0.05, env_noise=0.05),
    "per_trait": dict(width=16, seed=5, strategy_per_trait=1, mutation_rate=0.3, env_noise=0.05),
    "inherit": dict(width=16, seed=6, reproduction_inheritence=0.5),
    "harsh": dict(width=16, seed=7, payoff_cd=-30.0, payoff_dd=-5.0, cost_of_living=3.0, travel_cost=4.0,
                  init_agent_count=150),
    "dense": dict(width=16, seed=8, init_agent_count=180, max_agents=220, reproduce_min_energy=40.0,
                  reproduce_cost=10.0, payoff_cc=10.0),
    # reproduction threshold below the cost of living: parents die of the living cost, also right after
    # reproducing in a retry round (the regime of k_repro_commit's deferred-loser list)
    "dying_parents": dict(width=16, seed=10, reproduce_min_energy=1.5, reproduce_cost=0.5, cost_of_living=2.0,
                          init_energy_mu=4.0, init_energy_sigma=2.0, init_energy_min=1.0, init_agent_count=100),
    "dying_parents_noise": dict(width=16, seed=11, reproduce_min_energy=2.0, reproduce_cost=1.0, cost_of_living=2.5,
                                init_energy_mu=5.0, init_energy_sigma=3.0, init_energy_min=1.0, env_noise=0.1,
                                mutation_rate=0.05, strategy_per_trait=1, init_agent_count=80, max_agents=256),
    "tiny_cap": dict(width=16, seed=9, max_agents=60, reproduce_min_energy=52.0, reproduce_cost=5.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_literal(name):
    kw = CASES[name]
    steps = 25 if kw["width"] == 16 else 12
    orc = Oracle(Params(**kw))
    orc.init_population()
    lit = LiteralModel(Params(**kw))
    lit.load(orc)
    for s in range(steps):
        orc.step()
        lit.step()
        a, b = orc.planes(), lit.planes()
        for k in ("occ", "trait", "strat", "energy"):
            assert np.array_equal(a[k], b[k]), f"{name}: step {s}: plane {k} differs"
        assert orc.n == len(lit.agents)

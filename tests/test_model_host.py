"""Host-side mirror of the reference's Python surface: constants, run plan, JSON log."""
import json
import math

import numpy as np

from pdabm_b200 import model
from pdabm_b200.sim import SimConfig, pack_strategies, unpack_strategies


def test_constants_match_reference_defaults():
    """model.py:39-169."""
    assert model.MAX_AGENT_SPACES == 2 ** 18 and model.ENV_MAX == 512
    assert model.INIT_AGENT_COUNT == int(2 ** 18 * 0.16) == 41943
    assert model.AGENT_HARD_LIMIT == 2 ** 17
    assert model.STEP_COUNT == 100 and model.OUTPUT_EVERY_N_STEPS == 1
    assert (model.PAYOFF_CC, model.PAYOFF_DC, model.PAYOFF_CD, model.PAYOFF_DD) == (3.0, 5.0, -1.0, 0.0)
    assert model.AGENT_TRAVEL_COST == 0.5 * model.COST_OF_LIVING
    assert model.POPULATION_COUNT_BINS == 16
    assert not model.AGENT_STRATEGY_PURE and not model.AGENT_STRATEGY_PER_TRAIT
    assert model.MULTI_RUN_STEPS == 10000


def test_sim_config_from_globals():
    cfg = model.sim_config_from_globals(seed=7)
    assert cfg.width == 512 and cfg.seed == 7
    assert cfg.init_agent_count == 41943 and cfg.max_agents == 131072
    assert cfg.strategy_weights == [0.25] * 4
    c = cfg.to_c()
    assert c.width == 512 and c.rows == 512 and c.max_agents == 131072 and c.seed == 7


def test_runplan_matches_reference():
    """configure_runplan, model.py:1813-1838: 2 x 6 parameter points x MULTI_RUN_COUNT runs, seeds
    RANDOM_SEED + i, travel_cost = cost_of_living / 2, sub-directory naming."""
    base = SimConfig(width=64, seed=100)
    runs = model.configure_runplan(base, multi_run_count=3, multi_run_steps=50)
    assert len(runs) == 2 * 6 * 3
    assert [r["config"].seed for r in runs[:3]] == [100, 101, 102]
    costs = [0.1, 0.3, 1, 2 / 3, 1.5, 1.666]
    for k, r in enumerate(runs):
        pure, ci = divmod(k // 3, 6)
        assert r["config"].strategy_pure == pure
        assert r["config"].cost_of_living == costs[ci] and r["config"].travel_cost == costs[ci] / 2
        assert r["subdir"] == "pure%g_env_cost%g_%g_steps" % (pure, costs[ci], 50)
        assert r["steps"] == 50


def test_step_log_shape(tmp_path):
    """One JSON document per line with the keys src/simulation_analysis.R reads (:172-199)."""
    cfg = SimConfig(width=64, seed=5, cost_of_living=0.3, travel_cost=0.15, strategy_pure=1)
    log = model.StepLog(cfg, steps=2)
    log.frame(0, np.arange(16), 120)
    log.frame(1, np.arange(16) * 2, 240)
    path = tmp_path / "sub" / "run.json"
    log.write(str(path))
    lines = open(path).read().splitlines()
    assert len(lines) == 1
    doc = json.loads(lines[0])
    assert doc["config"]["random_seed"] == 5
    env = doc["config"]["environment"]
    assert env["strategy_pure"] == 1 and env["cost_of_living"] == 0.3 and env["travel_cost"] == 0.15
    assert [s["step_index"] for s in doc["steps"]] == [0, 1]
    assert doc["steps"][1]["environment"]["population_strat_count"] == list(range(0, 32, 2))
    assert doc["steps"][1]["agents"]["prisoner"]["default"]["count"] == 240


def test_strategy_packing_roundtrip():
    rng = np.random.default_rng(1)
    s = rng.integers(0, 4, size=(8, 8, 4), dtype=np.uint8)
    p = pack_strategies(s)
    assert p.dtype == np.uint8 and np.array_equal(unpack_strategies(p), s)
    assert pack_strategies(np.array([1, 2, 3, 0])) == 1 | (2 << 2) | (3 << 4)


def test_argv_subset():
    a = model._parse_argv(["model.py", "-s", "10", "-r", "42", "--out-log", "x.json"])
    assert a == {"steps": 10, "seed": 42, "log_file": "x.json"}

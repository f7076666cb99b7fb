"""User-facing surface of the model: the reference's configuration constants, its step loop,
its JSON step log and a CUDAEnsemble-style multi-run driver -- driving libpdabm.so instead of
pyflamegpu.

Everything a user edits in /root/reference/src/model.py:39-169 exists here under the same
name with the same default; ``main()`` follows model.py:1841-2188 (single run or ensemble);
the log has the shape FLAMEGPU2's JSON step logger writes and src/simulation_analysis.R
reads.
"""
from __future__ import annotations

import json
import math
import os
import random
import sys
from time import strftime
from typing import Dict, List, Optional

import numpy as np

from .sim import SimConfig, Simulation

__VERSION__ = "v0.0.5-b200"

##########################################
# SIMULATION CONFIGURATION               #  (model.py:34-169, same names and defaults)
##########################################
RANDOM_SEED: int = random.randint(0, 2 ** 32 // 2 - 1)
MAX_AGENT_SPACES: int = 2 ** 18
INIT_AGENT_COUNT: int = int(MAX_AGENT_SPACES * 0.16)
AGENT_HARD_LIMIT: int = int(MAX_AGENT_SPACES * 0.5)
STEP_COUNT: int = 100
WRITE_LOG: bool = True
LOG_FILE: str = f"data/{strftime('%Y-%m-%d %H-%M-%S')}_{RANDOM_SEED}.json"
VERBOSE_OUTPUT: bool = False
DEBUG_OUTPUT: bool = False
OUTPUT_EVERY_N_STEPS: int = 1
USE_VISUALISATION: bool = False  # the OpenGL visualiser is out of scope
MAX_PLAY_DISTANCE: int = 1
COST_OF_LIVING: float = 1.0
REPRODUCE_MIN_ENERGY: float = 100.0
REPRODUCE_COST: float = 50.0
ALLOW_IMMEDIATE_SPACE_OCCUPATION: bool = True
REPRODUCTION_INHERITENCE: float = 0.0
MAX_CHILDREN_PER_STEP: int = 1
PAYOFF_CC: float = 3.0
PAYOFF_DC: float = 5.0
PAYOFF_CD: float = -1.0
PAYOFF_DD: float = 0.0
AGENT_TRAVEL_STRATEGIES: List[str] = ["random"]
AGENT_TRAVEL_STRATEGY: int = AGENT_TRAVEL_STRATEGIES.index("random")
AGENT_TRAVEL_COST: float = 0.5 * COST_OF_LIVING
MAX_ENERGY: float = 150.0
INIT_ENERGY_MU: float = 50.0
INIT_ENERGY_SIGMA: float = 10.0
INIT_ENERGY_MIN: float = 5.0
ENV_NOISE: float = 0.0
AGENT_STRATEGY_COOP: int = 0
AGENT_STRATEGY_DEFECT: int = 1
AGENT_STRATEGY_TIT_FOR_TAT: int = 2
AGENT_STRATEGY_RANDOM: int = 3
AGENT_STRATEGIES: dict = {
    "always_coop": {"name": "always_coop", "id": AGENT_STRATEGY_COOP, "proportion": 1 / 4},
    "always_defect": {"name": "always_defect", "id": AGENT_STRATEGY_DEFECT, "proportion": 1 / 4},
    "tit_for_tat": {"name": "tit_for_tat", "id": AGENT_STRATEGY_TIT_FOR_TAT, "proportion": 1 / 4},
    "random": {"name": "random", "id": AGENT_STRATEGY_RANDOM, "proportion": 1 / 4},
}
AGENT_TRAIT_COUNT: int = 4
AGENT_STRATEGY_PURE: bool = False
AGENT_STRATEGY_PER_TRAIT: bool = False
AGENT_TRAIT_MUTATION_RATE: float = 0.0
MULTI_RUN = False
MULTI_RUN_STEPS = 10000
MULTI_RUN_COUNT = 1

# derived constants (model.py:191-228)
ENV_MAX: int = math.ceil(math.sqrt(MAX_AGENT_SPACES))
AGENT_WEIGHTS: List[float] = [AGENT_STRATEGIES[s]["proportion"] for s in AGENT_STRATEGIES]
AGENT_STRATEGY_IDS: List[int] = [AGENT_STRATEGIES[s]["id"] for s in AGENT_STRATEGIES]
AGENT_STRATEGY_COUNT: int = len(AGENT_STRATEGY_IDS)
POPULATION_COUNT_BINS: int = AGENT_STRATEGY_COUNT ** 2


def sim_config_from_globals(**overrides) -> SimConfig:
    """Collect the module constants into a SimConfig (what add_env_vars & co. do with
    environment properties, model.py:1678-1743)."""
    g = globals()
    env_max = math.ceil(math.sqrt(g["MAX_AGENT_SPACES"]))
    if AGENT_TRAIT_COUNT != 4 or AGENT_STRATEGY_COUNT != 4:
        raise ValueError("the kernels are specialised for 4 traits x 4 strategies (model.py:126, :152)")
    weights = [0.0] * 4
    for s in g["AGENT_STRATEGIES"].values():
        weights[s["id"]] = float(s["proportion"])
    cfg = dict(
        width=env_max, seed=g["RANDOM_SEED"], init_agent_count=g["INIT_AGENT_COUNT"],
        max_agents=g["AGENT_HARD_LIMIT"], payoff_cc=g["PAYOFF_CC"], payoff_dc=g["PAYOFF_DC"],
        payoff_cd=g["PAYOFF_CD"], payoff_dd=g["PAYOFF_DD"], env_noise=g["ENV_NOISE"],
        cost_of_living=g["COST_OF_LIVING"], travel_cost=g["AGENT_TRAVEL_COST"],
        reproduce_min_energy=g["REPRODUCE_MIN_ENERGY"], reproduce_cost=g["REPRODUCE_COST"],
        reproduction_inheritence=g["REPRODUCTION_INHERITENCE"],
        max_children_per_step=g["MAX_CHILDREN_PER_STEP"], max_energy=g["MAX_ENERGY"],
        init_energy_mu=g["INIT_ENERGY_MU"], init_energy_sigma=g["INIT_ENERGY_SIGMA"],
        init_energy_min=g["INIT_ENERGY_MIN"], mutation_rate=g["AGENT_TRAIT_MUTATION_RATE"],
        strategy_pure=int(g["AGENT_STRATEGY_PURE"]), strategy_per_trait=int(g["AGENT_STRATEGY_PER_TRAIT"]),
        strategy_weights=weights,
    )
    cfg.update(overrides)
    return SimConfig(**cfg)


class StepLog:
    """FLAMEGPU2 JSON step log (configure_logging, model.py:180-187): agent count and the
    ``population_strat_count`` environment array every ``frequency`` steps, frame 0 after init."""

    def __init__(self, cfg: SimConfig, steps: int, frequency: int = 1, environment: Optional[Dict] = None):
        self.frequency = max(1, int(frequency))
        env = {"strategy_pure": int(cfg.strategy_pure), "cost_of_living": cfg.cost_of_living,
               "travel_cost": cfg.travel_cost, "env_noise": cfg.env_noise,
               "mutation_rate": cfg.mutation_rate, "strategy_per_trait": int(cfg.strategy_per_trait),
               "max_agents": cfg.max_agents, "env_max": cfg.width,
               "payoff_cc": cfg.payoff_cc, "payoff_cd": cfg.payoff_cd,
               "payoff_dc": cfg.payoff_dc, "payoff_dd": cfg.payoff_dd}
        if environment:
            env.update(environment)
        self.doc = {"config": {"random_seed": int(cfg.seed), "steps": int(steps), "environment": env},
                    "steps": []}

    def frame(self, step_index: int, counts, n_agents: int):
        self.doc["steps"].append({
            "step_index": int(step_index),
            "environment": {"population_strat_count": [int(c) for c in counts]},
            "agents": {"prisoner": {"default": {"count": int(n_agents)}}},
        })

    def write(self, path: str):
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(path, "w") as f:
            f.write(json.dumps(self.doc))  # one complete JSON document on one line
            f.write("\n")


def simulate(cfg: SimConfig, steps: int, log_file: Optional[str] = None,
             output_every_n_steps: int = 1, stream=None, verbose: bool = False) -> StepLog:
    """The step loop of ``simulation.simulate()`` (model.py:2176): init_fn, then per step the
    pipeline + step_fn + exit_condition_fn (model.py:1579-1588), logging every N steps."""
    if stream is None:  # an explicit stream lets small worlds replay one CUDA graph per step
        import torch
        stream = torch.cuda.Stream(device=cfg.device)
    sim = Simulation(cfg, stream=stream)
    try:
        sim.init_population()
        log = StepLog(cfg, steps, output_every_n_steps)
        counts, n = sim.counts()
        log.frame(0, counts, n)  # frame 0: counts pre-filled by init_fn (model.py:1517-1522)
        # The logged counts stay on the device (pd_set_step_log): up to `cap` frames are queued without a host
        # synchronisation in between, then read back together.
        every = max(1, int(output_every_n_steps))
        cap = max(1, min(1024, -(-steps // every)))
        sim.enable_step_log(cap)
        done, extinct = 0, False
        while done < steps and not extinct:
            queued = 0
            while done < steps and queued < cap:
                if every == 1:
                    chunk = min(cap - queued, steps - done)
                    sim.step(chunk, want_counts="every")
                    queued += chunk
                else:
                    chunk = min(every, steps - done)
                    sim.step(chunk, want_counts=True)
                    queued += 1
                done += chunk
            for row in sim.read_step_log():
                log.frame(int(row[0]), row[2:], int(row[1]))
                if verbose:
                    print(f"step {int(row[0])}: agents {int(row[1])}")
                if row[1] <= 0:  # exit_condition_fn (model.py:1583-1588)
                    extinct = True
                    break
        sim.counts()  # raises if a sparse list overflowed during the run
        if log_file:
            log.write(log_file)
        return log
    finally:
        sim.close()


def configure_runplan(base: SimConfig, multi_run_count: int, multi_run_steps: int) -> List[Dict]:
    """RunPlanVector of model.py:1813-1838: seeds RANDOM_SEED+i, pure in {0,1} x six costs of
    living, travel_cost = cost/2, one output sub-directory per parameter point."""
    runs: List[Dict] = []
    for pure_strategy in [0, 1]:
        for cost_of_living in [0.1, 0.3, 1, 2 / 3, 1.5, 1.666]:
            subdir = "pure%g_env_cost%g_%g_steps" % (pure_strategy, cost_of_living, multi_run_steps)
            for i in range(multi_run_count):
                cfg = SimConfig(**{**base.__dict__, "seed": base.seed + i, "strategy_pure": pure_strategy,
                                   "cost_of_living": cost_of_living, "travel_cost": cost_of_living / 2})
                runs.append({"config": cfg, "steps": multi_run_steps, "subdir": subdir, "index": i})
    return runs


def run_ensemble(runs: List[Dict], out_directory: str = "data", devices: Optional[List[int]] = None,
                 streams_per_device: int = 4, output_every_n_steps: int = 1) -> List[str]:
    """CUDAEnsemble.simulate (model.py:1801-1810, :2188): independent runs, round-robin over
    devices and CUDA streams, one JSON file per run.  No communication between runs.
    Under torch.distributed (one process per GPU) every rank takes the runs i with i mod world == rank on its
    own device (dist.shard_runs) and rank 0 gets the list of all log files (dist.gather_logs)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from .dist import shard_runs, gather_logs, rank_world
    rank, world = rank_world()
    if world > 1:
        runs = shard_runs(runs, rank, world)
        if devices is None:
            devices = [torch.cuda.current_device()]
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    slots = [(d, s) for s in range(streams_per_device) for d in devices]
    streams = {}
    for d, s in slots:
        with torch.cuda.device(d):
            streams[(d, s)] = torch.cuda.Stream(device=d)
    paths: List[str] = []

    def one(job):
        k, run = job
        d, s = slots[k % len(slots)]
        cfg = SimConfig(**{**run["config"].__dict__, "device": d})
        path = os.path.join(out_directory, run["subdir"], f"{run['index']}_{cfg.seed}.json")
        simulate(cfg, run["steps"], log_file=path, output_every_n_steps=output_every_n_steps,
                 stream=streams[(d, s)])
        return path

    with ThreadPoolExecutor(max_workers=len(slots)) as pool:
        for p in pool.map(one, list(enumerate(runs))):
            paths.append(p)
    return gather_logs(paths)


def _parse_argv(argv: List[str]) -> Dict:
    """The FLAMEGPU2 argv subset the reference forwards (model.py:1791, :1808): -s steps, -r seed."""
    out: Dict = {}
    it = iter(argv[1:])
    for a in it:
        if a in ("-s", "--steps"):
            out["steps"] = int(next(it))
        elif a in ("-r", "--random"):
            out["seed"] = int(next(it))
        elif a in ("-d", "--device"):
            out["device"] = int(next(it))
        elif a == "--out-log":
            out["log_file"] = next(it)
    return out


def main(argv: Optional[List[str]] = None):
    """model.py:1841-2188."""
    argv = sys.argv if argv is None else argv
    args = _parse_argv(argv)
    print(f"env_max (grid width): {ENV_MAX}")
    print(f"max agent count: {MAX_AGENT_SPACES}")
    seed = args.get("seed", RANDOM_SEED)
    print(f"random seed: {seed}")
    cfg = sim_config_from_globals(seed=seed, device=args.get("device", 0))
    if not MULTI_RUN:
        print("Running simulation...")
        log = simulate(cfg, args.get("steps", STEP_COUNT),
                       log_file=(args.get("log_file", LOG_FILE) if WRITE_LOG else None),
                       output_every_n_steps=OUTPUT_EVERY_N_STEPS, verbose=VERBOSE_OUTPUT)
        last = log.doc["steps"][-1]
        print(f"done: step {last['step_index']}, agents {last['agents']['prisoner']['default']['count']}")
    else:
        print("Configuring run plan...")
        runs = configure_runplan(cfg, MULTI_RUN_COUNT, args.get("steps", MULTI_RUN_STEPS))
        print("Running simulation...")
        run_ensemble(runs, out_directory="data", output_every_n_steps=OUTPUT_EVERY_N_STEPS)


if __name__ == "__main__":
    main()

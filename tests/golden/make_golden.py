"""Generates tests/golden/trajectories.json from the NumPy oracle (the reference itself cannot be
run or imported here -- see oracle/__init__.py -- so these are regression pins of OUR restatement,
not reference outputs).  Re-run only when the specified semantics change on purpose:

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pd_oracle import Oracle, Params  # noqa: E402

CASES = {
    "defaults_w32": (dict(width=32, seed=12345), 70),
    "defaults_w64": (dict(width=64, seed=12345), 40),
    "config2_like_w64": (dict(width=64, seed=12345, env_noise=0.05, mutation_rate=0.01), 70),
    "per_trait_w64": (dict(width=64, seed=2024, strategy_per_trait=1, env_noise=0.05, mutation_rate=0.01), 70),
    "pure_w32": (dict(width=32, seed=7, strategy_pure=1, mutation_rate=0.02), 70),
    "inherit_w32": (dict(width=32, seed=9, reproduction_inheritence=0.5, env_noise=0.1), 70),
}


def plane_digest(planes) -> str:
    h = hashlib.sha256()
    occ = planes["occ"].astype(np.uint8)
    h.update(occ.tobytes())
    h.update(np.where(occ != 0, planes["trait"], 0).astype(np.uint8).tobytes())
    h.update(np.where(occ[..., None] != 0, planes["strat"], 0).astype(np.uint8).tobytes())
    h.update(np.where(occ != 0, planes["energy"], np.float32(0)).astype("<f4").tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name, (kw, steps) in CASES.items():
        o = Oracle(Params(**kw))
        o.init_population()
        frames = [{"step": 0, "n": o.n, "counts": o.population_strat_count.tolist(), "sha256": plane_digest(o.planes())}]
        for s in range(steps):
            o.step()
            frames.append({"step": s + 1, "n": o.n, "counts": o.population_strat_count.tolist(),
                           "sha256": plane_digest(o.planes())})
        out[name] = {"params": kw, "frames": frames}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectories.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""The statistical-comparison harness (tests/stat_compare.py) on CPU with the oracle as the engine."""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("stat_compare", os.path.join(ROOT, "tests", "stat_compare.py"))
sc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sc)


def test_bands_and_reference_log_coverage(tmp_path):
    counts = sc.run_ensemble("oracle", 32, 10, [1, 2, 3, 4, 5, 6])
    assert counts.shape == (6, 11, 16)
    assert (counts.sum(axis=2)[:, 0] == int(32 * 32 * 0.16)).all()
    mean, lo, hi = sc.bands(counts)
    assert np.all(lo <= mean) and np.all(mean <= hi)
    # a "reference" log in the FLAMEGPU2 step-log shape whose trajectory is one of the ensemble's own runs
    doc = {"config": {"random_seed": 1}, "steps": [
        {"step_index": t, "environment": {"population_strat_count": counts[0, t].tolist()}} for t in range(11)]}
    d = tmp_path / "data" / "sub"
    d.mkdir(parents=True)
    (d / "run.json").write_text(json.dumps(doc) + "\n")
    ref = sc.load_reference_logs(str(tmp_path / "data"))
    assert len(ref) == 1 and ref[0].shape == (11, 16)
    assert sc.coverage(ref, lo, hi) > 0.8

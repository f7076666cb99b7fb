"""bench.py prints ONE JSON line with the keys the driver reads (the contract in the task statement):
the reference arm runs on CPU, our arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--workload", "config1", "--steps", "2", "--warmup", "3", "--presteps", "5"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = _run(["--workload", "config1", "--steps", "5", "--warmup", "3", "--presteps", "5", "--e2e-steps", "2"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "kernels"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["gpu_launches"] == 45 and d["graph_replays"] == 5
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 512 * 512 * 6 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}

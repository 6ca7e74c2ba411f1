"""bench.py contract on the CPU side: the reference arm prints one JSON line with the keys the
driver reads, and the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--workload", "cfg1", "--steps", "2", "--warmup", "1", "--cpu-budget", "5")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "matrix elements/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["rk_integrals_per_s"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "cfg1" in cb["sample"] and cb["value"] == d["value"]
    assert d["steps"] == 2 and d["warmup"] == 1 and cb["omp_threads_set"] == cb["cores"]
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["config"]["workload"].startswith("cfg1: k=8 n_b=96")


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--workload", "cfg1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)

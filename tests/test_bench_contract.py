"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (`--impl reference`: the reference
algorithm on the host cores, one JSON line with the same metric / unit / config as the GPU arm) and the refusal of the
GPU arm to run, or fall back to anything, without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "c1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                        # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("S-matrix frequency-point solves/sec") and d["unit"] == "solves/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "configs[0]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "points per step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_defaults_to_the_headline_workload():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert "configs[4]" in d["config"]["workload"] and d["scaling"] == "strong" and d["dtype"] == "c128"


def test_gpu_arm_refuses_to_run_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = run_bench("--steps", "1", "--warmup", "1", "--no-per-config", "--no-cpu")
    assert out.returncode != 0 and out.stdout.strip() == ""       # no number from a fallback path
    assert "no CPU fallback" in out.stderr

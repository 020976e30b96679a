"""CPU (and one GPU test): the shape of bench.py's reference arm (`--impl reference`: the reference's CPU implementation of the path on host
cores, one JSON line) and of the error line of our arm on a machine without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def _run(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                       env={**os.environ, **(env or {})})
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    return r.returncode, lines


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    rc, lines = _run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1")
    assert rc == 0 and len(lines) == 1                       # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and "synthetic" in d["data"] and "workload" in d["config"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert isinstance(base.get("north_star", ""), str)


def test_reference_arm_non_zero_ranks_exit_quietly():
    rc, lines = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert rc == 0 and lines == []


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return                                                # covered by the GPU run of bench.py itself
    rc, lines = _run("--steps", "1", "--warmup", "1")
    assert rc != 0 and len(lines) == 1 and "error" in json.loads(lines[0])


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_line_carries_every_contract_key():
    rc, lines = _run("--gpus", "1", "--steps", "20", "--warmup", "3")
    assert rc == 0 and len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 20                               # one kernel of ours per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"]
    assert 5000 < d["value"] < 8000 and abs(d["value"] - r["achieved"]) / d["value"] < 0.05
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 ** 31 and e["d2h_bytes_per_step"] == 8 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and isinstance(k["reasons"], list)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"]))

"""bench.py output contract: the one JSON line the driver parses (keys, types, units), for the reference arm on the CPU and for
our arm on the GPU (small batch, few steps: the values are not benchmark numbers, only the shape of the line is checked)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
             "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict, "e2e": dict}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def _check_base(d, steps, warmup):
    for k, t in BASE_KEYS.items():
        assert k in d, k
        assert isinstance(d[k], t) or (t is float and isinstance(d[k], int)), (k, type(d[k]))
    assert "vs_baseline" in d and d["vs_baseline"] is None            # BASELINE.md holds no published number for this metric
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] in base["metric"] and d["unit"] == "ticks/s" and d["higher_is_better"] is True
    assert d["steps"] == steps and d["warmup"] == warmup and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["value"] > 0 and d["e2e"]["value"] > 0


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "2", "--warmup", "3")   # W >= 3 is enforced by bench.py
    _check_base(d, 2, 3)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "instances" in cb["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


@pytest.mark.gpu
def test_our_arm_line(cuda_device):
    d = _run("--steps", "4", "--warmup", "3", "--batch", "4096", "--no-cpu")
    _check_base(d, 4, 3)
    assert "impl" not in d or d["impl"] != "reference"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] >= 4 * 5                                   # at least five kernels per tick
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert {"p50", "p99", "max"} <= set(d["tick_latency_ms"])
    nr = d["next_rows"]
    assert nr["ik_osqp_solved_frac"] == 1.0 and nr["ik_osqp_tick_ms"] > 0 and nr["traj_producer_tick_ms"] > 0

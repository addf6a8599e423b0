"""bench.py's one-line JSON contract: the reference arm (CPU, runs here) and the CUDA arm (B200) on the smallest BASELINE workload (C1)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]                      # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line():
    from oracle import oracle
    if oracle.reference_core() is None:
        pytest.skip("oracle/_ref/libg2o_ref_core.so was not built (no reference tree at build time)")
    d = _run(["--impl", "reference", "--workload", "c1", "--steps", "3", "--warmup", "3"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "lm_iterations_per_second" and d["unit"] == "LM iterations/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert len(d["lm"]["trials"]) == 6 and len(d["lm"]["pcg_iterations"]) == 6     # warm-up + timed iterations, so that both arms' windows can be compared


@pytest.mark.gpu
def test_cuda_arm_line():
    d = _run(["--workload", "c1", "--steps", "3", "--warmup", "3", "--no-cpu"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "lm_iterations_per_second" and d["unit"] == "LM iterations/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    e = d["e2e"]
    assert e["value"] > 0 and e["value"] <= d["value"] * 1.05 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > e["h2d_bytes_per_step"] - 1
    assert d["gpu_launches"] > 0 and d["active_products"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert "workload" in d["config"] and "C1" in d["config"]["workload"]
    assert len(d["lm"]["trials"]) == 6

"""GPU: `bench.py` (our arm) prints ONE JSON line on stdout carrying every key of the driver's contract, measured through
the C ABI (gpu_launches > 0), with a traffic figure only when profiles/gemm_traffic.json belongs to the loaded build."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "64", "--no-side",
           "--no-cpu-baseline", "--settle-s", "0.1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]                     # the contract: one JSON line, nothing else on stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "clip_zero_shot_ad_images_per_s" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] - 64 * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    assert d["gpu_launches"] >= 3 * 60                            # ~66 kernels of ours per step
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None                                  # the committed capture is for batch 512, not 64
    e = d["e2e"]
    assert e["unit"] == "images/s" and e["value"] > 0
    assert e["h2d_bytes_per_step"] == 64 * 224 * 224 * 3 and e["d2h_bytes_per_step"] > 0     # uint8 NHWC feed
    assert abs(e["value"] - d["value"]) / d["value"] < 0.5        # measured, not a copy of `value`; same order of magnitude
    assert e["value"] != d["value"]
    c = d["clocks"]
    assert "sm_mhz" in c and "sm_max_mhz" in c and "reasons" in c

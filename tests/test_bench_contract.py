"""bench.py contract: the JSON line of the reference arm (CPU, runs here) and of the B200 arm (gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "256", "--height", "192"])
    assert d["impl"] == "reference" and COMMON <= set(d)
    assert d["metric"].startswith("georeferenced+resampled") and d["unit"] == "Mpix/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["higher_is_better"] is True


@pytest.mark.gpu
def test_b200_arm_json_line():
    d = _run(["--steps", "4", "--warmup", "3", "--width", "1064", "--height", "708", "--repeats", "2"])
    assert COMMON | {"roofline", "gpu_launches", "clocks", "frames_per_s"} <= set(d)
    assert "impl" not in d and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["gpu_launches"] >= 4 * 4
    r = d["roofline"]
    # the dominant kernel is bound by the FP64 pipe; the HBM view of the same launch rides along
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    h = r["hbm"]
    assert h["unit"] == "GB/s" and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-12
    assert d["parity"]["pass"] is True and d["parity"]["mask_mismatches"] == 0
    # only the image rows that hold georeferenced pixels are uploaded (pipeline sparseUpload)
    full = 1064 * 708 * 3
    assert d["e2e"]["h2d_bytes_full_frame"] == full and 0.4 * full < d["e2e"]["h2d_bytes_per_step"] < 0.8 * full
    assert d["e2e"]["h2d_bytes_per_step"] % (1064 * 3) == 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0
    assert d["value"] > 100 * d["cpu_baseline"]["value"]

"""CPU test of the bench.py contract: the reference arm (the CPU port on the host cores) prints ONE JSON line with the
keys the driver reads; the GPU arm's line is checked for the same keys in the gpu run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _one_json_line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _one_json_line(r.stdout)
    assert d["impl"] == "reference" and KEYS <= set(d)
    assert d["unit"] == "samples/s" and d["value"] > 1e5 and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu",
                        "--neval-per-block", str(1 << 22)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _one_json_line(r.stdout)
    assert KEYS | {"roofline", "clocks"} <= set(d)
    assert d["gpu_launches"] >= 3 and d["value"] > 1e9 and d["e2e"]["value"] > 1e9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and 0 < rf["frac"] < 1.2
    assert abs(d["result"]["integral"] - d["result"]["exact"]) < 7 * d["result"]["sigma"]

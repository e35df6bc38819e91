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
    # the driver marks the arms same_config only when the workload strings are identical: the reference arm names the
    # GPU arm's workload verbatim and says elsewhere that each CPU step is a bounded sample of it
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.Workload("gauss10", 1 << 26).text
    assert d["cpu_baseline"]["single_thread"]["cores"] == 1 and d["cpu_baseline"]["single_thread"]["value"] > 1e5


def test_ncu_numbers_are_tied_to_the_kernel_sources(tmp_path, monkeypatch):
    """roofline.traffic / limiters_ncu come from a committed ncu capture: printed only while the kernel sources hash to
    the capture's csrc_sha256 (measurement hygiene 9)"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench
    from csrc_hash import csrc_sha256
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    wl = bench.Workload("gauss10", 1 << 26)
    (prof / wl.ncu).write_text(json.dumps({"csrc_sha256": csrc_sha256(), "dram_bytes_per_launch": 123, "issue_active_pct": 60.0}))
    t, lim = bench.ncu_capture(wl)
    assert t == 123 and lim["issue_active_pct"] == 60.0
    (prof / wl.ncu).write_text(json.dumps({"csrc_sha256": "0" * 64, "dram_bytes_per_launch": 123}))
    t, lim = bench.ncu_capture(wl)
    assert t is None and "stale" in lim
    os.remove(prof / wl.ncu)
    t, lim = bench.ncu_capture(wl)
    assert t is None and "missing" in lim


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
    # correctness rides in the line: partitioned vs unpartitioned sums of both benched workloads
    assert len(d["parity"]) == 2 and all(p["ok"] and p["maps_identical_across_ranks"] for p in d["parity"])
    assert all(p["max_rel_obs"] <= 1e-10 and p["max_rel_hist"] <= 1e-10 for p in d["parity"])
    # the north star's second workload (example/bubble.jl under vegasmc) in the same line
    b = d["bubble"]
    assert b["unit"] == "steps/s" and b["value"] > 1e8 and b["e2e"]["value"] > 1e8 and 0 < b["roofline"]["frac"] < 1.2
    assert abs(b["result"]["integral"] - b["result"]["exact"]) < 7 * b["result"]["sigma"]

#!/usr/bin/env python
"""profiles/*_ncu.json from an ncu report (run here, no GPU needed):
    python tools/make_ncu_json.py REP.ncu-rep KERNEL_REGEX UNITS_PER_LAUNCH OUT.json [launch-index]
Takes the first (or the given) profiled launch whose kernel name matches and whose duration is above 1 ms (the
fixed-point handles also launch a twin kernel that returns at once), and writes the numbers bench.py prints under
roofline.traffic / roofline.limiters_ncu — stamped with the sha256 of the kernel sources (tools/csrc_hash.py) and the git
commit, so that bench.py can refuse to print them once the sources have changed."""
import csv
import io
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from csrc_hash import csrc_sha256

rep, pat, units, out = sys.argv[1], re.compile(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
want = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units_row = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name, default=None):
    for h, i in col.items():
        if h == name or h.endswith(name):
            try:
                return float(r[i].replace(",", ""))
            except ValueError:
                return default
    return default


def in_unit(r, name, unit_to):
    """value of a metric converted with its unit column (ms / us / ns; Kbyte / Mbyte / byte)"""
    for h, i in col.items():
        if h == name or h.endswith(name):
            v, u = float(r[i].replace(",", "")), units_row[i].lower()
            scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3,
                     "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
            return v * scale
    return None


picked = []
for r in rows[2:]:
    if pat.search(r[col["Kernel Name"]]) and (in_unit(r, "gpu__time_duration.sum", "ms") or 0) > 1.0:
        picked.append(r)
r = picked[want]
inst = num(r, "smsp__inst_executed.sum")
dur = in_unit(r, "gpu__time_duration.sum", "ms")
rd, wr = in_unit(r, "dram__bytes_read.sum", "byte") or 0.0, in_unit(r, "dram__bytes_write.sum", "byte") or 0.0
wave = num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
cyc = num(r, "sm__cycles_elapsed.max")
grid = num(r, "launch__grid_size")
commit = os.environ.get("MCI_GIT_COMMIT") or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                        cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
doc = {
    "kernel": r[col["Kernel Name"]], "units_per_launch": units,
    "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}",
    "csrc_sha256": csrc_sha256(), "git_commit": commit,
    "gpu_time_ms": dur, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "warp_instructions": inst, "warp_instructions_per_sample_warp": inst / (units / 32.0) if inst else None,
    "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
    "threads_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
    "registers_per_thread": num(r, "launch__registers_per_thread"), "grid": grid, "block": num(r, "launch__block_size"),
    "smem_wavefronts": wave, "sm_cycles_elapsed_max": cyc,
    # one wavefront per clock per SM
    "smem_pipe_busy_frac": (wave / (cyc * min(grid or 148, 148))) if wave and cyc else None,
}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc, indent=1))

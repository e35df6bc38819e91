#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed):  python tools/ncu_summary.py rep.ncu-rep [regex ...]
Prints the metrics whose names match any regex (default: a roofline-oriented selection) per profiled launch."""
import csv
import io
import re
import subprocess
import sys

DEFAULT = [r"gpu__time_duration\.sum", r"^launch__(registers_per_thread|occupancy_limit|grid_size|block_size|shared_mem_per_block|waves)",
           r"sm__warps_active\.avg\.pct_of_peak_sustained_active", r"smsp__issue_active\.avg\.pct", r"smsp__inst_executed\.sum$",
           r"smsp__thread_inst_executed\.sum$", r"sm__inst_executed_pipe_[a-z0-9_]+\.sum$",
           r"sm__pipe_[a-z0-9_]+_cycles_active\.avg\.pct_of_peak_sustained_active", r"sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active",
           r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active", r"smsp__average_warp_latency_issue_stalled",
           r"dram__bytes_(read|write)\.sum$", r"dram__throughput\.avg\.pct", r"l1tex__data_bank_conflicts_pipe_lsu\.sum$",
           r"smsp__thread_inst_executed_per_inst_executed", r"sm__throughput\.avg\.pct", r"lts__t_bytes\.sum$",
           r"smsp__inst_executed_op_shared_atom", r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", r"sm__cycles_elapsed\.max"]


def main():
    rep = sys.argv[1]
    pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            name = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
            if any(p.search(h) for p in pats):
                try:
                    if float(v.replace(",", "")) == 0.0:
                        continue
                except ValueError:
                    pass
                print(f"  {name} [{u}] = {v}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Times the fill kernel of one problem under each histogram policy (and with adapt off) — a tuning aid,
not a bench line.  Usage: python tools/sweep_fill.py [case] [solver] [log2 neval_per_block] [walkers] [policies e.g. 1,n] [iterations]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcintegration_jl_b200 as mci
from cases import make_spec

name = sys.argv[1] if len(sys.argv) > 1 else "gauss10"
solver = sys.argv[2] if len(sys.argv) > 2 else "vegas"
lg = int(sys.argv[3]) if len(sys.argv) > 3 else 24
walkers = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
ALL = {"0": (0, True), "1": (1, True), "2": (2, True), "3": (3, True), "4": (4, True), "5": (5, True), "n": (0, False)}
sel = sys.argv[5].split(",") if len(sys.argv) > 5 else list(ALL)
niter = int(sys.argv[6]) if len(sys.argv) > 6 else 8
rows = []
for policy, adapt in (ALL[k] for k in sel):
    kw = dict(n_block=16, neval_per_block=1 << lg, hist_policy=policy)
    if solver != "vegas":
        kw["walkers"] = walkers
    spec, _ = make_spec(name, solver, **kw)
    if not adapt:
        for L in spec.leaves:
            L["adapt"] = False
    eng = mci.CudaEngine(spec)
    for it in range(3):
        eng.run_iteration(it)
        if adapt:
            eng.refine()
    t = []
    for it in range(3, max(4, niter)):
        eng.run_iteration(it)
        t.append(eng.last_timing()[0])
        if adapt:
            eng.refine()
    info = eng.launch_info()
    ms = sorted(t)[len(t) // 2]
    n = spec.n_block * spec.neval_per_block
    rows.append(dict(lib=os.path.basename(os.environ.get("MCIB200_LIB", "libmcib200.so")), case=name, solver=solver, policy=policy, adapt=adapt, hist_mode=info["hist_mode"], grid=info["grid"],
                     threads=info["threads"], smem=info["smem_bytes"], fill_ms=ms, gsamples_per_s=n / ms / 1e6,
                     kernel=info["kernel"]))
    print(json.dumps(rows[-1]), flush=True)
    eng.close()

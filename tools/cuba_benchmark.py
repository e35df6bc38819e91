#!/usr/bin/env python
"""The reference's own timing workload (example/benchmark/cuba/benchmark.jl:139-158): 11 integrands on the unit cube
evaluated together, dof = [[3]] * 11, integrate(...; neval=1e5) = 10 iterations, under :vegas and :vegasmc.  The
reference reports 0.246 s (:vegas) and 0.495 s (:vegasmc) on one CPU thread.  Prints, per solver, the warm wall time of
the whole integrate() call on the GPU (median of 5; the first call of a process also pays ~1 s of NVRTC for the user
source), the per-integral estimates next to the reference's printed ones, and the same at neval = 1e8 (1000x the
evaluations).  One JSON line per (solver, neval)."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcintegration_jl_b200 as mci
from mcintegration_jl_b200.integrands import CUBA11_REFERENCE, cuba11

f = cuba11()
for solver in ("vegas", "vegasmc"):
    for neval in (1e5, 1e8):
        kw = dict(var=None, dof=[[3]] * 11, solver=solver, neval=neval, niter=10)
        t0 = time.perf_counter()
        mci.integrate(f, seed=1, **{**kw, "var": mci.Continuous(0.0, 1.0)})
        first = time.perf_counter() - t0
        ts = []
        for k in range(5):
            t0 = time.perf_counter()
            res = mci.integrate(f, seed=100 + k, **{**kw, "var": mci.Continuous(0.0, 1.0)})
            ts.append(time.perf_counter() - t0)
        ref = CUBA11_REFERENCE[solver]
        dev = [(float(res.mean[k]) - m) / math.hypot(float(res.stdev[k]), e) for k, (m, e) in enumerate(ref)]
        print(json.dumps(dict(workload="example/benchmark/cuba/benchmark.jl: 11 integrands, 3-D, dof=[[3]]*11", solver=solver,
                              neval_per_iteration=neval, niter=10, wall_s_warm_median=sorted(ts)[2], wall_s_first_call=first,
                              reference_wall_s=CUBA11_REFERENCE["wall_s"][solver] if neval == 1e5 else None,
                              reference_hardware="one CPU thread, hardware not stated (example/benchmark/cuba/benchmark.jl:119-120,146-147)",
                              evals_per_s=neval * 10 / sorted(ts)[2], mean=[float(m) for m in res.mean],
                              sigma=[float(s) for s in res.stdev], reference_mean=[m for m, _ in ref],
                              reference_sigma=[e for _, e in ref], deviation_in_combined_sigma=dev)), flush=True)

#!/usr/bin/env python
"""Tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcintegration_jl_b200 as mci
from cases import make_spec

CASES = [("gauss10", "vegas", {}), ("sphere2", "vegas", {}), ("bubble", "vegas", {}), ("genz6_comp", "vegas", dict(hist_policy=2)),
         ("gauss10", "vegas", dict(hist_policy=5)), ("genz6_comp", "vegas", dict(hist_policy=5)), ("complex2", "vegas", {}),
         ("complex2", "vegasmc", dict(walkers=96)), ("complex1", "mcmc", dict(walkers=96)),
         ("bubble", "vegasmc", dict(walkers=96)), ("product4", "mcmc", dict(walkers=96)), ("sphere2_offset", "mcmc", dict(walkers=96)),
         ("discrete2", "vegasmc", dict(walkers=96)), ("bubble_fermik", "mcmc", dict(walkers=96))]
for jit in (2, 1):
    for name, solver, kw in CASES:
        kw = dict(kw)
        npb = 3000 if solver == "vegas" else 96 * 30
        spec, _ = make_spec(name, solver, n_block=4, neval_per_block=npb, **kw)
        spec.jit = jit
        eng = mci.CudaEngine(spec)
        for it in range(5 if kw.get("hist_policy") == 5 else 2):  # the fixed-point cells take over once the histogram has settled
            st = eng.run_iteration(it)
            if solver != "vegas":
                eng.reweight(1.0)
            eng.refine()
        print("ok", name, solver, "jit" if jit == 1 else "aot", eng.kernel_name(), int(st.neval), flush=True)
        eng.close()

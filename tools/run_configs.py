#!/usr/bin/env python
"""Runs the BASELINE.json configurations C2..C5 at their full sizes on one GPU through the public integrate() API and
prints, per config, the estimate, sigma, chi2, deviation from the exact value in sigma, wall time and samples/s.
Evidence for parity at sizes the CPU oracle cannot reach (size-independent property: |mean - exact| <= k sigma, and
sigma shrinking like 1/sqrt(neval)).  Usage: python tools/run_configs.py [scale]  (scale < 1 shrinks every neval)"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import integrands as I
from cases import case

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
PI = math.pi


def run(tag, f, exact, *, solver, neval, niter, measure=None, walkers=None, **cfg):
    t0 = time.perf_counter()
    res = mci.integrate(f, solver=solver, neval=neval * scale, niter=niter, seed=1234, measure=measure, walkers=walkers, **cfg)
    dt = time.perf_counter() - t0
    mean = [np.ravel(m).tolist() for m in res.mean]
    std = [np.ravel(s).tolist() for s in res.stdev]
    dev = None
    if exact is not None:
        dev = [[(m - e) / s if s > 0 else 0.0 for m, s, e in zip(mm, ss, np.ravel(ee))] for mm, ss, ee in zip(mean, std, exact)]
    out = dict(config=tag, solver=solver, neval_per_iter=neval * scale, niter=niter, mean=mean, sigma=std,
               chi2=[np.ravel(c).tolist() for c in res.chi2], exact=[np.ravel(e).tolist() for e in exact] if exact else None,
               dev_sigma=dev, wall_s=dt, timing=res.timing, samples_per_s=neval * scale * niter / dt)
    print(json.dumps(out), flush=True)


para = I.BubblePara()
b = case("bubble")
# CUDA context, library load and the first launch are paid once per process, not per integral: keep them out of C2's time
mci.integrate(I.gauss_nd, solver="vegas", neval=1e4, niter=2, seed=1, var=mci.Continuous(-5.0, 5.0), dof=[[10]])
run("C2 10-D Gaussian, Continuous(-5,5), dof=[[10]]", I.gauss_nd, [[math.erf(5 / math.sqrt(2)) ** 10]], solver="vegas",
    neval=2 ** 30, niter=10, var=mci.Continuous(-5.0, 5.0), dof=[[10]])
run("C3-A 16-D Genz corner peak, pool", I.genz_corner_peak(16, "pool"), [[4.17984992413838e-05]], solver="vegas",
    neval=2 ** 28, niter=20, var=mci.Continuous(0.0, 1.0), dof=[[16]])
run("C3-B 16-D Genz corner peak, 16-leaf composite", I.genz_corner_peak(16, "comp"), [[4.17984992413838e-05]], solver="vegas",
    neval=2 ** 28, niter=20, var=mci.Continuous([(0.0, 1.0)] * 16), dof=[[1]])
run("C4 example/bubble.jl (exact = finite-temperature Lindhard)", b["f"], [[I.lindhard_finite_temperature(para, q) for q in para.extQ]],
    solver="vegasmc", neval=1e8, niter=10, measure="bin_by_discrete", **b["cfg"]())
run("C5 20-D 4-component product", I.product4(), [[1.0]] * 4, solver="mcmc", neval=1e9, niter=10,
    var=mci.Continuous(0.0, 1.0), dof=[[20]] * 4)
run("C5 integrand under vegas", I.product4(), [[1.0]] * 4, solver="vegas", neval=2 ** 28, niter=10,
    var=mci.Continuous(0.0, 1.0), dof=[[20]] * 4)

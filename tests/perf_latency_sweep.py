#!/usr/bin/env python
"""Time to solution of integrate() against neval (BASELINE C2, niter = 10, block = 16): the whole call — engine creation,
10 x (fill, statistics to the host, train!) and the final average — on the GPU, next to the CPU port of the reference
loop (oracle, all host cores) on the same problem.  Shows where the crossover lies for small jobs: the reference's
default is neval = 1e4 (src/main.jl:74).  One JSON line per size.
Lives under tests/ (not collected by pytest) because it runs the oracle as the CPU yardstick and checks the two means
against each other: only tests/, smoke() and bench.py's CPU legs execute anything from oracle/.
    python tests/perf_latency_sweep.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import integrands
from oracle_lib import oracle_integrate

kw = dict(var=None, dof=[[10]], solver="vegas", niter=10, block=16, seed=1234)
mk = lambda: mci.Continuous(-5.0, 5.0)
mci.integrate(integrands.gauss_nd, neval=1e4, **{**kw, "var": mk()})  # library load, CUDA context
for lg in range(4, 10):
    neval = 10 ** lg
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = mci.integrate(integrands.gauss_nd, neval=neval, **{**kw, "var": mk()})
        t.append(time.perf_counter() - t0)
    row = dict(neval_per_iteration=neval, gpu_s=sorted(t)[1], mean=float(res.mean[0]), stdev=float(res.stdev[0]))
    if lg <= 7:
        t0 = time.perf_counter()
        ref = oracle_integrate(integrands.gauss_nd, neval=neval, **{**kw, "var": mk()})
        row.update(cpu_port_s=time.perf_counter() - t0, cpu_threads=min(os.cpu_count() or 1, 16), cpu_mean=float(ref.mean[0]))
    print(json.dumps(row), flush=True)

#!/usr/bin/env python
"""example/bubble.jl of MCIntegration.jl — the polarisation bubble of free electrons at four external momenta — on a
B200, with the built-in device integrand (mcintegration.jl_b200/csrc/mci_integrands.cuh, Integrand<MCI_F_BUBBLE>) and
the bin-by-Discrete measure of bubble.jl:81-84.  Prints the estimate next to the T=0 Lindhard function the reference
compares with (bubble.jl:24-38) and next to the exact finite-temperature value.
    python examples/bubble.py [vegas|vegasmc|mcmc|fermik]
`fermik` is test/bubble_FermiK.jl: the loop momentum is a FermiK{3} variable instead of (R, theta, phi), under :mcmc."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import integrands

para = integrands.BubblePara()                       # rs=1, beta=25, spin=2, me=0.5, Qsize=4 (bubble.jl:10-22)
R = mci.Continuous(0.0, 1.0, alpha=3.0)              # bubble.jl:94-99
theta = mci.Continuous(0.0, math.pi, alpha=3.0)
phi = mci.Continuous(0.0, 2 * math.pi, alpha=3.0)
T = mci.Continuous(0.0, para.β, alpha=3.0)
Ext = mci.Discrete(1, para.Qsize, adapt=False)
solver = sys.argv[1] if len(sys.argv) > 1 else "vegasmc"
if solver == "fermik":                               # bubble_FermiK.jl:93-105
    K = mci.FermiK(3, para.kF, 0.2 * para.kF, 10.0 * para.kF)
    res = mci.integrate(integrands.bubble_fermik(para), measure="bin_by_discrete", var=(T, K, Ext), dof=[[1, 1, 1]],
                        obs=[[0.0] * para.Qsize], solver="mcmc", neval=1e8, niter=10, seed=2024)
else:
    res = mci.integrate(integrands.bubble(para), measure="bin_by_discrete", var=(R, theta, phi, T, Ext), dof=[[1, 1, 1, 1, 1]],
                        obs=[[0.0] * para.Qsize], solver=solver, neval=1e8, niter=10, seed=2024)
print(f"{'q/kF':>8s}  {'avg':>12s}    {'err':>10s}  {'Lindhard T=0':>13s}  {'exact finite T':>14s}")
for i, q in enumerate(para.extQ):
    print(f"{q / para.kF:8.4f}  {res.mean[0][i]:12.8f} ± {res.stdev[0][i]:10.2e}  {para.lindhard(q):13.8f}  "
          f"{integrands.lindhard_finite_temperature(para, q):14.8f}")
mci.report(res.config)

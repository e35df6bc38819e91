#!/usr/bin/env python
"""The README quick start of MCIntegration.jl (README.md:20-27, docs/src/index.md:25-50) on a B200:

    julia> f(x, c) = log(x[1]) / sqrt(x[1])
    julia> integrate(f; var = Continuous(0, 1), neval = 1e5, solver = :vegas)

with the integrand given as CUDA C++ source (compiled by NVRTC into a kernel specialised to the problem)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import integrands

f = integrands.user_integrand("""
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  const double x = X(0, 1);          // x[1] of the first (only) variable
  w[0] = log(x) / sqrt(x);
}""")
for solver in ("vegas", "vegasmc", "mcmc"):
    res = mci.integrate(f, var=mci.Continuous(0.0, 1.0), dof=[[1]], neval=1e7, solver=solver, seed=1)
    print(f"{solver:8s} {res.mean[0]: .6f} +- {res.stdev[0]:.6f}   (exact -4, chi2/dof {res.chi2[0]:.2f})")
mci.report(res)

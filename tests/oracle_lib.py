"""Test-side binding of the CPU oracle (oracle/, test infrastructure).  Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi
from mcintegration_jl_b200.engine import HandleEngine, IterationStats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libmcioracle.so")

_lib = None


def build_oracle():
    src = [os.path.join(ORACLE_DIR, f) for f in ("mci_oracle.c", "mci_oracle.h")] + [
        os.path.join(ROOT, "include", "mcib200.h")]
    if os.path.exists(ORACLE_SO) and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in src):
        return ORACLE_SO
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_SO


def oracle():
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(build_oracle())
    _abi.declare(lib, "orc_")
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    u32p = C.POINTER(C.c_uint32)
    lib.orc_run_iteration.restype = C.c_int
    lib.orc_run_iteration.argtypes = [C.c_void_p, C.c_uint64, C.c_int32]
    lib.orc_packed_ptr.restype = dp
    lib.orc_packed_ptr.argtypes = [C.c_void_p]
    lib.orc_packed_len.restype = C.c_int64
    lib.orc_packed_len.argtypes = [C.c_void_p]
    lib.orc_set_shift_mode.argtypes = [C.c_void_p, C.c_int32]
    lib.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
    lib.orc_uniforms.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32, dp]
    lib.orc_locate.restype = C.c_int32
    lib.orc_locate.argtypes = [dp, C.c_int32, C.c_double]
    lib.orc_smooth.argtypes = [dp, C.c_int32, C.c_double, dp]
    lib.orc_rescale.restype = C.c_int
    lib.orc_rescale.argtypes = [dp, C.c_int32, C.c_double]
    lib.orc_train_continuous.restype = C.c_int
    lib.orc_train_continuous.argtypes = [dp, C.c_int32, dp, C.c_double, dp]
    lib.orc_train_discrete.restype = C.c_int
    lib.orc_train_discrete.argtypes = [dp, C.c_int32, C.c_double, dp, dp]
    lib.orc_map_continuous.argtypes = [dp, C.c_int32, C.c_double, ip, dp, dp]
    lib.orc_maxdof.argtypes = [ip, C.c_int32, C.c_int32, ip]
    lib.orc_mean_std.argtypes = [dp, dp, C.c_int32, C.c_int32, dp, dp]
    lib.orc_average.argtypes = [dp, dp, C.c_int32, C.c_int32, dp, dp, dp]
    lib.orc_do_reweight.argtypes = [dp, dp, C.c_int32, C.c_double, dp]
    lib.orc_probability_identity.argtypes = [dp, dp, dp, C.c_uint64, C.c_int32, dp, dp, dp]
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class OracleEngine(HandleEngine):
    """The oracle behind the same engine interface as CudaEngine (tests only)."""
    prefix = "orc_"

    def __init__(self, spec, nthreads=None, shift_mode=0):
        super().__init__(oracle(), spec)
        self.nthreads = nthreads or min(os.cpu_count() or 1, spec.n_block)
        self.lib.orc_set_shift_mode(self.h, shift_mode)

    def run_local(self, it):
        self._check(self.lib.orc_run_iteration(self.h, int(it), int(self.nthreads)))

    def packed(self):
        n = self.lib.orc_packed_len(self.h)
        return np.ctypeslib.as_array(self.lib.orc_packed_ptr(self.h), shape=(n,))

    def run_iteration(self, it) -> IterationStats:
        self.run_local(it)
        return self.fetch_iteration()


def oracle_integrate(integrand, *, solver="vegas", config=None, neval=1e4, niter=10, block=16, gamma=1.0, adapt=True,
                     reweight_goal=None, ignore=None, measure=None, measurefreq=1, thermal_ratio=0.1, walkers=1,
                     shift_mode=0, nthreads=None, chain_burnin=None, idx=1, **kwargs):
    """integrate() with the oracle as engine: same host driver (mcintegration_jl_b200.main.drive)."""
    from mcintegration_jl_b200.main import _MEASURES, _SOLVERS, _standardize_block, default_chain_burnin, drive
    from mcintegration_jl_b200.integrands import resolve

    f = resolve(integrand)
    if config is None:
        config = mci.Configuration(**kwargs)
    nevalperblock, block = _standardize_block(int(neval), int(block), 1)
    W = 1 if solver == "vegas" else int(walkers)
    nevalperblock = (nevalperblock // W) * W
    spec = mci.problem_from_config(config, f, solver=_SOLVERS[solver], measure_kind=_MEASURES[measure], n_block=block,
                                   neval_per_block=nevalperblock, measurefreq=measurefreq, walkers_per_block=W,
                                   thermal_ratio=thermal_ratio, chain_burnin=default_chain_burnin(W, chain_burnin),
                                   seed=config.next_seed(), mcmc_start_idx=int(idx))  # main.jl:136, like integrate()
    eng = OracleEngine(spec, nthreads=nthreads, shift_mode=shift_mode)
    try:
        eng.set_maps(*mci.maps_from_config(config))
        return drive(eng, config, solver=solver, niter=niter, block=block, gamma=gamma, adapt=adapt,
                     reweight_goal=reweight_goal, ignore=ignore)
    finally:
        eng.close()

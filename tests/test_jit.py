"""NVRTC path (csrc/mci_jit.cu): user integrands given as CUDA C++ source — the drop-in for the reference's "any
closure f(x, c)" (src/main.jl:25-33) — and problem-specialised builds of the built-in integrands (jit=1).
CPU part: the generated program compiles (mci_jit_check needs no GPU) and compiler errors come back readable.
GPU part: a user-source Gaussian reproduces the built-in one; jit=1 kernels pass the same L1 parity against the oracle
as the ahead-of-time kernels."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi, integrands as I
from cases import make_spec

GAUSS_SRC = """
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  double s = 0.0;
  for (int i = 1; i <= 10; ++i) { double x = X(0, i); s += x * x; }
  w[0] = userdata[0] * exp(-0.5 * s);
}
"""
# two integrands on a tuple of variables: f1 = x*y (x in [0,1), y in [0,2)), f2 = k (Discrete 1..4)
MIXED_SRC = """
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  w[0] = X(0, 1) * X(1, 1);
  w[1] = X(2, 1);
}
"""


# test/montecarlo.jl:52-92 (Sphere3): two integrands, obs = [0.0, [0.0, 0.0]], measure doubles the second entry
SPHERE3_F = """
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  const double r2 = X(0, 1) * X(0, 1) + X(0, 2) * X(0, 2);
  w[0] = r2 < 1.0 ? 1.0 : 0.0;
  w[1] = r2 + X(0, 3) * X(0, 3) < 1.0 ? 1.0 : 0.0;
}
"""
SPHERE3_M = """
template <class XA> __device__ void mci_user_measure(const XA &X, double *obs, const double *relw, const double *userdata) {
  obs[0] += relw[0];
  obs[1] += relw[1];
  obs[2] += relw[1] * 2.0;
}
"""
# test/montecarlo.jl:173-185 (TestComplex2): (x, i x^2) as two real integrands + a measure that files re / im parts
COMPLEX2_F = """
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  w[0] = X(0, 1);
  w[1] = X(0, 1) * X(0, 1);
}
"""
COMPLEX2_M = """
template <class XA> __device__ void mci_user_measure(const XA &X, double *obs, const double *relw, const double *userdata) {
  obs[0] += relw[0];   /* Re obs1 */
  obs[3] += relw[1];   /* Im obs2; obs = [re1, im1, re2, im2] */
}
"""


def _check(spec):
    lib = _abi.load_library()
    p, keep = spec.to_ctypes()
    log = C.create_string_buffer(16384)
    rc = lib.mci_jit_check(C.byref(p), log, 16384)
    return rc, log.value.decode()


def _user_gauss_spec(solver=_abi.SOLVER_VEGAS, **kw):
    cfg = mci.Configuration(var=mci.Continuous(-5.0, 5.0), dof=[[10]], seed=1234)
    f = I.user_integrand(GAUSS_SRC, 1, [(2 * math.pi) ** -5])
    return mci.problem_from_config(cfg, f, solver=solver, measure_kind=0, n_block=16, neval_per_block=4096, **kw), cfg


@pytest.mark.parametrize("solver", [_abi.SOLVER_VEGAS, _abi.SOLVER_VEGASMC, _abi.SOLVER_MCMC])
def test_user_source_compiles_for_every_solver(solver):
    spec, _ = _user_gauss_spec(solver, walkers_per_block=64 if solver else 1)
    rc, log = _check(spec)
    assert rc == 0, log


def test_user_source_errors_are_reported():
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[1]], seed=1)
    bad = I.user_integrand("template <class XA> __device__ void mci_user_integrand(const XA &X, const double *u, double *w) "
                           "{ w[0] = no_such_symbol; }")
    spec = mci.problem_from_config(cfg, bad, solver=_abi.SOLVER_VEGAS, measure_kind=0, n_block=16, neval_per_block=64)
    rc, log = _check(spec)
    assert rc == _abi.ERR_UNSUPPORTED
    assert "no_such_symbol" in log and "user_source(1)" in log
    spec.user_source = None  # MCI_F_USER without a source
    rc, log = _check(spec)
    assert rc == _abi.ERR_INVALID and "user_source" in log


@pytest.mark.parametrize("name,solver", [("gauss10", "vegas"), ("bubble", "vegas"), ("bubble", "vegasmc"), ("product4", "mcmc"),
                                          ("discrete2", "vegasmc"), ("genz6_comp", "vegas"), ("sphere2_offset", "mcmc"),
                                          ("bubble_fermik", "mcmc")])
def test_builtin_integrands_compile_specialised(name, solver):
    spec, _ = make_spec(name, solver, walkers=64 if solver != "vegas" else 1)
    spec.jit = 1
    rc, log = _check(spec)
    assert rc == 0, log


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_user_gaussian_matches_builtin():
    spec_u, _ = _user_gauss_spec()
    spec_b, _ = make_spec("gauss10", "vegas", neval_per_block=4096)
    u, b = mci.CudaEngine(spec_u), mci.CudaEngine(spec_b)
    try:
        assert "[jit]" in u.kernel_name() and "MCI_F_GAUSS_ND" in b.kernel_name()
        for it in range(3):
            a, c = u.run_iteration(it), b.run_iteration(it)
            np.testing.assert_allclose(a.block_obs, c.block_obs, rtol=1e-12)  # libdevice exp vs exp_nonpos: <= 2 ulp per weight
            assert np.array_equal(a.block_norm, c.block_norm)
            # the built-in handle switches to fixed-point histogram cells once its histogram has settled (one unit = 2^-58
            # of the previous iteration's largest bin, at most half a unit lost per add: tests/test_gpu_round2.py)
            np.testing.assert_allclose(a.hist, c.hist, rtol=1e-10, atol=c.hist.max() * 2.0 ** -38)
            b.refine()
            u.set_maps(*b.get_maps())
        xu, bu, ju, wu = u.trace_samples(2, 5, 1000, 256)
        xb, bb, jb, wb = b.trace_samples(2, 5, 1000, 256)
        assert np.array_equal(bu, bb) and np.array_equal(xu, xb) and np.array_equal(ju, jb)
        np.testing.assert_allclose(wu, wb, rtol=1e-14)
    finally:
        u.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["vegas", "vegasmc", "mcmc"])
def test_user_integrand_end_to_end(solver):
    """integrate() with a source integrand on a tuple of variables (Continuous, Continuous, Discrete), two integrands:
    int x*y over [0,1)x[0,2) = 1 and sum_{k=1..4} k = 10; the reference's 7 sigma rule."""
    f = I.user_integrand(MIXED_SRC, 2)
    res = mci.integrate(f, solver=solver, neval=4e5, niter=10, seed=5, walkers=64,
                        var=(mci.Continuous(0.0, 1.0), mci.Continuous(0.0, 2.0), mci.Discrete(1, 4)), dof=[[1, 1, 0], [0, 0, 1]])
    for k, ex in enumerate((1.0, 10.0)):
        assert abs(res.mean[k] - ex) < 7 * res.stdev[k] + 1e-12, (k, res.mean[k], res.stdev[k])


def test_user_measure_compiles():
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[2], [3]], obs=[0.0, [0.0, 0.0]], neighbor=[(1, 3), (1, 2)], seed=3)
    f = I.user_integrand(SPHERE3_F, 2)
    for solver in (_abi.SOLVER_VEGAS, _abi.SOLVER_VEGASMC, _abi.SOLVER_MCMC):
        spec = mci.problem_from_config(cfg, f, solver=solver, measure_kind=_abi.MEASURE_USER, n_block=16, neval_per_block=640,
                                       walkers_per_block=64 if solver else 1, measure_source=SPHERE3_M)
        rc, log = _check(spec)
        assert rc == 0, log
    spec = mci.problem_from_config(cfg, f, solver=0, measure_kind=_abi.MEASURE_USER, n_block=16, neval_per_block=640)
    rc, log = _check(spec)  # user measure requested but the source does not define it
    assert rc == _abi.ERR_UNSUPPORTED and "mci_user_measure" in log


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["vegas", "vegasmc", "mcmc"])
def test_sphere3_with_user_measure(solver):
    """test/montecarlo.jl:52-92,272,307,350: check_vector(Sphere3(...), [pi/4, [4pi/3/8, 4pi/3/4]])"""
    res = mci.integrate(I.user_integrand(SPHERE3_F, 2), measure=I.user_measure(SPHERE3_M), solver=solver, neval=4e5, niter=10,
                        seed=8, walkers=64, var=mci.Continuous(0.0, 1.0), dof=[[2], [3]], obs=[0.0, [0.0, 0.0]],
                        neighbor=[(1, 3), (1, 2)])
    want = [[math.pi / 4], [4 * math.pi / 3 / 8, 4 * math.pi / 3 / 4]]
    for k in range(2):
        m, s = np.ravel(res.mean[k]), np.ravel(res.stdev[k])
        for mm, ss, ee in zip(m, s, want[k]):
            assert abs(mm - ee) < 7 * ss, (solver, k, mm, ss, ee)


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["vegas", "vegasmc", "mcmc"])
def test_complex2_as_real_pairs(solver):
    """test/montecarlo.jl:173-185,296,329,377: check_complex(TestComplex2, [0.5, i/3]) with complex numbers carried as
    (re, im) pairs: observables [[re1, im1], [re2, im2]]"""
    res = mci.integrate(I.user_integrand(COMPLEX2_F, 2), measure=I.user_measure(COMPLEX2_M), solver=solver, neval=2e5, niter=10,
                        seed=9, walkers=64, var=mci.Continuous(0.0, 1.0), dof=[[1], [1]], obs=[[0.0, 0.0], [0.0, 0.0]])
    (re1, im1), (re2, im2) = res.mean[0], res.mean[1]
    assert abs(re1 - 0.5) < 7 * res.stdev[0][0] and im1 == 0.0
    assert abs(im2 - 1.0 / 3) < 7 * res.stdev[1][1] and re2 == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,solver", [("gauss10", "vegas"), ("sphere2", "vegas"), ("bubble", "vegas"), ("genz6_comp", "vegas"),
                                          ("discrete2", "vegas"), ("bubble", "vegasmc"), ("sphere2", "vegasmc"),
                                          ("product4_d6", "vegasmc"), ("bubble", "mcmc"), ("sphere2_offset", "mcmc"),
                                          ("product4_d6", "mcmc"), ("lattice_comp", "mcmc"), ("bubble_fermik", "mcmc")])
def test_specialised_kernels_keep_parity(name, solver):
    """jit=1: the same templates compiled against compile-time layout tables must reproduce the oracle like the
    ahead-of-time kernels do (tolerances of tests/test_gpu_parity.py)."""
    from oracle_lib import OracleEngine
    kw = dict(neval_per_block=5000) if solver == "vegas" else dict(n_block=4, neval_per_block=96 * 200, walkers=96)
    spec, _ = make_spec(name, solver, **kw)
    spec.jit = 1
    gpu, cpu = mci.CudaEngine(spec), OracleEngine(spec)
    try:
        assert "[jit]" in gpu.kernel_name()
        for it in range(2):
            a, b = gpu.run_iteration(it), cpu.run_iteration(it)
            tol = 1e-11 if solver == "vegas" else 1e-9
            scale = np.abs(b.block_obs).max() + 1e-300
            np.testing.assert_allclose(a.block_obs, b.block_obs, rtol=tol, atol=tol * scale)
            np.testing.assert_allclose(a.block_norm, b.block_norm, rtol=tol)
            np.testing.assert_allclose(a.hist, b.hist, rtol=1e-9)
            np.testing.assert_allclose(a.visited, b.visited, rtol=1e-9)
            assert a.neval == b.neval
            cpu.refine()
            gpu.set_maps(*cpu.get_maps())
    finally:
        gpu.close(); cpu.close()


# a user integrand on a FermiK{2} momentum: the occupied area of a Fermi disc smeared by a Gaussian (any smooth positive f)
FERMIK_SRC = """
template <class XA> __device__ void mci_user_integrand(const XA &X, const double *userdata, double *w) {
  double kx = X(0, 1), ky = X(1, 1);              // the two Cartesian components of K[1]
  double k = sqrt(kx * kx + ky * ky), d = (k - userdata[0]) / userdata[1];
  w[0] = exp(-0.5 * d * d);
}
"""


def _fermik_user_spec(solver):
    cfg = mci.Configuration(var=mci.FermiK(2, 1.0, 0.2, 10.0), dof=[[1]], seed=5)
    f = I.user_integrand(FERMIK_SRC, 1, [1.0, 0.1])
    return mci.problem_from_config(cfg, f, solver=solver, measure_kind=0, n_block=16, neval_per_block=4096, walkers_per_block=64), cfg


def test_fermik_user_source_compiles_under_mcmc_only():
    """A FermiK variable reaches a user integrand as D component leaves; like in the reference (test/bubble_FermiK.jl:2,
    139-140) only :mcmc takes it"""
    rc, log = _check(_fermik_user_spec(_abi.SOLVER_MCMC)[0])
    assert rc == 0, log
    for solver in (_abi.SOLVER_VEGAS, _abi.SOLVER_VEGASMC):
        rc, log = _check(_fermik_user_spec(solver)[0])
        assert rc == _abi.ERR_UNSUPPORTED and "mcmc only" in log, log


def test_fermik_needs_an_integrand_that_takes_it():
    cfg = mci.Configuration(var=mci.FermiK(3, 1.0, 0.2, 10.0), dof=[[1]], seed=5)
    spec = mci.problem_from_config(cfg, I.discrete_x, solver=_abi.SOLVER_MCMC, measure_kind=0, n_block=16, neval_per_block=64)
    rc, log = _check(spec)
    assert rc == _abi.ERR_INVALID and "FermiK" in log, log


@pytest.mark.gpu
def test_fermik_user_integrand_end_to_end():
    """int d^2k exp(-(|k| - kF)^2 / 2 s^2) = 2 pi (s^2 exp(-kF^2/2s^2) + kF s sqrt(pi/2) (1 + erf(kF / (s sqrt 2)))) on the
    plane, sampled with FermiK{2} proposals only (create!/remove! through the normalisation integrand, shift!)"""
    kF, s = 1.0, 0.1
    exact = 2 * math.pi * (s * s * math.exp(-kF * kF / (2 * s * s)) + kF * s * math.sqrt(math.pi / 2) * (1 + math.erf(kF / (s * math.sqrt(2)))))
    res = mci.integrate(I.user_integrand(FERMIK_SRC, 1, [kF, s]), var=mci.FermiK(2, kF, 0.2, 10.0), dof=[[1]], solver="mcmc",
                        neval=2e6, niter=10, block=16, seed=5)
    assert abs(res.mean[0] - exact) < 6.0 * res.stdev[0], (res.mean[0], res.stdev[0], exact)
    assert res.stdev[0] < 0.02 * exact


def test_disk_cache_round_trip(tmp_path):
    """MCIB200_JIT_CACHE=<dir>: the first process builds with NVRTC and leaves <dir>/mcib200_sm100a_<hash>.cubin, the
    second loads it (same lowered names, same cubin), a truncated file is ignored and rebuilt.  (CPU: compile only.)"""
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    prog = ("import sys, ctypes as C; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from mcintegration_jl_b200 import _abi\nfrom cases import make_spec\n"
            "spec, _ = make_spec('product4', 'mcmc', walkers=64); spec.jit = 1\n"
            "p, keep = spec.to_ctypes(); log = C.create_string_buffer(4096)\n"
            "rc = _abi.load_library().mci_jit_check(C.byref(p), log, 4096); print(rc, log.value.decode()); sys.exit(rc != 0)\n"
            % (os.path.dirname(here), here))
    env = dict(os.environ, MCIB200_JIT_CACHE=str(tmp_path), MCIB200_JIT_DUMP=str(tmp_path / "dump.bin"))

    def run():
        r = subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        return open(tmp_path / "dump.bin", "rb").read()

    first = run()
    files = [f for f in os.listdir(tmp_path) if f.startswith("mcib200_sm100a_") and f.endswith(".cubin")]
    assert len(files) == 1
    path = tmp_path / files[0]
    stored = open(path, "rb").read()
    assert stored.endswith(first) and len(stored) > len(first)  # header (names) + the very cubin NVRTC produced
    assert run() == first                                        # served from the file
    assert open(path, "rb").read() == stored
    open(path, "wb").write(stored[: len(stored) // 2])           # a torn file must not be trusted
    assert len(run()) == len(first)
    assert len(open(path, "rb").read()) == len(stored)           # ... and is replaced by a fresh build

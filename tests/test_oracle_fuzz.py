"""Randomised differential test of the C oracle (oracle/mci_oracle.c) against the independent pure-Python restatement of
the reference's solver loops (tests/golden/make_golden.py: Pool / loop_vegas / loop_vegasmc / loop_mcmc, written from the
Julia sources, sharing no code with the oracle).  Where tests/golden/golden.json pins a handful of hand-made problems,
this walks seeded random ones: grids, number of integrands and their dof (with and without padding between them),
reweight factors, seeds, iteration numbers, block counts, chain lengths, measurement frequency and thermal ratio.
Everything must agree to rounding (1e-12 relative; integer counters exactly)."""
import importlib.util
import os
import random

import numpy as np
import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi, integrands
from mcintegration_jl_b200.configuration import _neighbor
from oracle_lib import OracleEngine

_spec = importlib.util.spec_from_file_location(
    "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)


# ---- the integrands, restated from the reference's tests (1-based slots: x[1], x[2], ...) -------------------------------
def _sphere2(x):  # test/montecarlo.jl:19-24
    r2 = x[1] * x[1] + x[2] * x[2]
    return [1.0 if r2 < 1.0 else 0.0, 1.0 if r2 + x[3] * x[3] < 1.0 else 0.0]


def _poly3(x):  # test/thread.jl
    return [x[1], x[1] * x[1], x[1] * x[1] * x[1]]


def _x2y2(x):  # vegas/montecarlo.jl:69
    return [x[1] * x[1] + x[2] * x[2]]


def _product4(a, d):
    def f(x):
        out = []
        for k in range(4):
            p = 1.0
            for i in range(1, d + 1):
                p *= (1 + a[k] * (x[i] - 0.5))
            out.append(p)
        return out
    return f


def _problem(rnd):
    """-> (device integrand, python integrand, dof rows without the normalisation row, lower, upper)"""
    kind = rnd.choice(["sphere2", "poly3", "x2y2", "product4"])
    if kind == "sphere2":
        return integrands.sphere2, _sphere2, [[2], [3]], 0.0, 1.0
    if kind == "poly3":
        return integrands.poly3, _poly3, [[1], [1], [1]], 0.0, 1.0
    if kind == "x2y2":
        return integrands.x2_plus_y2, _x2y2, [[2]], 0.0, 1.0
    d = rnd.randint(1, 5)
    a = (0.25, 0.5, 0.75, 1.0)
    return integrands.product4(a), _product4(a, d), [[d]] * 4, 0.0, 1.0


def _case(seed):
    rnd = random.Random(seed)
    f_dev, f_py, dof, lo, hi = _problem(rnd)
    nb = rnd.randint(3, 12)
    cuts = sorted(rnd.uniform(lo + 0.02, hi - 0.02) for _ in range(nb - 1))
    grid = [lo] + cuts + [hi]
    Nd = len(dof) + 1
    rw = [rnd.uniform(0.2, 1.0) for _ in range(Nd)]
    tot = sum(rw)
    rw = list(np.array(rw) / tot)
    return dict(f_dev=f_dev, f_py=f_py, dof=dof, grid=grid, rw=rw, seed=rnd.randint(1, 10 ** 6), it=rnd.randint(0, 9),
                n_block=rnd.randint(1, 3), n=rnd.randint(20, 160), mf=rnd.randint(1, 3), thermal=rnd.choice([0.0, 0.1, 0.25]))


def _engine(c, solver, npb):
    cfg = mci.Configuration(var=mci.Continuous(c["grid"][0], c["grid"][-1], grid=c["grid"]), dof=c["dof"], seed=c["seed"])
    spec = mci.problem_from_config(cfg, c["f_dev"], solver=solver, measure_kind=0, n_block=c["n_block"], neval_per_block=npb,
                                   measurefreq=c["mf"], walkers_per_block=1, thermal_ratio=c["thermal"])
    eng = OracleEngine(spec, nthreads=1)
    eng.set_maps(grids=np.array(c["grid"]), reweight=np.array(c["rw"]))
    return eng, cfg


def _restated_problem(c):
    flat = [d[0] for d in c["dof"]] + [0]
    maxdof = max(flat)
    return MG.Prob(lambda: MG.Pool(c["grid"], maxdof), c["f_py"], flat, _neighbor(None, len(flat)))


SEEDS = list(range(1, 13))


@pytest.mark.parametrize("seed", SEEDS)
def test_vegas_random_problems(seed):
    c = _case(1000 + seed)
    want = MG.loop_vegas(_restated_problem(c), c["seed"], c["it"], c["n_block"], c["n"], c["mf"])
    eng, _ = _engine(c, _abi.SOLVER_VEGAS, c["n"])
    try:
        st = eng.run_iteration(c["it"])
        B1 = c["n_block"] + 1
        np.testing.assert_allclose(st.block_obs, want["obs"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"], rtol=1e-9, atol=1e-9)
    finally:
        eng.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_vegasmc_random_problems(seed):
    c = _case(2000 + seed)
    want = MG.loop_vegasmc(_restated_problem(c), c["rw"], c["seed"], c["it"], c["n_block"], c["n"], c["mf"])
    eng, _ = _engine(c, _abi.SOLVER_VEGASMC, c["n"])
    try:
        st = eng.run_iteration(c["it"])
        B1 = c["n_block"] + 1
        np.testing.assert_allclose(st.block_obs, want["obs"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st.visited - B1 * 1e-8, want["visited"], rtol=1e-9, atol=1e-6)
        assert st.neval == want["neval"]
        assert round(float(st.propose[1, 0, 0] - B1 * 1e-8), 3) == want["propose_cv"]
        assert round(float(st.accept[1, 0, 0] - B1 * 1e-10), 3) == want["accept_cv"]
    finally:
        eng.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_mcmc_random_problems(seed):
    c = _case(3000 + seed)
    want = MG.loop_mcmc(_restated_problem(c), c["rw"], c["seed"], c["it"], c["n_block"], c["n"], c["mf"], c["thermal"])
    eng, _ = _engine(c, _abi.SOLVER_MCMC, c["n"])
    try:
        st = eng.run_iteration(c["it"])
        B1 = c["n_block"] + 1
        np.testing.assert_allclose(st.block_obs, want["obs"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st.visited - B1 * 1e-8, want["visited"], rtol=1e-9, atol=1e-6)
        assert st.neval == want["neval"]
        assert np.array_equal(np.round(st.propose - B1 * 1e-8, 3), np.array(want["propose"]))
        assert np.array_equal(np.round(st.accept - B1 * 1e-10, 3), np.array(want["accept"]))
    finally:
        eng.close()

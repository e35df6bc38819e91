"""CPU tests: the oracle reproduces the reference's statistical known answers with the reference's own
rule |mean - expect| < 7 sigma (test/runtests.jl:4-9) for all three solvers, at the reference's sizes
(test/montecarlo.jl:262-387: neval 1e5 x 10 iterations for mcmc / vegasmc, 2e5 x 10 for vegas), with the
reference's decomposition (one chain per block, walkers=1).  This is the strongest pin available for the
solver loops: the reference holds no fixed-seed vectors and Julia is not in the image (oracle/mci_oracle.h)."""
import numpy as np
import pytest

from cases import case
from oracle_lib import oracle_integrate


def check(res, expect, ratio=7.0):
    for o, ex in enumerate(expect):
        m, s = np.ravel(res.mean[o]), np.ravel(res.stdev[o])
        for mm, ss, ee in zip(m, s, np.ravel(ex)):
            assert abs(mm - ee) < ratio * ss, (o, mm, ss, ee)


VEGAS = ["sphere1", "sphere2", "sphere2_offset", "discrete_x", "discrete2", "log_over_sqrt", "lattice_pool", "lattice_comp",
         "lattice_vec", "hypersphere3", "poly3", "x2y2", "gauss4", "product4_d6"]
VEGASMC = ["sphere1", "sphere2", "sphere2_offset", "discrete_x", "discrete2", "log_over_sqrt", "lattice_pool",
           "lattice_comp", "lattice_vec", "hypersphere3", "poly3"]
MCMC = ["sphere1", "sphere2", "sphere2_offset", "discrete_x", "discrete2", "lattice_pool", "lattice_comp", "lattice_vec",
        "poly3"]


@pytest.mark.parametrize("name", VEGAS)
def test_vegas_known_answers(name):
    c = case(name)
    res = oracle_integrate(c["f"], solver="vegas", neval=2e5, niter=10, seed=1234, **c["cfg"]())
    check(res, c["expect"])
    if name == "log_over_sqrt":
        assert res.stdev[0] < 4e-4  # test/montecarlo.jl:317


@pytest.mark.parametrize("name", VEGASMC)
def test_vegasmc_known_answers(name):
    c = case(name)
    res = oracle_integrate(c["f"], solver="vegasmc", neval=1e5, niter=10, seed=1234, **c["cfg"]())
    check(res, c["expect"])
    if name == "log_over_sqrt":
        assert res.stdev[0] < 7e-4  # test/montecarlo.jl:364


@pytest.mark.parametrize("name", MCMC)
def test_mcmc_known_answers(name):
    c = case(name)
    res = oracle_integrate(c["f"], solver="mcmc", neval=1e5, niter=10, seed=1234, **c["cfg"]())
    check(res, c["expect"])


def check_complex(res, expect, ratio=7.0):
    """test/runtests.jl:11-21: real and imaginary parts are held to 7 sigma separately"""
    for o, ex in enumerate(expect):
        m, s = complex(res.mean[o]), complex(res.stdev[o])
        assert abs(m.real - ex.real) < ratio * s.real + 1e-300, (o, m, s, ex)
        assert abs(m.imag - ex.imag) < ratio * s.imag + 1e-300, (o, m, s, ex)


@pytest.mark.parametrize("solver,neval", [("vegas", 2e5), ("vegasmc", 1e5), ("mcmc", 1e5)])
@pytest.mark.parametrize("name", ["complex1", "complex2"])
def test_complex_known_answers(name, solver, neval):
    """test/montecarlo.jl:292-296, 325-330, 370-374: TestComplex1 -> 0.5 + i/3, TestComplex2 -> (0.5, i/3)"""
    c = case(name)
    res = oracle_integrate(c["f"], solver=solver, neval=neval, niter=10, seed=1234, **c["cfg"]())
    check_complex(res, c["expect"])
    assert all(np.iscomplexobj(np.asarray(m)) for m in res.mean)


@pytest.mark.parametrize("solver", ["vegasmc", "mcmc"])
def test_many_short_chains_with_burnin(solver):
    """The GPU decomposition — many short chains per block — on the oracle: 64 chains x 97 measured steps each with
    200 extra burn-in steps (chain_burnin) still reproduces the known answer (sphere volumes)."""
    from mcintegration_jl_b200 import _abi
    from mcintegration_jl_b200.main import drive
    from cases import make_spec
    from oracle_lib import OracleEngine
    spec, cfg = make_spec("sphere2", solver, n_block=16, neval_per_block=64 * 97, walkers=64, chain_burnin=200, seed=77)
    eng = OracleEngine(spec)
    try:
        res = drive(eng, cfg, solver=solver, niter=10, block=16)
    finally:
        eng.close()
    check(res, case("sphere2")["expect"])


def test_stateful_shift_form_agrees():
    """The reference's VEGAS loop updates prob by the ratio form of shift! (sampler.jl:383-384); the device
    evaluates the equivalent create! form 1/(N*dx).  Both oracle modes must agree to rounding."""
    c = case("gauss4")
    a = oracle_integrate(c["f"], solver="vegas", neval=5e4, niter=5, seed=5, shift_mode=0, **c["cfg"]())
    b = oracle_integrate(c["f"], solver="vegas", neval=5e4, niter=5, seed=5, shift_mode=1, **c["cfg"]())
    assert abs(a.mean[0] - b.mean[0]) <= 1e-9 * abs(a.mean[0])
    assert abs(a.stdev[0] - b.stdev[0]) <= 1e-6 * a.stdev[0]


@pytest.mark.parametrize("solver,ratio", [("mcmc", 10.0), ("vegas", 20.0), ("vegasmc", 10.0)])
def test_bubble_vs_lindhard(solver, ratio):
    """test/bubble.jl:93-133 verbatim: train with neval=1e5, block=8, 10 iterations, then ONE warm-restart
    iteration (config=result.config) with neval=1e6, block=64; finite-T polarisation against the T=0 Lindhard
    function with the reference's loose sigma ratios."""
    import mcintegration_jl_b200 as mci
    c = case("bubble")
    para = c["para"]
    cfg = mci.Configuration(seed=3, **c["cfg"]())
    res = oracle_integrate(c["f"], solver=solver, measure="bin_by_discrete", neval=1e5, block=8, config=cfg)
    res = oracle_integrate(c["f"], solver=solver, measure="bin_by_discrete", neval=1e6, block=64, niter=1,
                           config=res.config)
    avg, std = res.mean[0], res.stdev[0]
    for i, q in enumerate(para.extQ):
        assert abs(avg[i] - para.lindhard(q)) < ratio * std[i], (i, avg[i], std[i], para.lindhard(q))

"""CPU tests of the host-side mirror of the reference interface: variables (test/variable.jl),
Configuration normalisation (test/utility.jl:12-28, src/configuration.jl:105-227), block standardisation
(src/main.jl:220-234), marshalling into the POD problem, Result/average/report (src/statistics.jl)."""
import io
import math
from contextlib import redirect_stdout

import numpy as np
import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi, integrands
from mcintegration_jl_b200.configuration import _maxdof, _neighbor
from mcintegration_jl_b200.main import _standardize_block, block_statistics
from mcintegration_jl_b200.variables import is_variable


def test_variables_like_reference():
    """test/variable.jl:1-34"""
    d = mci.Discrete([1, 4])
    assert (d.lower, d.upper) == (1, 4)
    assert not is_variable(int)
    N = 7
    x = mci.Continuous(0.0, 1.0, N, grid=[0.0, 0.1, 0.4, 1.0])
    y = mci.Discrete(1, 6, distribution=np.random.default_rng(0).random(6))
    cxy = mci.CompositeVar(x, y)
    assert is_variable(x) and is_variable(y) and is_variable(cxy)
    assert x.poolsize() == N + 1  # length(x) == N + 1
    assert x.nbins == 3
    x.resize(20)
    assert x.poolsize() == 20
    cxy.resize(30)
    assert cxy.size == 30 and cxy[0].poolsize() == 30 and cxy[1].poolsize() == 30
    assert math.isclose(y.distribution.sum(), 1.0) and y.accumulation[0] == 0.0 and math.isclose(y.accumulation[-1], 1.0)
    with pytest.raises(AssertionError):
        mci.Continuous(1.0, 1.0)  # upper > lower (variable.jl:140)
    with pytest.raises(AssertionError):
        mci.Continuous(0.0, 1.0, 4, offset=3)  # offset + 1 < size (:138)


def test_continuous_defaults():
    """variable.jl:137-153: ninc=1000 grid POINTS -> 999 increments, alpha=2, adapt, size 16 (+1 cache slot)"""
    x = mci.Continuous(-5.0, 5.0)
    assert len(x.grid) == 1000 and x.nbins == 999 and x.alpha == 2.0 and x.adapt and x.size == 17
    assert x.grid[0] == -5.0 and x.grid[-1] == 5.0 and np.all(np.diff(x.grid) > 0)
    comp = mci.Continuous([(0.0, 1.0), (0.0, 2.0)], alpha=3.0)
    assert isinstance(comp, mci.CompositeVar) and len(comp) == 2 and comp[1].upper == 2.0 and comp[0].alpha == 3.0
    dd = mci.Discrete([(1, 3), (1, 4)])
    assert isinstance(dd, mci.CompositeVar) and [l.nbins for l in dd.leaves()] == [3, 4]


def test_maxdof_reference_fixture():
    """test/utility.jl:12-15"""
    assert _maxdof([[1, 2, 3, 5], [3, 1, 2, 7], [2, 4, 1, 2]]) == [3, 4, 3, 7]


def test_configuration_normalisation():
    X, Y, Z = mci.Continuous(0.0, 1.0), mci.Continuous(0.0, 1.0), mci.Continuous(0.0, 1.0)
    c = mci.Configuration(var=[X, Y, Z])  # test/utility.jl:17-27: a vector of variables becomes a tuple
    assert isinstance(c.var, tuple) and c.dof == [[1, 1, 1], [0, 0, 0]] and c.N == 1 and c.norm == 2
    c = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[2], [3]])
    assert c.dof == [[2], [3], [0]] and c.maxdof == [3] and c.N == 2
    assert np.allclose(c.reweight, 1 / 3) and c.normalization == 1e-10 and np.all(c.visited == 1e-8)
    assert c.propose.shape == (3, 3, 3) and c.observable == [0.0, 0.0]
    assert mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=3).dof == [[3], [0]]
    assert mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[2, 3]).dof == [[2], [3], [0]]
    m = np.array([[1, 2], [3, 4]])  # each column is one integral (configuration.jl:139-140)
    assert mci.Configuration(var=(X, Y), dof=m).dof == [[1, 3], [2, 4], [0, 0]]
    c = mci.Configuration(var=(X,), dof=[[1]], obs=[[0.0] * 4])
    assert c.obs_len() == [4]
    with pytest.raises(AssertionError):
        mci.Configuration(var=(X, Y), dof=[[1]])  # wrong dof width
    with pytest.raises(AssertionError):
        mci.Configuration(var=X, dof=[[1]], reweight=[1.0, -1.0])
    with pytest.raises(TypeError):
        mci.Configuration(var=3.0)


def test_pool_auto_resize():
    """configuration.jl:156-160: maxdof + offset >= poolsize - 2 -> resize to maxdof + 2 + offset"""
    v = mci.Continuous(0.0, 1.0)
    mci.Configuration(var=v, dof=[[20]])
    assert v.poolsize() == 22
    v = mci.Continuous(0.0, 1.0)
    mci.Configuration(var=v, dof=[[10]])
    assert v.poolsize() == 17


def test_default_and_custom_neighbors():
    """configuration.jl:201-227"""
    assert _neighbor(None, 5) == [[5, 2], [1, 3], [2, 4], [3], [1]]
    assert _neighbor(None, 2) == [[2], [1]]
    assert _neighbor(None, 3) == [[3, 2], [1], [1]]
    assert _neighbor([(1, 3), (2, 3)], 3) == [[3], [3], [1, 2]]  # test/montecarlo.jl:29
    with pytest.raises(AssertionError):
        _neighbor([(1, 2)], 3)  # not connected


def test_standardize_block():
    """main.jl:220-234"""
    assert _standardize_block(10000, 16, 1) == (625, 16)
    assert _standardize_block(10000, 16, 3) == (10000 // 15, 15)
    assert _standardize_block(10000, 2, 4) == (2500, 4)
    with pytest.raises(AssertionError):
        _standardize_block(8, 16)


def test_marshalling_round_trip():
    para = integrands.BubblePara()
    assert math.isclose(para.kF, 1.9191582926775128, rel_tol=1e-15)
    assert math.isclose(para.β, 6.7876339745650816, rel_tol=1e-14)
    for q, want in zip(para.extQ, (-0.09722569490554467, -0.09517413768359333, -0.08866785115184726, -0.07620341547901538)):
        assert math.isclose(para.lindhard(q), want, rel_tol=1e-9)
    cfg = mci.Configuration(var=(mci.Continuous(0.0, 1.0, alpha=3.0), mci.Discrete(1, 4, adapt=False)), dof=[[1, 1]],
                            obs=[[0.0] * 4], seed=9)
    spec = mci.problem_from_config(cfg, integrands.discrete_x, solver=_abi.SOLVER_VEGASMC, measure_kind=1, n_block=8,
                                   neval_per_block=100)
    p, keep = spec.to_ctypes()
    assert (p.n_topvar, p.n_leaf, p.n_integrand, p.n_block, p.seed) == (2, 2, 1, 8, 9)
    assert [p.leaf[i].kind for i in range(2)] == [0, 1] and p.leaf[0].nbins == 999 and p.leaf[1].nbins == 4
    assert p.leaf[0].alpha == 3.0 and p.leaf[1].adapt == 0 and p.leaf[1].lower_i == 1
    assert [p.dof[i] for i in range(2)] == [1, 1] and p.obs_len[0] == 4
    assert [p.neighbor_ptr[i] for i in range(3)] == [0, 1, 2] and [p.neighbor_idx[i] for i in range(2)] == [1, 0]
    g, d, a, r = mci.maps_from_config(cfg)
    assert len(g) == 1000 and np.allclose(d, 0.25) and np.allclose(a, [0, .25, .5, .75, 1]) and np.allclose(r, 0.5)
    g2 = g.copy(); g2[1:-1] += 1e-4
    mci.maps_to_config(cfg, g2, d, a, np.array([0.3, 0.7]))
    assert np.array_equal(cfg.var[0].grid, g2) and np.allclose(cfg.reweight, [0.3, 0.7])


def test_python_callable_is_refused_with_guidance():
    with pytest.raises(TypeError) as e:
        mci.integrate(lambda x, c: x[0], solver="vegas")
    assert "GPU" in str(e.value)
    with pytest.raises(ValueError):
        mci.integrate(integrands.gauss_nd, solver="MC")  # main.jl:263 "Solver MC is not supported!"


def test_genz_exact_value():
    """c_i = 1.85 i / 136, exact 4.17984992413838e-05"""
    f = integrands.genz_corner_peak(16)
    assert math.isclose(sum(f.userdata), 1.85, rel_tol=1e-14)
    g6 = integrands.genz_corner_peak(6, "comp")
    assert 0 < integrands.genz_corner_peak_exact(g6.userdata) < 1


def test_result_average_and_report():
    """statistics.jl:16-63, 137-220 incl. the `max <= init` quirk (Appendix B.12)"""
    from mcintegration_jl_b200.engine import IterationStats
    rng = np.random.default_rng(1)
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[1], [1]], obs=[0.0, [0.0, 0.0]])
    hist = []
    for it in range(6):
        obs = rng.normal(1.0, 0.01, size=(16, 3)) * 100.0
        st = IterationStats(block_obs=obs, block_norm=np.full(16, 100.0), hist=np.zeros(1), visited=np.zeros(3),
                            propose=np.zeros((3, 3, 3)), accept=np.zeros((3, 3, 3)), neval=1600)
        mean, std = block_statistics(st, cfg, 16)
        assert isinstance(mean[0], float) and mean[1].shape == (2,)
        hist.append((mean, std, {}))
    res = mci.Result.from_history(hist, 1, cfg, 9600)
    assert abs(res.mean[0] - 1.0) < 7 * res.stdev[0] and res.mean[1].shape == (2,)
    m, s, chi2 = res[0]
    assert chi2 >= 0
    one = mci.Result.from_history(hist[:1], 1, cfg, 1600)  # niter=1 with ignore=1 returns iteration 1
    assert one.mean[0] == hist[0][0][0]
    buf = io.StringIO()
    with redirect_stdout(buf):
        mci.report(res)
    assert "Integral 1" in buf.getvalue() or "integral" in buf.getvalue().lower()


def test_report_config_tables():
    """report(config) (configuration.jl:345-464) on counters shaped like the kernels' output"""
    cfg = mci.Configuration(var=(mci.Continuous(0.0, 1.0), mci.Discrete(1, 4)), dof=[[1, 1], [2, 0]])
    cfg.neval = 1000
    cfg.propose = np.full((3, 3, 3), 10.0)
    cfg.accept = np.full((3, 3, 3), 4.0)
    cfg.visited = np.array([300.0, 300.0, 400.0])
    buf = io.StringIO()
    with redirect_stdout(buf):
        mci.report(cfg)
    out = buf.getvalue()
    assert "ChangeIntegrand" in out and "SwapVariable" in out and "Norm ->" in out and "0.400000" in out
    assert "Integrand evaluation = 1000" in out and "Discrete" in out


def test_successive_calls_draw_fresh_samples():
    """main.jl:136: every integrate() call seeds its solver streams with rand(config.rng, 1:10^6), so two calls on one
    Configuration — a warm restart, or a plain repeat — use independent samples.  (The host driver is shared with the
    oracle engine, which stands in for the GPU here.)"""
    from oracle_lib import oracle_integrate
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[2]], seed=42)
    a = oracle_integrate(integrands.sphere1, solver="vegas", neval=2000, niter=2, adapt=False, config=cfg)
    b = oracle_integrate(integrands.sphere1, solver="vegas", neval=2000, niter=2, adapt=False, config=cfg)
    assert a.mean[0] != b.mean[0]
    # the sequence of seeds is itself reproducible from Configuration(seed=...)
    cfg2 = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[2]], seed=42)
    a2 = oracle_integrate(integrands.sphere1, solver="vegas", neval=2000, niter=2, adapt=False, config=cfg2)
    assert a2.mean[0] == a.mean[0]


def test_complex_block_statistics_and_average():
    """main.jl:281,285,305-308 and statistics.jl:208-215: real and imaginary parts are squared, square-rooted and
    averaged separately; std = r_std + i_std im"""
    from mcintegration_jl_b200.engine import IterationStats
    from mcintegration_jl_b200.statistics import average, mean_std
    rng = np.random.default_rng(3)
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0), dof=[[1], [1]], type=complex)
    B = 8
    z = rng.normal(size=(B, 2)) + 1j * rng.normal(size=(B, 2))
    norm = np.full(B, 100.0)
    flat = np.empty((B, 4))
    flat[:, 0::2], flat[:, 1::2] = (z * 100.0).real, (z * 100.0).imag
    st = IterationStats(flat, norm, np.zeros(0), np.zeros(3), np.zeros((3, 3, 3)), np.zeros((3, 3, 3)), 0)
    mean, std = block_statistics(st, cfg, B)
    for o in range(2):
        assert np.isclose(mean[o], z[:, o].mean())
        mr, sr = mean_std(z[:, o].real.sum(), (z[:, o].real ** 2).sum(), B)
        mi, si = mean_std(z[:, o].imag.sum(), (z[:, o].imag ** 2).sum(), B)
        assert np.isclose(std[o].real, sr) and np.isclose(std[o].imag, si)
    means = [complex(1 + 0.01 * k, 2 - 0.01 * k) for k in range(5)]
    stds = [complex(0.1, 0.2)] * 5
    m, e, c2 = average(means, stds, init=2)
    mr, er, cr = average([x.real for x in means], [x.real for x in stds], init=2)
    mi, ei, ci = average([x.imag for x in means], [x.imag for x in stds], init=2)
    assert np.isclose(m, mr + 1j * mi) and np.isclose(e, er + 1j * ei) and np.isclose(c2, cr + 1j * ci)

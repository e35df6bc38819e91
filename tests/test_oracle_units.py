"""CPU tests: the oracle (oracle/mci_oracle.c) against
  - every deterministic fixture the reference's own tests hold for this path
    (test/utility.jl, test/statistics.jl, test/mpi_test.jl), and
  - the golden vectors of tests/golden/golden.json (independent pure-Python restatement,
    tests/golden/make_golden.py; Philox = published Random123 known-answer vectors).
Bit-exact for integers / indices; doubles are compared exactly where both sides evaluate the same
expression with one rounding per operation, and to 4 ulp where libm (log, pow) is involved."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from oracle_lib import dptr, iptr, oracle

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))
ULP4 = 4 * 2.0 ** -52


def arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def test_philox_known_answers():
    lib = oracle()
    for k in GOLD["philox_kat"]:
        ctr = (C.c_uint32 * 4)(*k["ctr"])
        key = (C.c_uint32 * 2)(*k["key"])
        out = (C.c_uint32 * 4)()
        lib.orc_philox4x32_10(ctr, key, out)
        assert list(out) == k["out"]


def test_uniform_stream_layout():
    lib = oracle()
    for k in GOLD["uniforms"]:
        out = np.zeros(len(k["u"]))
        lib.orc_uniforms(k["seed"], k["iter"], k["block"], k["idx"], len(out), dptr(out))
        assert out.tolist() == k["u"]
        assert np.all((out >= 0) & (out < 1))


def test_locate_reference_fixture():
    """test/utility.jl:1-10"""
    lib = oracle()
    g = GOLD["locate"]
    acc = arr(g["acc"])
    for p, want in zip(g["p"], g["idx"]):
        got = lib.orc_locate(dptr(acc), len(acc), p)
        assert got == (want if want is not None else -1), (p, got, want)
    eps = 2.0 ** -52
    grid = arr([0.0, 0.1, 0.3, 0.5])
    assert lib.orc_locate(dptr(grid), 4, eps) == 1
    assert lib.orc_locate(dptr(grid), 4, 0.5 - eps) == 3
    assert lib.orc_locate(dptr(grid), 4, grid[0]) == 1
    assert [lib.orc_locate(dptr(grid), 4, p) for p in (0.05, 0.2, 0.31)] == [1, 2, 3]


def test_maxdof_reference_fixture():
    """test/utility.jl:12-15"""
    lib = oracle()
    dof = np.array([[1, 2, 3, 5], [3, 1, 2, 7], [2, 4, 1, 2]], dtype=np.int32)
    out = np.zeros(4, dtype=np.int32)
    lib.orc_maxdof(iptr(dof), 3, 4, iptr(out))
    assert out.tolist() == [3, 4, 3, 7]


@pytest.mark.parametrize("do_shift", [0, 1])
def test_probability_identity(do_shift):
    """test/utility.jl:30-56: total_probability == probability(i) * padding_probability(i)"""
    lib = oracle()
    rng = np.random.default_rng(7)
    d6 = rng.random(6)
    d6 /= d6.sum()
    total = np.zeros(1)
    prob = np.zeros(2)
    pad = np.zeros(2)
    for seed in range(5):
        lib.orc_probability_identity(dptr(arr([0.0, 0.1, 0.4, 1.0])), dptr(arr([0.0, 0.2, 0.6, 1.0])), dptr(arr(d6)), seed,
                                     do_shift, dptr(total), dptr(prob), dptr(pad))
        assert total[0] > 0
        for i in range(2):
            assert math.isclose(total[0], prob[i] * pad[i], rel_tol=1e-13)


def test_map_continuous_golden():
    lib = oracle()
    for k in GOLD["map_continuous"]:
        g = arr(k["grid"])
        iy = C.c_int32()
        x, p = C.c_double(), C.c_double()
        lib.orc_map_continuous(dptr(g), len(g), k["y"], C.byref(iy), C.byref(x), C.byref(p))
        assert iy.value == k["iy"]
        assert x.value == k["x"]
        assert p.value == k["prob"]


def test_smooth_rescale_train_continuous_golden():
    lib = oracle()
    for k in GOLD["train_continuous"]:
        hist = arr(k["hist"])
        sm = np.zeros_like(hist)
        lib.orc_smooth(dptr(hist), len(hist), 6.0, dptr(sm))
        assert sm.tolist() == k["smooth"]
        rs = sm.copy()
        assert lib.orc_rescale(dptr(rs), len(rs), k["alpha"]) == 0
        np.testing.assert_allclose(rs, k["rescaled"], rtol=ULP4, atol=0)
        g = arr(k["grid"])
        new = np.zeros_like(g)
        assert lib.orc_train_continuous(dptr(g), len(g), dptr(hist), k["alpha"], dptr(new)) == 0
        np.testing.assert_allclose(new, k["newgrid"], rtol=1e-13, atol=1e-15)
        assert new[0] == g[0] and new[-1] == g[-1]
        assert np.all(np.diff(new) > 0)


def test_train_discrete_golden():
    lib = oracle()
    for k in GOLD["train_discrete"]:
        hist = arr(k["hist"])
        n = len(hist)
        d, a = np.zeros(n), np.zeros(n + 1)
        assert lib.orc_train_discrete(dptr(hist), n, k["alpha"], dptr(d), dptr(a)) == 0
        np.testing.assert_allclose(d, k["dist"], rtol=ULP4)
        np.testing.assert_allclose(a, k["accum"], rtol=ULP4, atol=0)
        assert a[0] == 0.0 and math.isclose(a[-1], 1.0, rel_tol=1e-15)


def test_train_asserts_like_the_reference():
    """variable.jl:212-213, common.jl:71: non-positive / non-finite histograms are errors"""
    lib = oracle()
    g = arr([0.0, 0.25, 0.5, 1.0])
    new = np.zeros(4)
    assert lib.orc_train_continuous(dptr(g), 4, dptr(arr([1.0, 0.0, 2.0])), 2.0, dptr(new)) != 0
    assert lib.orc_train_continuous(dptr(g), 4, dptr(arr([1.0, float("nan"), 2.0])), 2.0, dptr(new)) != 0
    assert lib.orc_train_continuous(dptr(g), 4, dptr(arr([1.0, float("inf"), 2.0])), 2.0, dptr(new)) != 0


def test_mean_std_reference_fixture():
    """test/statistics.jl:22-46: _mean_std equals mean and std/sqrt(block) of the block means"""
    lib = oracle()
    for k in GOLD["mean_std"]:
        m, s = np.zeros(1), np.zeros(1)
        lib.orc_mean_std(dptr(arr([k["sum"]])), dptr(arr([k["sqsum"]])), 1, k["block"], dptr(m), dptr(s))
        assert m[0] == k["mean"] and math.isclose(s[0], k["std"], rel_tol=1e-15)
        assert math.isclose(m[0], k["sample_mean"], rel_tol=1e-14)
        assert math.isclose(s[0], k["sample_std_over_sqrtB"], rel_tol=1e-9)
    rng = np.random.default_rng(3)
    for block in (2, 7, 64):  # vector observable
        x = rng.normal(size=(block, 5))
        m, s = np.zeros(5), np.zeros(5)
        lib.orc_mean_std(dptr(arr(x.sum(0))), dptr(arr((x * x).sum(0))), 5, block, dptr(m), dptr(s))
        np.testing.assert_allclose(m, x.mean(0), rtol=1e-13)
        np.testing.assert_allclose(s, x.std(0, ddof=1) / math.sqrt(block), rtol=1e-9)
    m, s = np.zeros(1), np.ones(1)  # block == 1 -> std 0 (main.jl:311-315)
    lib.orc_mean_std(dptr(arr([2.0])), dptr(arr([4.0])), 1, 1, dptr(m), dptr(s))
    assert m[0] == 2.0 and s[0] == 0.0


def test_reweight_fixed_point_reference_fixture():
    """test/mpi_test.jl:148-168: visited=[1,2,3,4], goal=[1,2,3,4] -> reweight converges to 1/4 each"""
    lib = oracle()
    k = GOLD["reweight"]["mpi_test"]
    rw = arr(k["start"])
    for want in k["rounds"]:
        lib.orc_do_reweight(dptr(rw), dptr(arr(k["visited"])), 4, k["gamma"], dptr(arr(k["goal"])))
        np.testing.assert_allclose(rw, want, rtol=ULP4)
    np.testing.assert_allclose(rw, [0.25] * 4, rtol=1e-3)
    for g in GOLD["reweight"]["generic"]:
        rw = arr(g["rw"])
        goal = None if g["goal"] is None else arr(g["goal"])
        lib.orc_do_reweight(dptr(rw), dptr(arr(g["visited"])), len(rw), g["gamma"], dptr(goal) if goal is not None else None)
        np.testing.assert_allclose(rw, g["out"], rtol=ULP4)
        assert math.isclose(rw.sum(), 1.0, rel_tol=1e-14)


def test_average_golden():
    lib = oracle()
    for k in GOLD["average"]:
        a, e, c = C.c_double(), C.c_double(), C.c_double()
        lib.orc_average(dptr(arr(k["means"])), dptr(arr(k["stds"])), len(k["means"]), k["init"], C.byref(a), C.byref(e), C.byref(c))
        np.testing.assert_allclose([a.value, e.value, c.value], k["out"], rtol=1e-12, atol=1e-300)


def test_host_statistics_match_oracle():
    """the product's host mirror (mcintegration_jl_b200.statistics) against the oracle restatement"""
    from mcintegration_jl_b200.statistics import average, mean_std
    lib = oracle()
    rng = np.random.default_rng(11)
    x = rng.normal(2.0, 0.1, size=(16, 3))
    m, s = mean_std(x.sum(0), (x * x).sum(0), 16)
    mo, so = np.zeros(3), np.zeros(3)
    lib.orc_mean_std(dptr(arr(x.sum(0))), dptr(arr((x * x).sum(0))), 3, 16, dptr(mo), dptr(so))
    np.testing.assert_allclose(m, mo, rtol=1e-15)
    np.testing.assert_allclose(s, so, rtol=1e-12)
    means = list(rng.normal(1.0, 0.01, 10))
    stds = list(np.abs(rng.normal(0.01, 0.001, 10)))
    for init in (1, 2):
        a, e, c = C.c_double(), C.c_double(), C.c_double()
        lib.orc_average(dptr(arr(means)), dptr(arr(stds)), 10, init, C.byref(a), C.byref(e), C.byref(c))
        pa, pe, pc = average(means, stds, init=init)
        np.testing.assert_allclose([pa, pe, pc], [a.value, e.value, c.value], rtol=1e-12)


# ---- whole solver loops against the independent pure-Python restatement (tests/golden/make_golden.py) -------------------
def _loop_engine(solver, npb, measurefreq, thermal=0.1):
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import _abi, integrands
    from oracle_lib import OracleEngine
    L = GOLD["loops"]
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0, grid=L["grid"]), dof=[[2], [3]], seed=L["seed"])
    spec = mci.problem_from_config(cfg, integrands.sphere2, solver=solver, measure_kind=0, n_block=L["n_block"],
                                   neval_per_block=npb, measurefreq=measurefreq, walkers_per_block=1, thermal_ratio=thermal)
    eng = OracleEngine(spec, nthreads=1)
    eng.set_maps(grids=arr(L["grid"]), reweight=arr(L["reweight"]))
    return eng, L


def _unpack(eng, raw):
    """[hist | block_obs | block_cnt | visited | propose | accept | neval] (csrc/mci_aux.cuh PackedLayout)"""
    nb, ot, hl, nd, pa = eng.n_block, eng.obs_total, eng.hist_len, eng.nd, eng.pa_len
    o = 0
    out = {}
    for name, n in (("hist", hl), ("obs", nb * ot), ("cnt", nb), ("visited", nd), ("propose", pa), ("accept", pa), ("neval", 1)):
        out[name] = np.array(raw[o:o + n]); o += n
    out["obs"] = out["obs"].reshape(nb, ot)
    out["propose"] = out["propose"].reshape(3, nd, pa // (3 * nd)); out["accept"] = out["accept"].reshape(3, nd, pa // (3 * nd))
    return out


def test_vegas_loop_against_independent_restatement():
    from mcintegration_jl_b200 import _abi
    L0 = GOLD["loops"]["vegas"]
    eng, L = _loop_engine(_abi.SOLVER_VEGAS, L0["npb"], L0["measurefreq"])
    try:
        eng.run_local(L["iter"])
        got, want = _unpack(eng, eng.packed()), L0["out"]
        np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-14)
        assert got["cnt"].tolist() == want["norm"]
        np.testing.assert_allclose(got["hist"], want["hist"], rtol=1e-14)
    finally:
        eng.close()


def test_vegasmc_loop_against_independent_restatement():
    from mcintegration_jl_b200 import _abi
    L0 = GOLD["loops"]["vegasmc"]
    eng, L = _loop_engine(_abi.SOLVER_VEGASMC, L0["steps"], L0["measurefreq"])
    try:
        eng.run_local(L["iter"])
        got, want = _unpack(eng, eng.packed()), L0["out"]
        np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-13)
        np.testing.assert_allclose(got["cnt"], want["norm"], rtol=1e-13)
        np.testing.assert_allclose(got["hist"], want["hist"], rtol=1e-13)
        np.testing.assert_allclose(got["visited"], want["visited"], rtol=1e-13)
        assert got["propose"][1, 0, 0] == want["propose_cv"] and got["accept"][1, 0, 0] == want["accept_cv"]
        assert int(got["neval"][0]) == want["neval"]
    finally:
        eng.close()


def test_mcmc_loop_against_independent_restatement():
    from mcintegration_jl_b200 import _abi
    L0 = GOLD["loops"]["mcmc"]
    eng, L = _loop_engine(_abi.SOLVER_MCMC, L0["steps"], L0["measurefreq"], L0["thermal_ratio"])
    try:
        eng.run_local(L["iter"])
        got, want = _unpack(eng, eng.packed()), L0["out"]
        np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-13)
        np.testing.assert_allclose(got["cnt"], want["norm"], rtol=1e-13)
        assert got["hist"].tolist() == want["hist"]          # integer visit counts: exact
        assert got["visited"].tolist() == want["visited"]
        assert got["propose"].tolist() == want["propose"] and got["accept"].tolist() == want["accept"]
        assert int(got["neval"][0]) == want["neval"]
    finally:
        eng.close()


def _discrete_loop_engine(engine_cls, solver, L0, **kw):
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import integrands
    D = GOLD["loops_discrete"]
    cfg = mci.Configuration(var=mci.Discrete([(1, 3), (1, 4)]), dof=[[1]], seed=D["seed"])
    spec = mci.problem_from_config(cfg, integrands.discrete2_const, solver=solver, measure_kind=0, n_block=D["n_block"],
                                   neval_per_block=L0.get("npb", L0.get("steps")), measurefreq=L0["measurefreq"],
                                   walkers_per_block=1, thermal_ratio=L0.get("thermal_ratio", 0.1))
    eng = engine_cls(spec, **kw)
    dist = np.concatenate([arr(d) for d in D["dist"]])
    # accumulation = [0, sum(d[1:1]), sum(d[1:2]), ...], every prefix summed left to right (variable.jl:314-315)
    accum = np.concatenate([np.concatenate([[0.0], [sum(d[: i + 1], 0.0) for i in range(len(d))]]) for d in D["dist"]])
    eng.set_maps(dist=dist, accum=accum, reweight=arr(D["reweight"]))
    return eng, D


@pytest.mark.parametrize("solver_name", ["vegas", "vegasmc", "mcmc"])
def test_discrete_composite_loops_against_independent_restatement(solver_name):
    """CompositeVar of two Discrete leaves with non-uniform distributions: locate, the index draws and the composite
    probability product, whole loops, against tests/golden/make_golden.py"""
    from mcintegration_jl_b200 import _abi
    from oracle_lib import OracleEngine
    solver = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC, "mcmc": _abi.SOLVER_MCMC}[solver_name]
    L0 = GOLD["loops_discrete"][solver_name]
    eng, D = _discrete_loop_engine(OracleEngine, solver, L0, nthreads=1)
    try:
        eng.run_local(D["iter"])
        got, want = _unpack(eng, eng.packed()), L0["out"]
        np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-13)
        np.testing.assert_allclose(got["cnt"], want["norm"], rtol=1e-13)
        np.testing.assert_allclose(got["hist"], want["hist"], rtol=1e-13)
        if solver_name != "vegas":
            np.testing.assert_allclose(got["visited"], want["visited"], rtol=1e-13)
            assert int(got["neval"][0]) == want["neval"]
        if solver_name == "mcmc":
            assert got["propose"].tolist() == want["propose"] and got["accept"].tolist() == want["accept"]
    finally:
        eng.close()


def fermik_loop_spec(F, solver=None):
    """The problem behind golden.json["loops_fermik"][i]: vars (T::Continuous, K::FermiK{dim}, Ext::Discrete), mcmc"""
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import _abi, integrands
    ud = F["userdata"]
    T = mci.Continuous(0.0, ud[1], grid=F["tgrid"], alpha=3.0)
    K = mci.FermiK(F["dim"], F["kF"], F["dk"], 10.0 * F["kF"])
    Ext = mci.Discrete(1, 4, adapt=False)
    cfg = mci.Configuration(var=(T, K, Ext), dof=[[1, 1, 1]], obs=[np.zeros(4)], seed=F["seed"])
    f = integrands.Integrand("bubble_fermik", _abi.F_BUBBLE_FERMIK, 1, tuple(ud))
    spec = mci.problem_from_config(cfg, f, solver=_abi.SOLVER_MCMC if solver is None else solver,
                                   measure_kind=_abi.MEASURE_BIN_BY_DISCRETE, n_block=F["n_block"], neval_per_block=F["steps"],
                                   measurefreq=F["measurefreq"], walkers_per_block=1, thermal_ratio=F["thermal_ratio"])
    return cfg, spec


@pytest.mark.parametrize("case", [0, 1])
def test_fermik_mcmc_loop_against_independent_restatement(case):
    """FermiK{3} and FermiK{2} proposals (sampler.jl:109-281) inside the whole :mcmc loop of test/bubble_FermiK.jl,
    against the pure-Python restatement in tests/golden/make_golden.py"""
    import mcintegration_jl_b200 as mci
    from oracle_lib import OracleEngine
    F = GOLD["loops_fermik"][case]
    cfg, spec = fermik_loop_spec(F)
    eng = OracleEngine(spec, nthreads=1)
    try:
        eng.set_maps(grids=arr(F["tgrid"]), reweight=arr(F["reweight"]))
        st = eng.run_iteration(F["iter"])
        want, B1 = F["out"], F["n_block"] + 1
        np.testing.assert_allclose(st.block_obs, want["obs"], rtol=1e-12)
        np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=1e-12)
        # histogram layout: T's 5 increments, nothing for K, then Ext's 4 values (adapt = false: never filled)
        np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"] + [0.0] * 4, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st.visited - B1 * 1e-8, want["visited"], rtol=1e-9, atol=1e-6)
        assert st.neval == want["neval"]
        assert np.array_equal(np.round(st.propose - B1 * 1e-8, 3), np.array(want["propose"]))
        assert np.array_equal(np.round(st.accept - B1 * 1e-10, 3), np.array(want["accept"]))
    finally:
        eng.close()


def test_fermik_is_mcmc_only():
    """test/bubble_FermiK.jl:2,139-140: the reference runs FermiK under :mcmc alone; vegas / vegasmc are refused"""
    from mcintegration_jl_b200 import _abi
    from mcintegration_jl_b200.engine import MCIError
    from oracle_lib import OracleEngine
    for solver in (_abi.SOLVER_VEGAS, _abi.SOLVER_VEGASMC):
        _, spec = fermik_loop_spec(GOLD["loops_fermik"][0], solver)
        with pytest.raises(MCIError, match="mcmc only"):
            OracleEngine(spec, nthreads=1)


def test_fermik_bubble_oracle_vs_lindhard():
    """test/bubble_FermiK.jl:91-136: :mcmc on (T, FermiK, Ext) against the polarisation, 5 sigma (the reference compares
    with the T = 0 Lindhard value; the exact finite-temperature value is used here)"""
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import integrands as I
    from oracle_lib import oracle_integrate
    para = I.BubblePara()
    res = oracle_integrate(I.bubble_fermik(para), solver="mcmc", measure="bin_by_discrete", neval=2e5, niter=10, block=16, seed=11,
                           var=(mci.Continuous(0.0, para.β, alpha=3.0), mci.FermiK(3, para.kF, 0.2 * para.kF, 10.0 * para.kF),
                                mci.Discrete(1, para.Qsize, adapt=False)), dof=[[1, 1, 1]], obs=[np.zeros(4)])
    for i, q in enumerate(para.extQ):
        exact = I.lindhard_finite_temperature(para, q)
        assert abs(res.mean[0][i] - exact) < 5.0 * res.stdev[0][i], (q, res.mean[0][i], res.stdev[0][i], exact)
        assert res.stdev[0][i] < 0.03 * abs(exact)


def vars_loop_spec(solver_name):
    """golden.json["loops_vars"]: var = (Continuous pool, Discrete(1,4)), sphere2 on the pool, dof = [[2, 1], [3, 0]]"""
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import _abi, integrands
    G = GOLD["loops_vars"]
    L0 = G[solver_name]
    cfg = mci.Configuration(var=(mci.Continuous(0.0, 1.0, grid=G["grid"]), mci.Discrete(1, 4)), dof=[[2, 1], [3, 0]], seed=G["seed"])
    solver = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC}[solver_name]
    spec = mci.problem_from_config(cfg, integrands.sphere2, solver=solver, measure_kind=0, n_block=G["n_block"],
                                   neval_per_block=L0.get("npb", L0.get("steps")), measurefreq=L0["measurefreq"], walkers_per_block=1)
    d = G["dist"]
    maps = dict(grids=arr(G["grid"]), dist=arr(d), accum=arr([0.0] + [sum(d[: i + 1], 0.0) for i in range(len(d))]), reweight=arr(G["reweight"]))
    return spec, maps, G, L0


def check_vars_loop(st, G, L0, solver_name, rtol=1e-12):
    want, B1 = L0["out"], G["n_block"] + 1
    np.testing.assert_allclose(st.block_obs, want["obs"], rtol=rtol)
    np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=rtol)
    np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"], rtol=1e-9, atol=1e-9)
    if solver_name == "vegasmc":
        np.testing.assert_allclose(st.visited - B1 * 1e-8, want["visited"], rtol=1e-9, atol=1e-6)
        assert st.neval == want["neval"]
        # propose/accept[2, 1, vi] (vegas_mc/updates.jl:90,92): row-major [3][Nd][max(Nd, Nv)] -> [1][0][vi]
        assert np.array_equal(np.round(st.propose[1, 0, :2] - B1 * 1e-8, 3), np.array(want["propose_cv"]))
        assert np.array_equal(np.round(st.accept[1, 0, :2] - B1 * 1e-10, 3), np.array(want["accept_cv"]))


@pytest.mark.parametrize("solver_name", ["vegas", "vegasmc"])
def test_two_variable_loops_against_independent_restatement(solver_name):
    """Several top-level variables with padding across them (the layout class of example/bubble.jl), whole loops"""
    from oracle_lib import OracleEngine
    spec, maps, G, L0 = vars_loop_spec(solver_name)
    eng = OracleEngine(spec, nthreads=1)
    try:
        eng.set_maps(**maps)
        check_vars_loop(eng.run_iteration(G["iter"]), G, L0, solver_name)
    finally:
        eng.close()


@pytest.mark.parametrize("solver_name", ["vegas", "vegasmc"])
def test_binned_measure_loops_against_independent_restatement(solver_name):
    """The bubble's measure (example/bubble.jl:81-84: bin the integrand by the last, non-adaptive Discrete variable) inside
    whole :vegas / :vegasmc loops.  The GPU runs this measure against the oracle in the L1 bubble tests."""
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import _abi, integrands
    from oracle_lib import OracleEngine
    G = GOLD["loops_binned"]
    L0 = G[solver_name]
    cfg = mci.Configuration(var=(mci.Continuous(0.0, 1.0, grid=G["grid"]), mci.Discrete(1, 4, adapt=False)), dof=[[1, 1]],
                            obs=[np.zeros(4)], seed=G["seed"])
    solver = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC}[solver_name]
    spec = mci.problem_from_config(cfg, integrands.discrete_x, solver=solver, measure_kind=_abi.MEASURE_BIN_BY_DISCRETE,
                                   n_block=G["n_block"], neval_per_block=L0.get("npb", L0.get("steps")),
                                   measurefreq=L0["measurefreq"], walkers_per_block=1)
    eng = OracleEngine(spec, nthreads=1)
    try:
        eng.set_maps(grids=arr(G["grid"]), reweight=arr(G["reweight"]))
        st = eng.run_iteration(G["iter"])
        want, B1 = L0["out"], G["n_block"] + 1
        np.testing.assert_allclose(st.block_obs, want["obs"], rtol=1e-12)
        np.testing.assert_allclose(st.block_norm - 1e-10, want["norm"], rtol=1e-12)
        # histogram: the pool's 8 increments, then the 4 values of the non-adaptive Discrete (never filled)
        np.testing.assert_allclose(st.hist - B1 * 1e-10, want["hist"] + [0.0] * 4, rtol=1e-9, atol=1e-9)
        if solver_name == "vegasmc":
            np.testing.assert_allclose(st.visited - B1 * 1e-8, want["visited"], rtol=1e-9, atol=1e-6)
            assert st.neval == want["neval"]
    finally:
        eng.close()


# ---- complex integrands (Configuration(type=ComplexF64), main.jl:276-286): the three loops on TestComplex2 -------------
def _complex_loop_engine(solver, npb, measurefreq, thermal=0.1):
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import integrands
    from oracle_lib import OracleEngine
    L = GOLD["loops_complex"]
    cfg = mci.Configuration(var=mci.Continuous(0.0, 1.0, grid=L["grid"]), dof=[[1], [1]], seed=L["seed"], type=complex)
    spec = mci.problem_from_config(cfg, integrands.complex2, solver=solver, measure_kind=0, n_block=L["n_block"],
                                   neval_per_block=npb, measurefreq=measurefreq, walkers_per_block=1, thermal_ratio=thermal)
    eng = OracleEngine(spec, nthreads=1)
    eng.set_maps(grids=arr(L["grid"]), reweight=arr(L["reweight"]))
    return eng, L


@pytest.mark.parametrize("solver_name", ["vegas", "vegasmc", "mcmc"])
def test_complex_loops_against_independent_restatement(solver_name):
    """golden.json["loops_complex"]: f = (x, x^2 im) (test/montecarlo.jl:172-198) through the three whole loops, restated
    in pure Python with Python's complex arithmetic (abs = hypot; complex x real component-wise, like Julia).  The
    observables cross the boundary as (re, im) pairs: [block][integrand][2]."""
    from mcintegration_jl_b200 import _abi
    sid = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC, "mcmc": _abi.SOLVER_MCMC}[solver_name]
    L0 = GOLD["loops_complex"][solver_name]
    eng, L = _complex_loop_engine(sid, L0.get("npb", L0.get("steps")), L0["measurefreq"], L0.get("thermal_ratio", 0.1))
    try:
        assert eng.obs_total == 4  # two complex observables = four doubles
        eng.run_local(L["iter"])
        got, want = _unpack(eng, eng.packed()), L0["out"]
        np.testing.assert_allclose(got["obs"], want["obs"], rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(got["cnt"], want["norm"], rtol=1e-13)
        np.testing.assert_allclose(got["hist"], want["hist"], rtol=1e-13)
        if solver_name != "vegas":
            np.testing.assert_allclose(got["visited"], want["visited"], rtol=1e-13)
            assert int(got["neval"][0]) == want["neval"]
        if solver_name == "mcmc":
            np.testing.assert_array_equal(got["propose"], np.array(want["propose"]))
            np.testing.assert_array_equal(got["accept"], np.array(want["accept"]))
    finally:
        eng.close()


def test_mcmc_start_integrand_and_burnin_counters():
    """`idx`, the integrand an :mcmc chain starts in (mcmc/montecarlo.jl:76): starting in the normalisation integrand needs
    no evaluation and is always a valid start.  Extra burn-in steps (mci_problem_t.chain_burnin) touch no statistic:
    propose / accept / neval count the reference's loop only."""
    from cases import make_spec
    from oracle_lib import OracleEngine
    stats = {}
    for start, burn in ((0, 0), (3, 0), (0, 500)):
        spec, _ = make_spec("sphere2", "mcmc", n_block=2, neval_per_block=400, walkers=1, chain_burnin=burn, seed=9)
        spec.mcmc_start_idx = start
        eng = OracleEngine(spec, nthreads=1)
        try:
            stats[(start, burn)] = eng.run_iteration(0)
        finally:
            eng.close()
    a, b, c = stats[(0, 0)], stats[(3, 0)], stats[(0, 500)]
    assert not np.array_equal(a.visited, b.visited)  # a different start, a different chain
    total = 2 * (400 + 40)                            # neval + nburnin steps per block, every one counted in visited
    for st in (a, b, c):
        assert abs(st.visited.sum() - 3 * 3e-8 - total) < 1e-6
        assert (st.propose - 3e-8).sum() <= total + 1e-6 and st.neval <= total  # burn-in steps add nothing

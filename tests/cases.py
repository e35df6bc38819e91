"""Problem cases shared by the CPU (oracle) tests and the GPU parity tests.  Each case mirrors a
testset of the reference (test/montecarlo.jl, test/bubble.jl) or a BASELINE.json config."""
from __future__ import annotations

import math

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import _abi, integrands as I

PI = math.pi


def case(name):
    """-> dict(integrand, config kwargs factory, measure, expect (list or None))"""
    para = I.BubblePara()
    C = {
        # BASELINE C1 / test/montecarlo.jl:129-134 (TestSingular1)
        "log_over_sqrt": dict(f=I.log_over_sqrt, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[1]]), expect=[-4.0]),
        # BASELINE C2
        "gauss10": dict(f=I.gauss_nd, cfg=lambda: dict(var=mci.Continuous(-5.0, 5.0), dof=[[10]]),
                        expect=[math.erf(5 / math.sqrt(2)) ** 10]),
        "gauss4": dict(f=I.gauss_nd, cfg=lambda: dict(var=mci.Continuous(-5.0, 5.0), dof=[[4]]),
                       expect=[math.erf(5 / math.sqrt(2)) ** 4]),
        # BASELINE C3, layouts A (pool) and B (composite, reduced to 6-D for test speed)
        "genz16_pool": dict(f=I.genz_corner_peak(16, "pool"), cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[16]]),
                            expect=[4.17984992413838e-05]),
        "genz6_comp": dict(f=I.genz_corner_peak(6, "comp"),
                           cfg=lambda: dict(var=mci.Continuous([(0.0, 1.0)] * 6), dof=[[1]]),
                           expect=[I.genz_corner_peak_exact(I.genz_corner_peak(6, "comp").userdata)]),
        "genz16_comp": dict(f=I.genz_corner_peak(16, "comp"),
                            cfg=lambda: dict(var=mci.Continuous([(0.0, 1.0)] * 16), dof=[[1]]),
                            expect=[4.17984992413838e-05]),
        # test/montecarlo.jl:4-9 (Sphere1), :19-49 (Sphere2/3)
        "sphere1": dict(f=I.sphere1, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[2]]), expect=[PI / 4]),
        "sphere2": dict(f=I.sphere2, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[2], [3]]),
                        expect=[PI / 4, 4 * PI / 3 / 8]),
        "sphere2_offset": dict(f=I.sphere2, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0, offset=2), dof=[[2], [3]],
                                                              neighbor=[(1, 3), (2, 3)]),
                               expect=[PI / 4, 4 * PI / 3 / 8]),
        # test/montecarlo.jl:104-120 (Discrete)
        "discrete_x": dict(f=I.discrete_x, cfg=lambda: dict(var=mci.Discrete(1, 3), dof=[[1]]), expect=[6.0]),
        "discrete2": dict(f=I.discrete2_const, cfg=lambda: dict(var=mci.Discrete([(1, 3), (1, 4)]), dof=[[1]]),
                          expect=[12.0]),
        # test/montecarlo.jl:136-180 (TestSingular2 in three layouts)
        "lattice_pool": dict(f=I.lattice_pool, cfg=lambda: dict(var=mci.Continuous(0.0, PI), dof=[[3]]), expect=[1.3932]),
        "lattice_comp": dict(f=I.lattice_comp,
                             cfg=lambda: dict(var=mci.CompositeVar(mci.Continuous(0.0, PI), mci.Continuous(0.0, PI),
                                                                   mci.Continuous(0.0, PI)), dof=[[1]]),
                             expect=[1.3932]),
        "lattice_vec": dict(f=I.lattice_comp, cfg=lambda: dict(var=mci.Continuous([(0.0, PI)] * 3), dof=[[1]]),
                            expect=[1.3932]),
        # test/montecarlo.jl:200-216 (hypersphere volumes, N=3)
        "hypersphere3": dict(f=I.hypersphere3, cfg=lambda: dict(var=mci.Continuous(-1.0, 1.0), dof=[[2], [3], [4]]),
                             expect=[0.9230, 0.94724, 0.96118]),
        # test/thread.jl
        "poly3": dict(f=I.poly3, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[1], [1], [1]]),
                      expect=[1 / 2, 1 / 3, 1 / 4]),
        "x2y2": dict(f=I.x2_plus_y2, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[2]]), expect=[2 / 3]),
        # test/montecarlo.jl:166-198 (TestComplex1 / TestComplex2): Configuration(type=ComplexF64)
        "complex1": dict(f=I.complex1, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[1]], type=complex),
                         expect=[0.5 + 1j / 3]),
        "complex2": dict(f=I.complex2, cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[1], [1]], type=complex),
                         expect=[0.5 + 0j, 1j / 3]),
        # BASELINE C5 integrand
        "product4": dict(f=I.product4(), cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[20]] * 4),
                         expect=[1.0, 1.0, 1.0, 1.0]),
        "product4_d6": dict(f=I.product4(), cfg=lambda: dict(var=mci.Continuous(0.0, 1.0), dof=[[6]] * 4),
                            expect=[1.0, 1.0, 1.0, 1.0]),
        # BASELINE C4: example/bubble.jl:89-102
        "bubble": dict(f=I.bubble(para), measure="bin_by_discrete",
                       cfg=lambda: dict(var=(mci.Continuous(0.0, 1.0, alpha=3.0), mci.Continuous(0.0, PI, alpha=3.0),
                                             mci.Continuous(0.0, 2 * PI, alpha=3.0), mci.Continuous(0.0, para.β, alpha=3.0),
                                             mci.Discrete(1, para.Qsize, adapt=False)),
                                        dof=[[1, 1, 1, 1, 1]], obs=[[0.0] * para.Qsize]),
                       expect=None, para=para),
        # test/bubble_FermiK.jl:91-105: the same diagram on (T, K::FermiK{3}, Ext), :mcmc only
        "bubble_fermik": dict(f=I.bubble_fermik(para), measure="bin_by_discrete",
                              cfg=lambda: dict(var=(mci.Continuous(0.0, para.β, alpha=3.0),
                                                    mci.FermiK(3, para.kF, 0.2 * para.kF, 10.0 * para.kF),
                                                    mci.Discrete(1, para.Qsize, adapt=False)),
                                               dof=[[1, 1, 1]], obs=[[0.0] * para.Qsize]),
                              expect=None, para=para),
    }
    return C[name]


SOLVER_ID = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC, "mcmc": _abi.SOLVER_MCMC}
MEASURE_ID = {None: _abi.MEASURE_DEFAULT, "bin_by_discrete": _abi.MEASURE_BIN_BY_DISCRETE}


def make_spec(name, solver, *, n_block=16, neval_per_block=4096, seed=1234, walkers=1, measurefreq=1,
              thermal_ratio=0.1, hist_policy=0, chain_burnin=0):
    c = case(name)
    cfg = mci.Configuration(seed=seed, **c["cfg"]())
    spec = mci.problem_from_config(cfg, c["f"], solver=SOLVER_ID[solver], measure_kind=MEASURE_ID[c.get("measure")],
                                   n_block=n_block, neval_per_block=neval_per_block, measurefreq=measurefreq,
                                   walkers_per_block=walkers, thermal_ratio=thermal_ratio, hist_policy=hist_policy,
                                   chain_burnin=chain_burnin)
    return spec, cfg

"""Host mirror of src/configuration.jl: normalises var/dof/obs/reweight/neighbor and marshals the
fields the solver loops consume (configuration.jl:37-60) into the POD `mci_problem_t`."""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass

import numpy as np

from . import _abi
from .variables import CompositeVar, Continuous, Discrete, FermiK, is_variable


def _maxdof(dof):
    """configuration.jl:229-236"""
    return [max(d[i] for d in dof) for i in range(len(dof[0]))]


def _neighbor(neighbor, Nd):
    """configuration.jl:201-227 (1-based vertex lists, like the reference)."""
    if neighbor is None:
        nb = [[d - 1, d + 1] for d in range(1, Nd + 1)]
        nb[0] = [2] if Nd == 2 else [Nd, 2]
        nb[-1] = [1]
        if Nd >= 3:
            nb[-2] = [Nd - 2]
        return nb
    if len(neighbor) > 0 and isinstance(neighbor[0], tuple):
        adj = [set() for _ in range(Nd)]
        for a, b in neighbor:
            adj[a - 1].add(b)
            adj[b - 1].add(a)
        # is_connected(g)
        seen, stack = {1}, [1]
        while stack:
            v = stack.pop()
            for w in adj[v - 1]:
                if w not in seen:
                    seen.add(w)
                    stack.append(w)
        assert len(seen) == Nd, "The neighbor graph is not connected."
        return [sorted(a) for a in adj]
    nb = [list(n) for n in neighbor]
    assert len(nb) == Nd, f"{Nd} elements are expected for neighbor={neighbor}"
    return nb


class Configuration:
    """Configuration(; var, dof, type, obs, reweight, seed, neighbor, userdata) (configuration.jl:105-194)."""

    def __init__(self, *, var=None, dof=None, type=float, obs=None, reweight=None, seed=None, neighbor=None,
                 userdata=None, **_ignored):
        if var is None:
            var = (Continuous(0.0, 1.0),)
        if is_variable(var):
            single = True
            var = (var,)
        elif isinstance(var, (list, tuple)):
            single = False
            assert all(is_variable(v) for v in var), "All elements in var should be variables"
            var = tuple(var)
        else:
            raise TypeError("Configuration.var should be a variable, a vector of variables, or a tuple of variables.")
        Nv = len(var)
        if dof is None:
            dof = 1 if single else [[1] * Nv]
        if isinstance(dof, int):  # one integral with one variable (:134-136)
            assert Nv == 1, "Only one type of variable is allowed when dof is an integer"
            dof = [[dof]]
        elif isinstance(dof, np.ndarray) and dof.ndim == 2:  # each column is one integral (:139-140)
            dof = [list(map(int, dof[:, i])) for i in range(dof.shape[1])]
        else:
            dof = list(dof)
            if all(isinstance(d, (list, tuple, np.ndarray)) for d in dof):
                dof = [list(map(int, d)) for d in dof]
            elif all(isinstance(d, (int, np.integer)) for d in dof):
                dof = [[int(d)] for d in dof]
            else:
                raise TypeError("Configuration.dof should be a Vector{Int} or Vector{Vector{Int}}")
        dof.append([0] * Nv)  # normalisation diagram (:153)
        self.var = var
        self.dof = dof
        self.maxdof = _maxdof(dof)
        for i, m in enumerate(self.maxdof):  # pool auto-resize (:156-160)
            if m + var[i].offset >= var[i].poolsize() - 2:
                var[i].resize(m + 2 + var[i].offset)
        Nd = len(dof)
        assert Nd > 1, "At least one integrand is required."
        assert all(len(d) == Nv for d in dof), "Each element of `dof` should have the same dimension as `var`"
        if obs is None:
            obs = [0.0] * (Nd - 1)
        assert len(obs) == Nd - 1, "The number of observables should be equal to the number of integrands"
        self.type = type
        dt = np.complex128 if self.is_complex() else np.float64
        self.observable = [np.zeros(len(o), dtype=dt) if isinstance(o, (list, tuple, np.ndarray)) else dt(0.0) for o in obs]
        self.N = Nd - 1
        self.norm = Nd
        self.neighbor = _neighbor(neighbor, Nd)
        if reweight is None:
            reweight = np.ones(Nd)
        reweight = np.array(reweight, dtype=np.float64)
        assert len(reweight) == Nd, "Wrong reweight vector size!"
        assert np.all(reweight > 0), "All reweight factors should be positive."
        self.reweight = reweight / reweight.sum()
        self.seed = int(seed) if seed is not None else random.SystemRandom().randint(1, 1000000)
        # config.rng = MersenneTwister(seed) (configuration.jl:177-178): every integrate() call draws the seed of its
        # sample streams from it (main.jl:136), so successive calls on one Configuration use independent samples
        self.rng = random.Random(self.seed)
        self.userdata = userdata
        self.normalization = 1.0e-10
        self.visited = np.full(Nd, 1.0e-8)
        pa = max(Nd, Nv)
        self.propose = np.full((3, Nd, pa), 1.0e-8)
        self.accept = np.zeros((3, Nd, pa))
        self.neval = 0

    # ---- marshalling -------------------------------------------------------
    def leaves(self):
        out = []
        for v in self.var:
            out.extend(v.leaves())
        return out

    def obs_len(self):
        return [len(o) if isinstance(o, np.ndarray) else 1 for o in self.observable]

    def is_complex(self):
        """type = ComplexF64 (configuration.jl:108): integrand values and observables are complex"""
        return self.type is complex or self.type in (np.complex128, "ComplexF64", "complex")

    def next_seed(self):
        """rand(config.rng, 1:1000000): the seed one integrate() call gives its solver streams (main.jl:136)"""
        return self.rng.randint(1, 1000000)

    def __repr__(self):
        return f"Configuration(N={self.N}, var={self.var}, dof={self.dof[:-1]}, seed={self.seed})"


@dataclass
class ProblemSpec:
    """Everything mci_create needs, as plain Python; to_ctypes() keeps the buffers alive."""
    solver: int
    integrand_id: int
    measure_kind: int
    n_integrand: int
    topvar_nleaf: list
    topvar_offset: list
    leaves: list  # dicts: kind,nbins,adapt,lower_i,lower,upper,alpha
    dof: list  # [Ni][Nv]
    obs_len: list
    neighbor: list | None  # 1-based adjacency lists over Nd nodes or None
    userdata: list
    n_block: int
    neval_per_block: int
    seed: int
    measurefreq: int = 1
    walkers_per_block: int = 1
    thermal_ratio: float = 0.1
    device: int = 0
    hist_policy: int = 0
    jit: int = 0
    user_source: str | None = None
    chain_burnin: int = 0
    mcmc_start_idx: int = 0  # `idx` of MCMC.montecarlo (mcmc/montecarlo.jl:76), 1-based; 0 = default 1
    weight_kind: int = 0     # _abi.WEIGHT_REAL / WEIGHT_COMPLEX

    def to_ctypes(self):
        keep = []

        def arr(ct, vals):
            a = (ct * max(1, len(vals)))(*vals)
            keep.append(a)
            return a

        p = _abi.Problem()
        p.solver, p.integrand_id, p.measure_kind = self.solver, self.integrand_id, self.measure_kind
        p.n_integrand = self.n_integrand
        p.n_topvar = len(self.topvar_nleaf)
        p.n_leaf = len(self.leaves)
        p.topvar_nleaf = arr(C.c_int32, self.topvar_nleaf)
        p.topvar_offset = arr(C.c_int32, self.topvar_offset)
        leaf_arr = (_abi.Leaf * len(self.leaves))()
        for i, L in enumerate(self.leaves):
            leaf_arr[i] = _abi.Leaf(L["kind"], L["nbins"], int(L["adapt"]), L.get("lower_i", 0), L.get("lower", 0.0),
                                    L.get("upper", 0.0), L["alpha"])
        keep.append(leaf_arr)
        p.leaf = leaf_arr
        p.dof = arr(C.c_int32, [x for row in self.dof for x in row])
        p.obs_len = arr(C.c_int32, self.obs_len)
        if self.neighbor is not None:
            ptr, idx = [0], []
            for nb in self.neighbor:
                idx.extend(int(x) - 1 for x in nb)
                ptr.append(len(idx))
            p.neighbor_ptr = arr(C.c_int32, ptr)
            p.neighbor_idx = arr(C.c_int32, idx)
        p.userdata = arr(C.c_double, list(self.userdata))
        p.n_userdata = len(self.userdata)
        p.n_block = self.n_block
        p.neval_per_block = self.neval_per_block
        p.seed = self.seed
        p.measurefreq = self.measurefreq
        p.walkers_per_block = self.walkers_per_block
        p.thermal_ratio = self.thermal_ratio
        p.device = self.device
        p.hist_policy = self.hist_policy
        p.jit = self.jit
        p.chain_burnin = int(self.chain_burnin)
        p.mcmc_start_idx = int(self.mcmc_start_idx)
        p.weight_kind = int(self.weight_kind)
        if self.user_source is not None:
            src = self.user_source.encode()
            keep.append(src)
            p.user_source = src
        return p, keep


def problem_from_config(config: Configuration, integrand, *, solver: int, measure_kind: int, n_block: int,
                        neval_per_block: int, measurefreq: int = 1, walkers_per_block: int = 1,
                        thermal_ratio: float = 0.1, device: int = 0, hist_policy: int = 0,
                        default_neighbor: bool = False, jit: int = 0, chain_burnin: int = 0,
                        measure_source: str | None = None, seed: int | None = None, mcmc_start_idx: int = 0) -> ProblemSpec:
    leaves = []
    for L in config.leaves():
        if isinstance(L, Continuous):
            leaves.append(dict(kind=_abi.LEAF_CONTINUOUS, nbins=L.nbins, adapt=L.adapt, lower=L.lower, upper=L.upper,
                               alpha=L.alpha))
        elif isinstance(L, FermiK):  # one leaf per Cartesian component: (kF, δk, maxK) ride in (lower, upper, alpha)
            leaves.append(dict(kind=_abi.LEAF_FERMIK, nbins=0, adapt=False, lower=L.kF, upper=L.δk, alpha=L.maxK))
        else:
            leaves.append(dict(kind=_abi.LEAF_DISCRETE, nbins=L.nbins, adapt=L.adapt, lower_i=L.lower, alpha=L.alpha))
    sources = [x for x in (getattr(integrand, "source", None), measure_source) if x]
    assert integrand.n_integrand == config.N, (
        f"integrand {integrand.name} returns {integrand.n_integrand} values but dof has {config.N} rows")
    assert bool(getattr(integrand, "complex", False)) == config.is_complex(), (
        f"integrand {integrand.name} is {'complex' if getattr(integrand, 'complex', False) else 'real'}-valued: "
        f"Configuration(type=...) must say the same (configuration.jl:108)")
    return ProblemSpec(
        solver=solver, integrand_id=integrand.id, measure_kind=measure_kind, n_integrand=config.N,
        topvar_nleaf=[len(v.leaves()) for v in config.var], topvar_offset=[v.offset for v in config.var],
        leaves=leaves, dof=[list(d) for d in config.dof[:-1]], obs_len=config.obs_len(),
        neighbor=None if default_neighbor else config.neighbor, userdata=list(integrand.userdata),
        n_block=n_block, neval_per_block=neval_per_block, seed=config.seed if seed is None else int(seed),
        measurefreq=measurefreq, mcmc_start_idx=mcmc_start_idx,
        weight_kind=_abi.WEIGHT_COMPLEX if config.is_complex() else _abi.WEIGHT_REAL,
        walkers_per_block=walkers_per_block, thermal_ratio=thermal_ratio, device=device, hist_policy=hist_policy, jit=jit,
        user_source="\n".join(sources) if sources else None, chain_burnin=chain_burnin)


def maps_from_config(config: Configuration):
    """Concatenated maps in ABI order (include/mcib200.h, mci_set_maps)."""
    grids, dist, accum = [], [], []
    for L in config.leaves():
        if isinstance(L, Continuous):
            grids.append(L.grid)
        elif isinstance(L, Discrete):
            dist.append(L.distribution)
            accum.append(L.accumulation)
    cat = lambda xs: np.concatenate(xs).astype(np.float64) if xs else np.zeros(0)
    return cat(grids), cat(dist), cat(accum), np.array(config.reweight, dtype=np.float64)


def maps_to_config(config: Configuration, grids, dist, accum, reweight):
    go = do = ao = 0
    for L in config.leaves():
        if isinstance(L, Continuous):
            n = L.nbins + 1
            L.grid = np.array(grids[go:go + n])
            go += n
        elif isinstance(L, Discrete):
            n = L.nbins
            L.distribution = np.array(dist[do:do + n])
            L.accumulation = np.array(accum[ao:ao + n + 1])
            do += n
            ao += n + 1
    config.reweight = np.array(reweight)

"""Host-side mirrors of the reference's variable pools (src/distribution/variable.jl).

Only what the hot path's boundary needs lives here: bounds, learning rate, adapt flag and the
adaptive map itself (grid / distribution).  Sampling state (data/gidx/prob pools) lives on the
device inside the fill kernels.
"""
from __future__ import annotations

import numpy as np

MaxOrder = 16  # distribution.jl:59


def _is_bounds_list(x):
    return isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], (list, tuple))


class Continuous:
    """Continuous(lower, upper, size=16; offset=0, alpha=2.0, adapt=true, ninc=1000, grid=...)
    (variable.jl:137-153).  `Continuous([(l1,u1), (l2,u2), ...])` returns a CompositeVar
    (variable.jl:174-188)."""

    def __new__(cls, lower, upper=None, size=MaxOrder, **kw):
        if _is_bounds_list(lower):
            bounds = lower
            if upper is not None and not isinstance(upper, (list, tuple)):
                size = upper
            n = len(bounds)
            ninc = kw.pop("ninc", [1000] * n)
            grid = kw.pop("grid", [None] * n)
            offset = kw.get("offset", 0)
            adapt = kw.get("adapt", True)
            leaves = [Continuous(b[0], b[1], size, ninc=ninc[i], grid=grid[i], **kw) for i, b in enumerate(bounds)]
            return CompositeVar(*leaves, adapt=adapt, offset=offset, size=size)
        return super().__new__(cls)

    def __init__(self, lower, upper=None, size=MaxOrder, *, offset=0, alpha=2.0, adapt=True, ninc=1000, grid=None):
        if _is_bounds_list(lower):
            return
        assert offset + 1 < size  # variable.jl:138
        assert upper > lower + 2 * np.finfo(float).eps  # :140
        self.lower = float(lower)
        self.upper = float(upper)
        self.size = size + 1  # one more element as cache for the swap operation (:139)
        self.offset = int(offset)
        self.alpha = float(alpha)
        self.adapt = bool(adapt)
        if grid is None:
            # collect(LinRange(lower, upper, ninc)); Julia's lerp: (1-t)*a + t*b
            t = np.arange(ninc, dtype=np.float64) / (ninc - 1)
            grid = (1 - t) * self.lower + t * self.upper
            grid[0], grid[-1] = self.lower, self.upper
        self.grid = np.array(grid, dtype=np.float64)
        self.histogram = np.full(len(self.grid) - 1, 4.9406564584124654e-324 * 1e50)  # ones(N)*TINY (:149)

    kind = 0

    @property
    def nbins(self):
        return len(self.grid) - 1

    def poolsize(self):
        return self.size

    def resize(self, size):  # Base.resize! (variable.jl:155-160)
        assert size > self.size
        self.size = size

    def leaves(self):
        return [self]

    def __repr__(self):
        return (("Adaptive" if self.adapt else "Nonadaptive") + f" Continuous variable ∈ [{self.lower}, {self.upper})."
                + (f" Learning rate = {self.alpha}." if self.adapt else "")
                + (f" Offset = {self.offset}." if self.offset > 0 else ""))


class Discrete:
    """Discrete(lower, upper, size=16; distribution=nothing, offset=0, alpha=2.0, adapt=true)
    (variable.jl:299-325).  `Discrete([(l1,u1), ...])` returns a CompositeVar (:343-354)."""

    def __new__(cls, lower, upper=None, size=MaxOrder, **kw):
        if _is_bounds_list(lower):
            bounds = lower
            if upper is not None and not isinstance(upper, (list, tuple)):
                size = upper
            dist = kw.pop("distribution", [None] * len(bounds))
            leaves = [Discrete(b[0], b[1], size, distribution=dist[i], **kw) for i, b in enumerate(bounds)]
            return CompositeVar(*leaves, adapt=kw.get("adapt", True), offset=kw.get("offset", 0), size=size)
        return super().__new__(cls)

    def __init__(self, lower, upper=None, size=MaxOrder, *, distribution=None, offset=0, alpha=2.0, adapt=True):
        if _is_bounds_list(lower):
            return
        if isinstance(lower, (tuple, list)):  # Discrete((lo, hi)) (:327-329)
            if upper is not None:
                size = upper
            lower, upper = lower
        assert offset + 1 < size
        assert upper >= lower
        self.lower = int(lower)
        self.upper = int(upper)
        self.size = size + 1
        self.offset = int(offset)
        self.alpha = float(alpha)
        self.adapt = bool(adapt)
        n = self.upper - self.lower + 1
        if distribution is None:
            distribution = np.ones(n)
        distribution = np.array(distribution, dtype=np.float64)
        assert np.all(distribution >= 0.0) and len(distribution) == n
        self.set_distribution(distribution / distribution.sum())
        self.histogram = np.full(n, 1.0e-10)  # clearStatistics! (:322, :565)

    kind = 1

    def set_distribution(self, dist):
        """distribution + accumulation = [0, Σ₁, Σ₂, …] with every prefix recomputed (variable.jl:314-315)."""
        self.distribution = np.array(dist, dtype=np.float64)
        acc = [0.0]
        for i in range(1, len(dist) + 1):
            s = 0.0
            for k in range(i):
                s += float(self.distribution[k])
            acc.append(s)
        self.accumulation = np.array(acc)

    @property
    def nbins(self):
        return self.upper - self.lower + 1

    def poolsize(self):
        return self.size

    def resize(self, size):
        assert size > self.size
        self.size = size

    def leaves(self):
        return [self]

    def __repr__(self):
        return (("Adaptive" if self.adapt else "Nonadaptive")
                + f" Discrete variable ∈ [{self.lower}, ..., {self.upper}]."
                + (f" Learning rate = {self.alpha}." if self.adapt else "")
                + (f" Offset = {self.offset}." if self.offset > 0 else ""))


class FermiK:
    """FermiK(dim, kF, δk, maxK, size=16; offset=0) (variable.jl:1-21): a momentum proposed near the Fermi surface
    (sampler.jl:109-281).  Has no map and no histogram; like in the reference (test/bubble_FermiK.jl:2,139-140) it
    works under solver=:mcmc only."""

    def __init__(self, dim, kF, δk, maxK, size=MaxOrder, *, offset=0):
        assert offset + 1 < size
        assert dim in (2, 3), "FermiK is implemented for dim = 2 or 3 (sampler.jl:124-148)"
        self.dim = int(dim)
        self.kF = float(kF)
        self.δk = float(δk)
        self.maxK = float(maxK)
        self.size = size  # k = zeros(dim, size) .+ kF / sqrt(dim): no extra cache element (variable.jl:13)
        self.offset = int(offset)
        self.adapt = False
        self.histogram = np.array([0.0])

    kind = 2

    def poolsize(self):
        return self.size

    def resize(self, size):  # not defined by the reference for FermiK; pools only ever grow
        assert size > self.size
        self.size = size

    def leaves(self):
        return [self] * self.dim  # one ABI leaf per component

    def __repr__(self):
        return (f"{self.dim}D FermiK variable ∈ [0, {self.maxK})." + f" Max number = {self.size - 1 - self.offset}."
                + (f" Offset = {self.offset}." if self.offset > 0 else ""))


class CompositeVar:
    """CompositeVar(vargs...; adapt=true, offset=0, size=16) (variable.jl:415-427): the bundled
    leaves are sampled together; the probability is the product of the leaves' probabilities."""

    def __init__(self, *vargs, adapt=True, offset=0, size=MaxOrder):
        assert all(isinstance(v, (Continuous, Discrete)) for v in vargs), "all arguments should variables"
        for v in vargs:
            v.adapt = bool(adapt)
            v.offset = int(offset)
        self.vars = tuple(vargs)
        self.adapt = bool(adapt)
        self.offset = int(offset)
        self.size = size

    def poolsize(self):
        return self.size

    def resize(self, size):  # variable.jl:455-472
        assert size > self.size
        for v in self.vars:
            if size > v.size:
                v.resize(size)
        self.size = size

    def leaves(self):
        return list(self.vars)

    def __len__(self):
        return len(self.vars)

    def __getitem__(self, i):
        return self.vars[i]

    def __iter__(self):
        return iter(self.vars)

    def __repr__(self):
        return (("Adaptive" if self.adapt else "Nonadaptive") + f" CompositeVar with {len(self.vars)} components."
                + (f" Offset = {self.offset}." if self.offset > 0 else ""))


def is_variable(v):
    return isinstance(v, (Continuous, Discrete, CompositeVar, FermiK))

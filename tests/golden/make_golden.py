#!/usr/bin/env python
"""Generates tests/golden/golden.json — golden vectors for the deterministic pieces of the hot path.

The reference is Julia and cannot run in this image (no julia binary), so these vectors are NOT outputs of
the reference; they come from this independent pure-Python restatement of the cited reference lines
(IEEE doubles, one rounding per operation, no FMA, strictly left-to-right sums).  They pin the C oracle
(oracle/mci_oracle.c) and, through it, the CUDA path against a second, independently written statement of
the same arithmetic.  The Philox vectors are the published Random123 known-answer vectors
(Random123 kat_vectors, philox4x32 10 rounds).  The first cases of every group are the hand cases listed in
SURVEY.md 8(c).

Run:  python tests/golden/make_golden.py      (rewrites golden.json next to this file)
Paths cited are relative to numericalEFT/MCIntegration.jl v0.4.2.
"""
import json
import math
import os
import random

M32 = 0xFFFFFFFF


# ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) ----
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return [c0, c1, c2, c3]


def u52(hi, lo):
    """top 52 bits of (hi:lo) times 2^-52, in [0, 1)"""
    return float(((hi << 32) | lo) >> 12) * 2.0 ** -52


def uniforms(seed, it, block, idx, n):
    """the stream layout of DESIGN.md "RNG": key = (seed_lo, seed_hi ^ (iter * 0x9E3779B1 mod 2^32)),
    counter = (idx_lo, idx_hi, pair, block); each call yields two uniforms (words 0:1, then 2:3)."""
    key = [seed & M32, ((seed >> 32) ^ ((it * 0x9E3779B1) & M32)) & M32]
    out = []
    pair = 0
    while len(out) < n:
        o = philox4x32_10([idx & M32, (idx >> 32) & M32, pair, block], key)
        out.append(u52(o[0], o[1]))
        out.append(u52(o[2], o[3]))
        pair += 1
    return out[:n]


# ---- src/distribution/common.jl ----
def locate(acc, p):
    """common.jl:8-36; 1-based; None where the reference raises"""
    n = len(acc)
    if acc[0] > p or acc[-1] <= p:
        return None
    jl, ju = 1, n + 1
    while ju - jl > 1:
        jm = (jl + ju) // 2
        if p < acc[jm - 1]:
            ju = jm
        else:
            jl = jm
    return jl


def smooth(d, factor=6.0):
    """common.jl:43-54"""
    if len(d) <= 1:
        return list(d)
    new = list(d)
    new[0] = (d[0] * (factor + 1) + d[1]) / (factor + 2)
    new[-1] = (d[-1] * (factor + 1) + d[-2]) / (factor + 2)
    for i in range(1, len(d) - 1):
        new[i] = (d[i - 1] + d[i] * factor + d[i + 1]) / (factor + 2)
    return new


def seqsum(v):
    s = 0.0
    for x in v:
        s += x
    return s


def rescale(d, alpha):
    """common.jl:67-83 (output NOT renormalised)"""
    if len(d) == 1:
        return list(d)
    assert all(x > 0 for x in d)
    s = seqsum(d)
    out = [x / s for x in d]
    for i, x in enumerate(out):
        if 0 < x <= 0.99999999:
            out[i] = (-(1 - x) / math.log(x)) ** alpha
    return out


# ---- src/distribution/sampler.jl:293-305 (Continuous create!) ----
def map_continuous(grid, y):
    N = len(grid) - 1
    t = y * N
    iy = int(math.floor(t)) + 1
    dy = t - (iy - 1)
    x = grid[iy - 1] + dy * (grid[iy] - grid[iy - 1])
    prob = 1.0 / (N * (grid[iy] - grid[iy - 1]))
    return iy, x, prob


# ---- src/distribution/variable.jl:206-239 (train! Continuous), :369-382 (Discrete) ----
def train_continuous(grid, hist, alpha):
    avg = rescale(smooth(hist, 6.0), alpha)
    n = len(grid)
    new = [0.0] * n
    new[0], new[-1] = grid[0], grid[-1]
    j, acc = 0, 0.0
    f_ninc = seqsum(avg) / (n - 1)
    for i in range(2, n):
        while acc < f_ninc:
            j += 1
            acc += avg[j - 1]
        acc -= f_ninc
        new[i - 1] = grid[j] - (acc / avg[j - 1]) * (grid[j] - grid[j - 1])
    return avg, new


def train_discrete(hist, alpha):
    d = rescale(list(hist), alpha)
    s = seqsum(d)
    d = [x / s for x in d]
    acc = [0.0] + [seqsum(d[: i + 1]) for i in range(len(d))]
    return d, acc


# ---- src/main.jl:296-320, :322-346; src/statistics.jl:186-220 ----
def mean_std(s, s2, block):
    mean = s / block
    if block > 1:
        v = (s2 / block - mean * mean) / (block - 1)
        return mean, math.sqrt(v) if v > 0 else 0.0
    return mean, 0.0


def do_reweight(rw, visited, gamma, goal=None):
    avg = seqsum(visited)
    rw = list(rw)
    for i, v in enumerate(visited):
        rw[i] *= (avg ** gamma) if v <= 1 else ((avg / v) ** gamma)
    if goal is not None:
        gs = seqsum(goal)
        rw = [r * (g / gs) for r, g in zip(rw, goal)]
    s = seqsum(rw)
    return [r / s for r in rw]


def average(means, stds, init):
    """statistics.jl:186-220, scalar observable; init 1-based"""
    mx = len(means)
    if mx <= init:
        return means[0], stds[0], 0.0
    w = [1.0 / (s + 1.0e-10) ** 2 for s in stds[init - 1:]]
    d = means[init - 1:]
    ws = sum(w)
    mea = sum(di * wi / ws for di, wi in zip(d, w))
    err = 1.0 / math.sqrt(ws)
    chi2 = sum(wi * (di - mea) ** 2 for di, wi in zip(d, w))
    return mea, err, chi2 / ((mx - init + 1) - 1)


# ---- whole solver loops on one Continuous pool (restated from the Julia sources; the two documented deviations of the
# ---- oracle are followed so that the streams line up: initialize! creates the slots 1..maxdof only, and a chain
# ---- consumes one uniform for every rand(rng, 1:n), also when n == 1) ----------------------------------------------
class Stream:
    def __init__(self, seed, it, block, idx):
        self.a = (seed, it, block, idx)
        self.buf = []
        self.n = 0

    def rand(self):
        if self.n >= len(self.buf):
            self.buf = uniforms(*self.a, len(self.buf) + 64)
        u = self.buf[self.n]
        self.n += 1
        return u

    def randint(self, n):
        return min(int(math.floor(self.rand() * n)) + 1, n)


class Pool:
    """Continuous pool: create!/remove!/shift!/swap! of src/distribution/sampler.jl:293-408 (1-based slots)"""

    def __init__(self, grid, nslot):
        self.g, self.N = grid, len(grid) - 1
        self.data, self.gidx, self.prob = [0.0] * (nslot + 2), [0] * (nslot + 2), [0.0] * (nslot + 2)

    def _draw(self, rng):
        y = rng.rand()
        iy = min(int(math.floor(y * self.N)) + 1, self.N)
        dy = y * self.N - (iy - 1)
        return iy, self.g[iy - 1] + dy * (self.g[iy] - self.g[iy - 1])

    def create(self, pos, rng):
        iy, x = self._draw(rng)
        self.data[pos], self.gidx[pos] = x, iy
        self.prob[pos] = 1.0 / (self.N * (self.g[iy] - self.g[iy - 1]))
        return 1.0 / self.prob[pos]

    def remove(self, pos):
        iy = self.gidx[pos]
        return 1 / ((self.g[iy] - self.g[iy - 1]) * self.N)

    def shift(self, pos, rng):
        self.bk = (self.data[pos], self.gidx[pos], self.prob[pos])
        curr = self.gidx[pos]
        iy, x = self._draw(rng)
        self.data[pos], self.gidx[pos] = x, iy
        ratio = (self.g[curr] - self.g[curr - 1]) / (self.g[iy] - self.g[iy - 1])
        self.prob[pos] *= ratio
        return 1.0 / ratio

    def rollback(self, pos):
        self.data[pos], self.gidx[pos], self.prob[pos] = self.bk

    def swap(self, a, b):
        for arr in (self.data, self.gidx, self.prob):
            arr[a], arr[b] = arr[b], arr[a]

    def hist_len(self):
        return self.N

    def accumulate(self, pos, w, hist):  # variable.jl:196-200
        hist[self.gidx[pos] - 1] += w

    def probability(self, d):  # variable.jl:606-619
        p = 1.0
        for pos in range(1, d + 1):
            p *= self.prob[pos]
        return p

    def padding(self, d, maxdof):  # variable.jl:628-641
        p = 1.0
        for pos in range(d + 1, maxdof + 1):
            p *= self.prob[pos]
        return p


class DiscretePair:
    """CompositeVar of two Discrete leaves (variable.jl:272-325, 397-507; sampler.jl:13-22, 37-41, 57-97, 410-462):
    every operation acts on both leaves at the same slot, leaf 0 first; prob = 1.0 * p_leaf0 * p_leaf1."""

    def __init__(self, dists, lowers, nslot):
        self.dist = dists
        self.acc = []
        for d in dists:  # accumulation = [0, sum(d[1:1]), sum(d[1:2]), ...] each prefix recomputed (variable.jl:314-315)
            self.acc.append([0.0] + [seqsum(d[: i + 1]) for i in range(len(d))])
        self.lower = lowers
        self.val = [[0] * (nslot + 2) for _ in dists]
        self.lprob = [[0.0] * (nslot + 2) for _ in dists]
        self.prob = [0.0] * (nslot + 2)
        self.data = self.val  # integrand view: data[leaf][pos]

    def _locate(self, k, u):  # clamped into the table (quirk 15)
        acc = self.acc[k]
        if u < acc[0]:
            return 1
        jl, ju = 1, len(acc) + 1
        while ju - jl > 1:
            jm = (jl + ju) // 2
            if u < acc[jm - 1]:
                ju = jm
            else:
                jl = jm
        return min(jl, len(self.dist[k]))

    def create(self, pos, rng):
        prop, self.prob[pos] = 1.0, 1.0
        for k in range(len(self.dist)):
            g = self._locate(k, rng.rand())
            self.val[k][pos] = self.lower[k] + g - 1
            self.lprob[k][pos] = self.dist[k][g - 1]
            prop *= 1.0 / self.dist[k][g - 1]
            self.prob[pos] *= self.lprob[k][pos]
        return prop

    def remove(self, pos):
        prop = 1.0
        for k in range(len(self.dist)):
            prop *= self.dist[k][self.val[k][pos] - self.lower[k]]
        return prop

    def shift(self, pos, rng):
        self.bk = ([v[pos] for v in self.val], [p[pos] for p in self.lprob], self.prob[pos])
        prop, self.prob[pos] = 1.0, 1.0
        for k in range(len(self.dist)):
            curr = self.val[k][pos] - self.lower[k] + 1
            g = self._locate(k, rng.rand())
            self.val[k][pos] = self.lower[k] + g - 1
            ratio = self.dist[k][g - 1] / self.dist[k][curr - 1]
            self.lprob[k][pos] *= ratio
            prop *= 1.0 / ratio
            self.prob[pos] *= self.lprob[k][pos]
        return prop

    def rollback(self, pos):
        for k in range(len(self.dist)):
            self.val[k][pos], self.lprob[k][pos] = self.bk[0][k], self.bk[1][k]
        self.prob[pos] = self.bk[2]

    def swap(self, a, b):
        for arr in self.val + self.lprob + [self.prob]:
            arr[a], arr[b] = arr[b], arr[a]

    def hist_len(self):
        return sum(len(d) for d in self.dist)

    def accumulate(self, pos, w, hist):  # variable.jl:362-367, 474-478
        off = 0
        for k in range(len(self.dist)):
            hist[off + self.val[k][pos] - self.lower[k]] += w
            off += len(self.dist[k])

    probability = Pool.probability
    padding = Pool.padding


class Prob:
    def __init__(self, make_pool, f, dof, nbrs):
        self.make_pool, self.f, self.DOF, self.NBRS = make_pool, f, dof, nbrs
        self.MAXDOF, self.NI = max(dof), len(dof) - 1


def sphere2(x):  # test/montecarlo.jl:19-24
    r2 = x[1] * x[1] + x[2] * x[2]
    return [1.0 if r2 < 1.0 else 0.0, 1.0 if r2 + x[3] * x[3] < 1.0 else 0.0]


EPS0, TINY_ = 5e-324, 5e-324 * 1e50


def loop_vegas(Q, seed, it, nblock, npb, mf):
    """vegas/montecarlo.jl:117-187"""
    DOF, MAXDOF, NI, sphere2 = Q.DOF, Q.MAXDOF, Q.NI, Q.f
    hist = [0.0] * Q.make_pool().hist_len()
    obs, norm = [], []
    for b in range(nblock):
        o, n = [0.0] * NI, 0.0
        for ne in range(1, npb + 1):
            rng, P = Stream(seed, it, b, ne - 1), Q.make_pool()
            jac = 1.0
            for pos in range(1, MAXDOF + 1):
                P.create(pos, rng)
                jac /= P.prob[pos]
            pad = [P.padding(DOF[i], MAXDOF) if DOF[i] != MAXDOF else 1.0 for i in range(NI)]
            w = sphere2(P.data)
            if ne % mf == 0:
                for i in range(NI):
                    o[i] += w[i] * pad[i] * jac
                n += 1.0
            for i in range(NI):
                for pos in range(1, DOF[i] + 1):
                    P.accumulate(pos, (abs(w[i]) * jac) * (abs(w[i]) * jac), hist)
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist)


def loop_vegasmc(Q, rw, seed, it, nblock, steps, mf):
    """vegas_mc/montecarlo.jl:151-232 + updates.jl:45-106, one chain per block"""
    DOF, MAXDOF, NI, sphere2 = Q.DOF, Q.MAXDOF, Q.NI, Q.f
    hist = [0.0] * Q.make_pool().hist_len()
    obs, norm, visited, prop_c, acc_c, neval = [], [], [0.0] * (NI + 1), 0.0, 0.0, 0
    for b in range(nblock):
        rng, P = Stream(seed, it, b, 0), Q.make_pool()
        for pos in range(1, MAXDOF + 1):
            P.create(pos, rng)
        w = sphere2(P.data)
        pad = [P.padding(DOF[i], MAXDOF) for i in range(NI + 1)]
        cur = rw[NI] * pad[NI]
        for i in range(NI):
            cur += abs(w[i]) * rw[i] * pad[i]
        o, n = [0.0] * NI, 0.0
        for ne in range(1, steps + 1):
            rng.randint(1)  # vi = rand(rng, 1:length(var))
            pos = rng.randint(MAXDOF)
            prop = P.shift(pos, rng)
            if prop > EPS0:
                nw = sphere2(P.data)
                neval += 1
                npad = [P.padding(DOF[i], MAXDOF) for i in range(NI + 1)]
                new = rw[NI] * npad[NI]
                for i in range(NI):
                    new += abs(nw[i]) * rw[i] * npad[i]
                R = prop * new / cur
                prop_c += 1.0
                if rng.rand() < R:
                    acc_c += 1.0
                    w, pad, cur = nw, npad, new
                else:
                    P.rollback(pos)
            for i in range(NI):
                f2 = abs(w[i]) * abs(w[i]) / P.probability(DOF[i])
                wf2 = f2 * pad[i] / cur
                for p in range(1, DOF[i] + 1):
                    P.accumulate(p, wf2, hist)
            if ne % mf == 0 and ne >= steps / 100:
                for i in range(NI):
                    visited[i] += abs(w[i] * pad[i] * rw[i]) / cur
                    o[i] += w[i] * pad[i] / cur
                n += 1.0 * pad[NI] / cur
                visited[NI] += rw[NI] * pad[NI] / cur
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist, visited=visited, propose_cv=prop_c, accept_cv=acc_c, neval=neval)


def loop_mcmc(Q, rw, seed, it, nblock, steps, mf, thermal):
    """mcmc/montecarlo.jl:99-177 + updates.jl:1-147, one chain per block; counters [update][curr][new or vi] (0-based)"""
    DOF, MAXDOF, NI, sphere2, NBRS = Q.DOF, Q.MAXDOF, Q.NI, Q.f, Q.NBRS
    Nd = NI + 1
    hist = [0.0] * Q.make_pool().hist_len()
    obs, norm, visited, neval = [], [], [0.0] * Nd, 0
    prop_c = [[[0.0] * Nd for _ in range(Nd)] for _ in range(3)]
    acc_c = [[[0.0] * Nd for _ in range(Nd)] for _ in range(3)]
    NORM = NI
    for b in range(nblock):
        rng, P = Stream(seed, it, b, 0), Q.make_pool()
        curr = 0
        for _ in range(10000):
            for pos in range(1, MAXDOF + 1):
                P.create(pos, rng)
            weight = sphere2(P.data)[curr]
            prob = abs(weight) * rw[curr]
            if prob > TINY_:
                break
        o, n = [0.0] * NI, 0.0
        nburn = int(math.floor(steps * thermal))
        for i in range(1, steps + nburn + 1):
            visited[curr] += 1
            u = rng.randint(4)
            if u == 1:  # changeIntegrand
                nb = NBRS[curr]
                new = nb[rng.randint(len(nb)) - 1] - 1
                if new != curr:
                    prop = len(nb) / len(NBRS[new])
                    cd, nd = DOF[curr], DOF[new]
                    if cd < nd:
                        for pos in range(cd + 1, nd + 1):
                            prop *= P.create(pos, rng)
                    elif cd > nd:
                        for pos in range(nd + 1, cd + 1):
                            prop *= P.remove(pos)
                    if prop > EPS0:
                        nw = 0.0 if new == NORM else sphere2(P.data)[new]
                        neval += 1
                        nprob = rw[new] if new == NORM else abs(nw) * rw[new]
                        R = prop * nprob / prob
                        prop_c[0][curr][new] += 1.0
                        if rng.rand() < R:
                            acc_c[0][curr][new] += 1.0
                            curr, weight, prob = new, nw, nprob
            elif curr != NORM:
                rng.randint(1)  # vi
                cd = DOF[curr]
                if u == 2:  # swapVariable
                    i1, i2 = rng.randint(cd), rng.randint(cd)
                    if i1 != i2:
                        P.swap(i1, i2)
                        nw = sphere2(P.data)[curr]
                        neval += 1
                        nprob = abs(nw) * rw[curr]
                        R = 1.0 * nprob / prob
                        prop_c[2][curr][0] += 1.0
                        if rng.rand() < R:
                            acc_c[2][curr][0] += 1.0
                            weight, prob = nw, nprob
                        else:
                            P.swap(i1, i2)
                else:  # changeVariable
                    pos = rng.randint(cd)
                    prop = P.shift(pos, rng)
                    if prop > EPS0:
                        nw = sphere2(P.data)[curr]
                        neval += 1
                        nprob = abs(nw) * rw[curr]
                        R = prop * nprob / prob
                        prop_c[1][curr][0] += 1.0
                        if rng.rand() < R:
                            acc_c[1][curr][0] += 1.0
                            weight, prob = nw, nprob
                        else:
                            P.rollback(pos)
            if i % mf == 0 and i >= nburn:
                if curr == NORM:
                    n += 1.0 / rw[NORM]
                else:
                    for pos in range(1, DOF[curr] + 1):
                        P.accumulate(pos, 1.0, hist)
                    o[curr] += weight / prob
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist, visited=visited, propose=prop_c, accept=acc_c, neval=neval)


class FermiKPool:
    """FermiK{D}: create!/remove!/shift!/swap! of src/distribution/sampler.jl:109-281 (1-based slots; data[c][pos] is
    Cartesian component c of K[pos]; slot nslot+1 plays K[end], the rollback cache).  K.prob is not kept: only
    :vegas/:vegasmc read it, and the reference runs FermiK under :mcmc alone (test/bubble_FermiK.jl:2,139-140)."""

    def __init__(self, D, kF, dk, nslot):
        self.D, self.kF, self.dk = D, kF, dk
        self.data = [[kF / math.sqrt(D)] * (nslot + 2) for _ in range(D)]  # variable.jl:13

    def create(self, pos, rng):
        Kamp = self.kF + (rng.rand() - 0.5) * 2.0 * self.dk
        if Kamp <= 0.0:
            return 0.0
        phi = 2 * math.pi * rng.rand()
        if self.D == 3:
            th = math.pi * rng.rand()
            self.data[0][pos] = Kamp * math.cos(phi) * math.sin(th)
            self.data[1][pos] = Kamp * math.sin(phi) * math.sin(th)
            self.data[2][pos] = Kamp * math.cos(th)
            return 2 * self.dk * (2 * math.pi) * math.pi * (math.sin(th) * (Kamp * Kamp))
        self.data[0][pos] = Kamp * math.cos(phi)
        self.data[1][pos] = Kamp * math.sin(phi)
        return 2 * self.dk * (2 * math.pi) * Kamp

    def remove(self, pos):
        k = [d[pos] for d in self.data]
        if self.D == 3:
            Kamp = math.sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2])
        else:
            Kamp = math.sqrt(k[0] * k[0] + k[1] * k[1])
        if not (self.kF - self.dk < Kamp < self.kF + self.dk):
            return 0.0
        if self.D == 3:
            sint = math.sqrt(k[0] * k[0] + k[1] * k[1]) / Kamp
            if sint < 1.0e-15:
                return 0.0
            return 1.0 / (2 * self.dk * (2 * math.pi) * math.pi * sint * (Kamp * Kamp))
        return 1.0 / (2 * self.dk * (2 * math.pi) * Kamp)

    def shift(self, pos, rng):
        self.bk = [d[pos] for d in self.data]
        x = rng.rand()
        if x < 1.0 / 3:
            lam = 1.5
            ratio = 1.0 / lam + rng.rand() * (lam - 1.0 / lam)
            for d in self.data:
                d[pos] *= ratio
            return 1.0 if self.D == 2 else ratio
        if x < 2.0 / 3:
            phi = rng.rand() * (2 * math.pi)
            k = [d[pos] for d in self.data]
            if self.D == 3:
                th = math.acos(1 - 2 * rng.rand())
                Kamp = math.sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2])
                self.data[0][pos] = Kamp * math.cos(phi) * math.sin(th)
                self.data[1][pos] = Kamp * math.sin(phi) * math.sin(th)
                self.data[2][pos] = Kamp * math.cos(th)
            else:
                Kamp = math.sqrt(k[0] * k[0] + k[1] * k[1])
                self.data[0][pos] = Kamp * math.cos(phi)
                self.data[1][pos] = Kamp * math.sin(phi)
            return 1.0
        for d in self.data:
            d[pos] = d[pos] + (rng.rand() - 0.5) * self.dk
        return 1.0

    def rollback(self, pos):
        for d, b in zip(self.data, self.bk):
            d[pos] = b

    def swap(self, a, b):
        for d in self.data:
            d[a], d[b] = d[b], d[a]

    def hist_len(self):
        return 0

    def accumulate(self, pos, w, hist):  # the no-op fallback, variable.jl:551
        pass


class DiscreteOne(DiscretePair):
    """A plain Discrete variable (one leaf); adapt=False keeps it out of the histogram (variable.jl:363)."""

    def __init__(self, dist, lower, nslot, adapt=True):
        super().__init__([dist], [lower], nslot)
        self.adapt = adapt

    def hist_len(self):
        return len(self.dist[0]) if self.adapt else 0

    def accumulate(self, pos, w, hist):
        if self.adapt:
            super().accumulate(pos, w, hist)


def green(tau, w, beta):  # test/bubble_FermiK.jl:41-52
    if tau >= 0.0:
        return math.exp(-w * tau) / (1 + math.exp(-w * beta)) if w > 0.0 else math.exp(w * (beta - tau)) / (1 + math.exp(w * beta))
    return -math.exp(-w * (tau + beta)) / (1 + math.exp(-w * beta)) if w > 0.0 else -math.exp(-w * tau) / (1 + math.exp(w * beta))


def bubble_fermik(ud, D):
    """test/bubble_FermiK.jl:54-75 on vars (T, K, Ext); ud = [kF, beta, me, spin, dim, Qsize, extQ...]"""
    kF, beta, me, spin = ud[0], ud[1], ud[2], ud[3]

    def f(T, K, Ext):
        tau = T.data[1]
        k = [K.data[c][1] for c in range(D)] + [0.0] * (3 - D)
        kq0 = k[0] + ud[6 + Ext.val[0][1] - 1]
        w1 = ((k[0] * k[0] + k[1] * k[1] + k[2] * k[2]) - kF * kF) / (2 * me)
        w2 = ((kq0 * kq0 + k[1] * k[1] + k[2] * k[2]) - kF * kF) / (2 * me)
        phase = 1.0 / math.pow(2 * math.pi, float(D))
        return green(tau, w1, beta) * green(-tau, w2, beta) * spin * phase * math.cos(2 * math.pi * 0 * tau / beta)
    return f


def loop_mcmc_vars(make_vars, f, rw, seed, it, nblock, steps, mf, thermal, nbins_obs):
    """mcmc/montecarlo.jl:99-177 + updates.jl:1-147 for ONE integrand of dof [1, 1, ..., 1] over several variables
    (Nd = 2, neighbours [[2], [1]]), measure = bin by the last (Discrete) variable (test/bubble_FermiK.jl:80-88);
    counters [update][curr][new or vi] with dimension max(Nd, Nv)"""
    V0 = make_vars()
    Nv, Nd, NORM = len(V0), 2, 1
    pa = max(Nd, Nv)
    hoff = [0]
    for v in V0:
        hoff.append(hoff[-1] + v.hist_len())
    hist = [0.0] * hoff[-1]
    obs, norm, visited, neval = [], [], [0.0] * Nd, 0
    prop_c = [[[0.0] * pa for _ in range(Nd)] for _ in range(3)]
    acc_c = [[[0.0] * pa for _ in range(Nd)] for _ in range(3)]
    for b in range(nblock):
        rng, V = Stream(seed, it, b, 0), make_vars()
        curr = 0
        for _ in range(10000):
            for v in V:  # initialize!: variable.jl:576-580, slots 1..maxdof
                v.create(1, rng)
            weight = f(*V)
            prob = abs(weight) * rw[curr]
            if prob > TINY_:
                break
        o, n = [0.0] * nbins_obs, 0.0
        nburn = int(math.floor(steps * thermal))
        for i in range(1, steps + nburn + 1):
            visited[curr] += 1
            u = rng.randint(2 * Nv + 2)  # [changeIntegrand, swapVariable, changeVariable x 2Nv]: montecarlo.jl:117-120
            if u == 1:
                new = 1 - curr  # the only neighbour; rand(rng, [x]) still consumes a draw
                rng.randint(1)
                prop = 1.0  # length(neighbor[curr]) / length(neighbor[new]) = 1 / 1
                for v in V:
                    prop *= v.create(1, rng) if curr == NORM else v.remove(1)
                if prop > EPS0:
                    nw = 0.0 if new == NORM else f(*V)
                    neval += 1
                    nprob = rw[new] if new == NORM else abs(nw) * rw[new]
                    R = prop * nprob / prob
                    prop_c[0][curr][new] += 1.0
                    if rng.rand() < R:
                        acc_c[0][curr][new] += 1.0
                        curr, weight, prob = new, nw, nprob
            elif curr != NORM:
                vi = rng.randint(Nv) - 1
                if u == 2:  # swapVariable: both slots are drawn from 1:1, so nothing is ever swapped
                    rng.randint(1), rng.randint(1)
                else:  # changeVariable
                    v = V[vi]
                    if not (isinstance(v, DiscreteOne) and len(v.dist[0]) == 1):
                        rng.randint(1)
                        prop = v.shift(1, rng)
                        if prop > EPS0:
                            nw = f(*V)
                            neval += 1
                            nprob = abs(nw) * rw[curr]
                            R = prop * nprob / prob
                            prop_c[1][curr][vi] += 1.0
                            if rng.rand() < R:
                                acc_c[1][curr][vi] += 1.0
                                weight, prob = nw, nprob
                            else:
                                v.rollback(1)
            if i % mf == 0 and i >= nburn:
                if curr == NORM:
                    n += 1.0 / rw[NORM]
                else:
                    for k, v in enumerate(V):
                        sub = [0.0] * v.hist_len()
                        v.accumulate(1, 1.0, sub)
                        for j, x in enumerate(sub):
                            hist[hoff[k] + j] += x
                    o[V[-1].val[0][1] - V[-1].lower[0]] += weight / prob
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist, visited=visited, propose=prop_c, accept=acc_c, neval=neval)


def _prob_all(V, dofs):  # variable.jl:606-619: ONE running product over the variables in order
    p = 1.0
    for v, d in zip(V, dofs):
        for pos in range(1, d + 1):
            p *= v.prob[pos]
    return p


def _pad_all(V, dofs, maxdof):  # variable.jl:628-641
    p = 1.0
    for v, d, m in zip(V, dofs, maxdof):
        for pos in range(d + 1, m + 1):
            p *= v.prob[pos]
    return p


def loop_vegas_vars(make_vars, f, DOF, seed, it, nblock, npb, mf, nbins_obs=0):
    """vegas/montecarlo.jl:117-187 over SEVERAL top-level variables; DOF[i][v] incl. the normalisation row.
    nbins_obs > 0: measure = the bubble's (example/bubble.jl:81-84): obs[1][Ext[1]] += relative weight of integrand 1"""
    NI, Nv = len(DOF) - 1, len(DOF[0])
    maxdof = [max(d[v] for d in DOF) for v in range(Nv)]
    V0 = make_vars()
    hoff = [0]
    for v in V0:
        hoff.append(hoff[-1] + v.hist_len())
    hist = [0.0] * hoff[-1]
    obs, norm = [], []
    for b in range(nblock):
        o, n = [0.0] * (nbins_obs or NI), 0.0
        for ne in range(1, npb + 1):
            rng, V = Stream(seed, it, b, ne - 1), make_vars()
            jac = 1.0
            for v, m in zip(V, maxdof):
                for pos in range(1, m + 1):
                    v.create(pos, rng)
                    jac /= v.prob[pos]
            pad = [1.0 if DOF[i] == maxdof else _pad_all(V, DOF[i], maxdof) for i in range(NI)]
            w = f(*V)
            if ne % mf == 0:
                if nbins_obs:
                    o[V[-1].val[0][1] - V[-1].lower[0]] += w[0] * pad[0] * jac
                else:
                    for i in range(NI):
                        o[i] += w[i] * pad[i] * jac
                n += 1.0
            for k, v in enumerate(V):
                sub = [0.0] * v.hist_len()
                for i in range(NI):
                    for pos in range(1, DOF[i][k] + 1):
                        v.accumulate(pos, (abs(w[i]) * jac) * (abs(w[i]) * jac), sub)
                for j, x in enumerate(sub):
                    hist[hoff[k] + j] += x
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist)


def loop_vegasmc_vars(make_vars, f, DOF, rw, seed, it, nblock, steps, mf, nbins_obs=0):
    """vegas_mc/montecarlo.jl:151-232 + updates.jl:45-106 over several top-level variables, one chain per block;
    propose/accept[2, 1, vi] come back as lists over vi"""
    NI, Nv = len(DOF) - 1, len(DOF[0])
    maxdof = [max(d[v] for d in DOF) for v in range(Nv)]
    V0 = make_vars()
    hoff = [0]
    for v in V0:
        hoff.append(hoff[-1] + v.hist_len())
    hist = [0.0] * hoff[-1]
    obs, norm, visited, neval = [], [], [0.0] * (NI + 1), 0
    prop_c, acc_c = [0.0] * Nv, [0.0] * Nv
    for b in range(nblock):
        rng, V = Stream(seed, it, b, 0), make_vars()
        for v, m in zip(V, maxdof):
            for pos in range(1, m + 1):
                v.create(pos, rng)
        w = f(*V)
        pad = [_pad_all(V, DOF[i], maxdof) for i in range(NI + 1)]
        cur = rw[NI] * pad[NI]
        for i in range(NI):
            cur += abs(w[i]) * rw[i] * pad[i]
        o, n = [0.0] * (nbins_obs or NI), 0.0
        for ne in range(1, steps + 1):
            vi = rng.randint(Nv) - 1
            v = V[vi]
            if not (isinstance(v, DiscreteOne) and len(v.dist[0]) == 1) and maxdof[vi] > 0:
                pos = rng.randint(maxdof[vi])
                prop = v.shift(pos, rng)
                if prop > EPS0:
                    nw = f(*V)
                    neval += 1
                    npad = [_pad_all(V, DOF[i], maxdof) for i in range(NI + 1)]
                    new = rw[NI] * npad[NI]
                    for i in range(NI):
                        new += abs(nw[i]) * rw[i] * npad[i]
                    R = prop * new / cur
                    prop_c[vi] += 1.0
                    if rng.rand() < R:
                        acc_c[vi] += 1.0
                        w, pad, cur = nw, npad, new
                    else:
                        v.rollback(pos)
            for i in range(NI):
                f2 = abs(w[i]) * abs(w[i]) / _prob_all(V, DOF[i])
                wf2 = f2 * pad[i] / cur
                for k, vv in enumerate(V):
                    sub = [0.0] * vv.hist_len()
                    for p_ in range(1, DOF[i][k] + 1):
                        vv.accumulate(p_, wf2, sub)
                    for j, x in enumerate(sub):
                        hist[hoff[k] + j] += x
            if ne % mf == 0 and ne >= steps / 100:
                for i in range(NI):
                    visited[i] += abs(w[i] * pad[i] * rw[i]) / cur
                    if nbins_obs:
                        o[V[-1].val[0][1] - V[-1].lower[0]] += w[i] * pad[i] / cur
                    else:
                        o[i] += w[i] * pad[i] / cur
                n += 1.0 * pad[NI] / cur
                visited[NI] += rw[NI] * pad[NI] / cur
        obs.append(o)
        norm.append(n)
    return dict(obs=obs, norm=norm, hist=hist, visited=visited, propose_cv=prop_c, accept_cv=acc_c, neval=neval)


def main():
    rnd = random.Random(20261017)
    G = {}
    G["philox_kat"] = [
        dict(ctr=[0, 0, 0, 0], key=[0, 0], out=[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        dict(ctr=[M32] * 4, key=[M32] * 2, out=[0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        dict(ctr=[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], key=[0xA4093822, 0x299F31D0],
             out=[0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for k in G["philox_kat"]:  # the restatement above must reproduce the published vectors
        assert philox4x32_10(k["ctr"], k["key"]) == k["out"], k
    G["uniforms"] = [dict(seed=s, iter=it, block=b, idx=i, u=uniforms(s, it, b, i, 7))
                     for s, it, b, i in [(1234, 0, 0, 0), (1234, 3, 15, (1 << 33) + 7), (0xDEADBEEFCAFEF00D, 9, 2, 12345)]]

    # locate: test/utility.jl:1-10 plus clamping edges
    grid = [0.0, 0.1, 0.3, 0.5]
    eps = 2.0 ** -52
    pts = [eps, 0.5 - eps, grid[0], 0.05, 0.2, 0.31, 0.1, 0.3, 0.5, -0.1]
    G["locate"] = dict(acc=grid, p=pts, idx=[locate(grid, p) for p in pts])
    assert G["locate"]["idx"][:6] == [1, 3, 1, 1, 2, 3]

    # map: SURVEY 8(c) hand cases + random grids
    maps = []
    g3 = [0.0, 0.1, 0.4, 1.0]
    for y in (0.0, 0.5, 0.75, 1 - 2.0 ** -53, 1.0 / 3, 2.0 / 3):
        iy, x, p = map_continuous(g3, y)
        maps.append(dict(grid=g3, y=y, iy=iy, x=x, prob=p))
    for _ in range(6):
        n = rnd.choice([5, 17, 1000])
        cuts = sorted(rnd.uniform(-5, 5) for _ in range(n - 2))
        g = [-5.0] + cuts + [5.0]
        for _ in range(4):
            y = rnd.random()
            iy, x, p = map_continuous(g, y)
            maps.append(dict(grid=g, y=y, iy=iy, x=x, prob=p))
    G["map_continuous"] = maps

    # smooth / rescale / train!
    tc = []
    hand = [([0.0, 0.1, 0.4, 1.0], [4.0, 1.0, 0.25], 2.0), ([0.0, 0.25, 0.5, 0.75, 1.0], [1.0, 2.0, 3.0, 10.0], 2.0)]
    for grid_, hist, alpha in hand:
        avg, new = train_continuous(grid_, hist, alpha)
        tc.append(dict(grid=grid_, hist=hist, alpha=alpha, smooth=smooth(hist), rescaled=avg, newgrid=new))
    for n, alpha in [(8, 2.0), (64, 1.5), (1000, 2.0), (1000, 3.0)]:
        lo, hi = rnd.uniform(-3, 0), rnd.uniform(1, 7)
        g = [lo + (hi - lo) * i / (n - 1) for i in range(n)]
        g[-1] = hi
        hist = [1e-10 + rnd.random() ** 6 * 10 ** rnd.uniform(-3, 3) for _ in range(n - 1)]
        avg, new = train_continuous(g, hist, alpha)
        tc.append(dict(grid=g, hist=hist, alpha=alpha, smooth=smooth(hist), rescaled=avg, newgrid=new))
    G["train_continuous"] = tc
    td = []
    for hist, alpha in [([1.0, 2.0, 3.0], 2.0), ([5.0], 2.0), ([1e-10, 7.5, 0.3, 2.0, 9.0, 0.01], 1.5),
                        ([rnd.random() for _ in range(33)], 2.0)]:
        d, acc = train_discrete(hist, alpha)
        td.append(dict(hist=hist, alpha=alpha, dist=d, accum=acc))
    G["train_discrete"] = td

    # _mean_std against the plain sample statistics (test/statistics.jl:22-46)
    ms = []
    for block in (2, 16, 64):
        m = [rnd.gauss(1.0, 0.3) for _ in range(block)]
        s, s2 = seqsum(m), seqsum([x * x for x in m])
        mean, std = mean_std(s, s2, block)
        mu = sum(m) / block
        sample_std = math.sqrt(sum((x - mu) ** 2 for x in m) / (block - 1)) / math.sqrt(block)
        ms.append(dict(block=block, sum=s, sqsum=s2, mean=mean, std=std, sample_mean=mu, sample_std_over_sqrtB=sample_std))
    G["mean_std"] = ms

    # doReweight!: the fixed point of test/mpi_test.jl:148-168 and generic cases
    rw = [0.25] * 4
    seq = []
    cur = [1.0, 1.0, 1.0, 1.0]
    cur = [c / 4 for c in cur]
    for _ in range(5):
        cur = do_reweight(cur, [1.0, 2.0, 3.0, 4.0], 1.0, [1.0, 2.0, 3.0, 4.0])
        seq.append(cur)
    G["reweight"] = dict(
        mpi_test=dict(visited=[1.0, 2.0, 3.0, 4.0], goal=[1.0, 2.0, 3.0, 4.0], gamma=1.0, start=[0.25] * 4, rounds=seq),
        generic=[dict(rw=r, visited=v, gamma=g, goal=goal, out=do_reweight(r, v, g, goal))
                 for r, v, g, goal in [([0.5, 0.5], [10.0, 30.0], 1.0, None), ([0.2, 0.3, 0.5], [0.5, 200.0, 50.0], 0.5, None),
                                       ([0.2, 0.3, 0.5], [7.0, 200.0, 50.0], 2.0, [1.0, 1.0, 4.0])]])

    # average over iterations (statistics.jl:186-220)
    means = [rnd.gauss(2.0, 0.01) for _ in range(10)]
    stds = [abs(rnd.gauss(0.01, 0.002)) for _ in range(10)]
    G["average"] = [dict(means=means, stds=stds, init=i, out=list(average(means, stds, i))) for i in (1, 2, 10)]

    # whole solver loops: sphere2 (dof [[2],[3]]) on a hand-made non-uniform 8-increment grid
    lgrid = [0.0, 0.07, 0.19, 0.3, 0.46, 0.6, 0.71, 0.88, 1.0]
    lrw = [0.5, 0.3, 0.2]
    Q = Prob(lambda: Pool(lgrid, 3), sphere2, [2, 3, 0], [[3, 2], [1], [1]])  # neighbours: configuration.jl:201-209, Nd = 3
    G["loops"] = dict(grid=lgrid, reweight=lrw, seed=4242, iter=3, n_block=2,
                      vegas=dict(npb=60, measurefreq=2, out=loop_vegas(Q, 4242, 3, 2, 60, 2)),
                      vegasmc=dict(steps=300, measurefreq=1, out=loop_vegasmc(Q, lrw, 4242, 3, 2, 300, 1)),
                      mcmc=dict(steps=300, measurefreq=2, thermal_ratio=0.1, out=loop_mcmc(Q, lrw, 4242, 3, 2, 300, 2, 0.1)))
    # the same loops on Discrete([(1,3),(1,4)]) (a CompositeVar of two Discrete leaves), f = 1 (test/montecarlo.jl:113-120)
    # with non-uniform distributions: pins locate / the Discrete index draws / the composite probability product
    d1, d2 = [0.2, 0.5, 0.3], [0.1, 0.4, 0.15, 0.35]
    QD = Prob(lambda: DiscretePair([d1, d2], [1, 1], 1), lambda x: [1.0], [1, 0], [[2], [1]])
    drw = [0.6, 0.4]
    G["loops_discrete"] = dict(dist=[d1, d2], reweight=drw, seed=777, iter=1, n_block=2,
                               vegas=dict(npb=80, measurefreq=1, out=loop_vegas(QD, 777, 1, 2, 80, 1)),
                               vegasmc=dict(steps=250, measurefreq=1, out=loop_vegasmc(QD, drw, 777, 1, 2, 250, 1)),
                               mcmc=dict(steps=250, measurefreq=1, thermal_ratio=0.2, out=loop_mcmc(QD, drw, 777, 1, 2, 250, 1, 0.2)))

    # :mcmc on vars (T::Continuous, K::FermiK{D}, Ext::Discrete nonadaptive), test/bubble_FermiK.jl, D = 3 and 2:
    # pins the FermiK proposals (create!/remove!/shift!) and the multi-variable update mix
    G["loops_fermik"] = []
    for D, seedF in ((3, 31337), (2, 2718)):
        spin, me, beta0, rs, Qsize = 2, 0.5, 25.0, 1.0, 4
        kF = (9 * math.pi / (2 * spin)) ** (1 / 3) / rs if D == 3 else math.sqrt(4 / spin) / rs
        beta = beta0 / (kF ** 2 / 2 / me)
        extQ = [((1 - i / (Qsize - 1)) * 0.0 * kF + (i / (Qsize - 1)) * 1.5 * kF) for i in range(Qsize)]
        ud = [kF, beta, me, float(spin), float(D), float(Qsize)] + extQ
        tgrid = [0.0, 0.04 * beta, 0.15 * beta, 0.5 * beta, 0.87 * beta, beta]
        dk = 0.2 * kF
        frw = [0.7, 0.3]
        mk = lambda: [Pool(tgrid, 1), FermiKPool(D, kF, dk, 1), DiscreteOne([0.25] * 4, 1, 1, adapt=False)]
        G["loops_fermik"].append(dict(dim=D, userdata=ud, tgrid=tgrid, kF=kF, dk=dk, reweight=frw, seed=seedF, iter=2, n_block=2,
                                      steps=400, measurefreq=1, thermal_ratio=0.1,
                                      out=loop_mcmc_vars(mk, bubble_fermik(ud, D), frw, seedF, 2, 2, 400, 1, 0.1, Qsize)))

    # :vegas and :vegasmc over TWO top-level variables with padding across them: var = (Continuous pool on the hand grid,
    # Discrete(1,4) with a non-uniform distribution), sphere2 on the pool, dof = [[2, 1], [3, 0]]
    vdist = [0.1, 0.4, 0.15, 0.35]
    vdof = [[2, 1], [3, 0], [0, 0]]
    mkv = lambda: [Pool(lgrid, 3), DiscreteOne(vdist, 1, 1, adapt=True)]
    fv = lambda P, D: sphere2(P.data)
    G["loops_vars"] = dict(grid=lgrid, dist=vdist, reweight=lrw, seed=909, iter=2, n_block=2,
                           vegas=dict(npb=70, measurefreq=1, out=loop_vegas_vars(mkv, fv, vdof, 909, 2, 2, 70, 1)),
                           vegasmc=dict(steps=300, measurefreq=1, out=loop_vegasmc_vars(mkv, fv, vdof, lrw, 909, 2, 2, 300, 1)))

    # the bubble's measure (bin the one integrand by the last, non-adaptive Discrete variable) under :vegas / :vegasmc:
    # var = (Continuous pool, Discrete(1,4; adapt=false)), f = x, dof = [[1, 1]]
    bdof = [[1, 1], [0, 0]]
    mkb = lambda: [Pool(lgrid, 1), DiscreteOne([0.25] * 4, 1, 1, adapt=False)]
    fb = lambda P, D: [P.data[1]]
    brw = [0.8, 0.2]
    G["loops_binned"] = dict(grid=lgrid, reweight=brw, seed=5150, iter=4, n_block=2,
                             vegas=dict(npb=90, measurefreq=1, out=loop_vegas_vars(mkb, fb, bdof, 5150, 4, 2, 90, 1, nbins_obs=4)),
                             vegasmc=dict(steps=300, measurefreq=1,
                                          out=loop_vegasmc_vars(mkb, fb, bdof, brw, 5150, 4, 2, 300, 1, nbins_obs=4)))

    # complex integrands (Configuration(type=ComplexF64)): TestComplex2 of test/montecarlo.jl:172-198, f = (x, x^2 im),
    # dof = [[1], [1]], through the same three loops — Python's complex arithmetic plays Julia's: abs(z) = hypot(re, im),
    # z * real and z / real component by component.  Stored as (re, im) pairs in the library's observable layout
    # [block][integrand][re, im] (include/mcib200.h, MCI_WEIGHT_COMPLEX)
    QC = Prob(lambda: Pool(lgrid, 1), lambda x: [x[1] + 0j, x[1] * x[1] * 1j], [1, 1, 0], [[3, 2], [1], [1]])
    crw = [0.4, 0.35, 0.25]

    def split(o):
        o = dict(o)
        o["obs"] = [[c for z in blk for c in (complex(z).real, complex(z).imag)] for blk in o["obs"]]
        return o

    G["loops_complex"] = dict(grid=lgrid, reweight=crw, seed=1618, iter=5, n_block=2,
                              vegas=dict(npb=64, measurefreq=1, out=split(loop_vegas(QC, 1618, 5, 2, 64, 1))),
                              vegasmc=dict(steps=300, measurefreq=1, out=split(loop_vegasmc(QC, crw, 1618, 5, 2, 300, 1))),
                              mcmc=dict(steps=300, measurefreq=1, thermal_ratio=0.1, out=split(loop_mcmc(QC, crw, 1618, 5, 2, 300, 1, 0.1))))

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(out, "w") as f:
        json.dump(G, f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

"""Host post-processing mirrors: _mean_std (main.jl:296-320), average/Result (statistics.jl:16-63,186-220),
report (statistics.jl:137-172).  Tiny, runs once per iteration on block means the device produced."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


def mean_std(obs_sum, obs_sq_sum, block):
    """_mean_std: mean = Σm/B, std = sqrt(max(0,(Σm²/B − mean²)/(B−1)))."""
    obs_sum = np.asarray(obs_sum, dtype=np.float64)
    obs_sq_sum = np.asarray(obs_sq_sum, dtype=np.float64)
    mean = obs_sum / block
    if block > 1:
        v = (obs_sq_sum / block - mean ** 2) / (block - 1)
        std = np.sqrt(np.where(v < 0.0, 0.0, v))
    else:
        std = np.zeros_like(obs_sq_sum)
    return mean, std


def average(means, stds, init=1, max_=None):
    """statistics.jl:186-220 for one (possibly vector) observable.  init is 1-based."""
    if max_ is None:
        max_ = len(means)
    if max_ <= init:
        return means[0], stds[0], np.zeros_like(means[0])
    if np.iscomplexobj(np.asarray(means[-1])):  # real and imaginary parts are averaged separately (statistics.jl:208-215)
        part = lambda f, xs: [f(np.asarray(x)) for x in xs]
        mr, er, cr = average(part(np.real, means), part(np.real, stds), init, max_)
        mi, ei, ci = average(part(np.imag, means), part(np.imag, stds), init, max_)
        return mr + 1j * mi, er + 1j * ei, cr + 1j * ci
    data = [np.asarray(means[i]) for i in range(init - 1, max_)]
    weight = [1.0 / (np.asarray(stds[i]) + 1.0e-10) ** 2 for i in range(init - 1, max_)]
    wsum = sum(weight)
    mea = sum(d * w / wsum for d, w in zip(data, weight))
    err = 1.0 / np.sqrt(wsum)
    chi2 = sum(w * (d - mea) ** 2 for d, w in zip(data, weight))
    return mea, err, chi2 / ((max_ - init + 1) - 1)


@dataclass
class Result:
    """statistics.jl:16-63."""
    mean: list
    stdev: list
    chi2: list
    neval: int
    ignore: int
    config: object
    iterations: list = field(default_factory=list)  # [(mean[], std[], snapshot), ...]
    timing: dict = field(default_factory=dict)  # integrate(): seconds spent in create / comm_init / iterations (not in the reference)

    @classmethod
    def from_history(cls, history, ignore, config, neval):
        init = ignore + 1
        assert len(history) > 0
        N = len(history[-1][0])
        res = [average([h[0][o] for h in history], [h[1][o] for h in history], init=init, max_=len(history))
               for o in range(N)]
        return cls([r[0] for r in res], [r[1] for r in res], [r[2] for r in res], neval, ignore, config, history)

    def __getitem__(self, i):
        return self.mean[i], self.stdev[i], self.chi2[i]

    def dof(self):
        return (len(self.iterations) - (self.ignore + 1) + 1) - 1


def report(result, pick=None, name=None, verbose=0, io=None):
    """statistics.jl:137-172: per-iteration table + final estimate; report(config) -> report_config."""
    import sys

    if not isinstance(result, Result):
        return report_config(result, io=io)

    io = io or sys.stdout
    N = len(result.mean)
    for o in range(N):
        label = f"Integral {o + 1}" if name is None else str(name[o])
        bar = "=" * 36 + f"  {label:^16}  " + "=" * 36
        print(bar, file=io)
        print(f"{'iter':>6}  {'integral':>28}  {'wgt average':>28}  {'chi2/dof':>10}", file=io)
        print("-" * len(bar), file=io)
        for i in range(len(result.iterations)):
            m, s = result.iterations[i][0][o], result.iterations[i][1][o]
            if i < result.ignore:
                am, ae, c2 = m, s, 0.0
                tag = "ignore"
            else:
                am, ae, c2 = average([h[0][o] for h in result.iterations], [h[1][o] for h in result.iterations],
                                     init=result.ignore + 1, max_=i + 1)
                tag = f"{i + 1:6d}"
            pm = (lambda x: np.asarray(x).ravel()[0 if pick is None else pick])
            print(f"{tag:>6}  {pm(m):16.8g} ± {pm(s):<10.4g}  {pm(am):16.8g} ± {pm(ae):<10.4g}  {pm(c2):10.4f}", file=io)
        print("-" * len(bar), file=io)
        m, s, c2 = result.mean[o], result.stdev[o], result.chi2[o]
        print(f"{label} = {np.asarray(m).ravel()} ± {np.asarray(s).ravel()}   (reduced chi2 = {np.asarray(c2).ravel()})", file=io)
    return None


def report_config(config, io=None):
    """report(config::Configuration) (configuration.jl:345-464): proposal / acceptance tables of the three updates,
    visit counts and reweight factors, from the counters the kernels hand back (propose/accept[3, Nd, max(Nd, Nv)])."""
    import sys

    io = io or sys.stdout
    neval = max(float(config.neval), 1.0)
    propose, accept = np.asarray(config.propose), np.asarray(config.accept)
    Nd = len(config.visited)
    bar = "-" * 85
    print("\n===========================  Configuration  =========================================", file=io)
    print(bar, file=io)
    print(f"Integral num = {config.N}, dof = {config.dof}, with variables:", file=io)
    for vi, v in enumerate(config.var):
        print(f"{vi + 1}. {v}", file=io)
    print(bar, file=io)
    row = lambda label, p, a: print(f"{label:<22s} {p / neval * 100.0:11.6f}% {a / neval * 100.0:11.6f}% {a / p if p else float('nan'):12.6f}", file=io)
    print(f"{'ChangeIntegrand':<20s} {'Proposed':>12s} {'Accepted':>12s} {'Ratio  ':>12s}", file=io)
    for n in config.neighbor[Nd - 1]:
        row(f"Norm -> {n:2d}:", propose[0, Nd - 1, n - 1], accept[0, Nd - 1, n - 1])
    for idx in range(1, Nd):
        for n in config.neighbor[idx - 1]:
            row(f"  {idx} ->Norm:" if n == Nd else f"  {idx} -> {n:2d}:", propose[0, idx - 1, n - 1], accept[0, idx - 1, n - 1])
    print(bar, file=io)
    kinds = {"Continuous": "Continuous", "Discrete": "Discrete", "CompositeVar": "Composite", "FermiK": "FermiK"}
    for title, u in (("ChangeVariable", 1), ("SwapVariable", 2)):
        print(f"{title:<20s} {'Proposed':>12s} {'Accepted':>12s} {'Ratio  ':>12s}", file=io)
        for idx in range(1, Nd):
            for vi, v in enumerate(config.var):
                row(f"  {idx:2d} / {kinds.get(type(v).__name__, type(v).__name__):<11s}:", propose[u, idx - 1, vi], accept[u, idx - 1, vi])
        print(bar, file=io)
    print("Integrand            Visited      ReWeight", file=io)
    print(f"  Norm   :     {int(config.visited[-1]):12d} {config.reweight[-1]:12.6f}", file=io)
    for idx in range(1, Nd):
        print(f"  Order{idx:2d}:     {int(config.visited[idx - 1]):12d} {config.reweight[idx - 1]:12.6f}", file=io)
    print(bar, file=io)
    print(f"Integrand evaluation = {int(config.neval)}\n", file=io)


def do_reweight(reweight, visited, gamma, goal=None):
    """doReweight! (main.jl:322-346) — host copy for tests; the engine runs the device/oracle version."""
    reweight = np.array(reweight, dtype=np.float64)
    avgstep = float(np.sum(visited))
    for i, v in enumerate(visited):
        if v <= 1:
            reweight[i] *= avgstep ** gamma
        else:
            reweight[i] *= (avgstep / v) ** gamma
    if goal is not None:
        goal = np.asarray(goal, dtype=np.float64)
        reweight *= goal / goal.sum()
    return reweight / reweight.sum()

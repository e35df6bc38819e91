"""integrate(f; var, dof, solver, neval, ...) — the reference's driver (src/main.jl:71-218) with the
block loop + statistics reduction of one iteration replaced by ONE call into the C ABI
(mci_run_iteration) and train!/doReweight! replaced by mci_refine / mci_reweight on the device."""
from __future__ import annotations

import time

import numpy as np

from . import _abi
from .configuration import Configuration, maps_from_config, maps_to_config, problem_from_config
from .integrands import resolve
from .statistics import Result, mean_std, report

_SOLVERS = {"vegas": _abi.SOLVER_VEGAS, "vegasmc": _abi.SOLVER_VEGASMC, "mcmc": _abi.SOLVER_MCMC}
_MEASURES = {None: _abi.MEASURE_DEFAULT, "default": _abi.MEASURE_DEFAULT,
             "bin_by_discrete": _abi.MEASURE_BIN_BY_DISCRETE, "user": _abi.MEASURE_USER}

# concurrent chains per block for the chain solvers when the caller does not say (SURVEY §7 hard part 3), and the
# shortest chain the default will cut a block into: a chain starts from a draw of the VEGAS map, not of the target
# density, so chains much shorter than the equilibration time bias the estimate
DEFAULT_WALKERS = 4096
MIN_CHAIN_STEPS = 512
# unmeasured steps every chain makes first when a block is cut into several chains and the caller does not say:
# measured on C4 (tools/c4_bias_probe.py, against the finite-temperature Lindhard function) 1000 burn-in steps take
# 1525-step chains from 3-8 sigma off to < 1 sigma, and 4000 do the same for 190-step chains
DEFAULT_CHAIN_BURNIN = 1024


def default_chain_burnin(walkers, chain_burnin):
    """None -> 0 for the reference's decomposition (one chain per block), DEFAULT_CHAIN_BURNIN otherwise"""
    if chain_burnin is not None:
        return int(chain_burnin)
    return 0 if walkers <= 1 else DEFAULT_CHAIN_BURNIN


def _standardize_block(neval, nblock, nworker=1):
    """main.jl:220-234"""
    assert neval > nblock, f"neval={neval} should be larger than nblock = {nblock}"
    if nblock > nworker:
        nblock = (nblock // nworker) * nworker
    else:
        nblock = nworker
    return int(neval // nblock), int(nblock)


def split_observables(config, flat):
    """flat [obs_total doubles] -> list per integrand (scalar or vector), like config.observable.  Complex
    observables arrive as interleaved (re, im) doubles (include/mcib200.h, MCI_WEIGHT_COMPLEX)."""
    flat = np.asarray(flat, dtype=np.float64)
    if config.is_complex():
        flat = flat[0::2] + 1j * flat[1::2]
    out, o = [], 0
    for n, proto in zip(config.obs_len(), config.observable):
        if isinstance(proto, np.ndarray):
            out.append(np.array(flat[o:o + n]))
        else:
            out.append(complex(flat[o]) if config.is_complex() else float(flat[o]))
        o += n
    return out


def block_statistics(stats, config, block):
    """main.jl:275-287 + _mean_std: per-block normalised means, their sum and squared sum.  A complex observable is
    handled through its (re, im) doubles: the reference squares and square-roots the two parts separately
    (main.jl:281,285,305-308), which is exactly the element-wise real arithmetic below."""
    m = stats.block_obs / stats.block_norm[:, None]
    mean, std = mean_std(m.sum(axis=0), (m * m).sum(axis=0), block)
    return split_observables(config, mean), split_observables(config, std)


def _warn_nonfinite(engine, solver, it):
    """debug = true: the solver loops warn about non-finite integrand values (vegas/montecarlo.jl:176-178,
    vegas_mc/montecarlo.jl:189-191, mcmc/montecarlo.jl:142-144).  The kernels count them; the host prints."""
    import warnings

    n = engine.nonfinite_count() if hasattr(engine, "nonfinite_count") else 0
    if n:
        what = "abs of the integrand" if solver == "vegas" else "integrand probability"
        warnings.warn(f"{what} is not finite at {n} evaluation(s) of iteration {it + 1}", RuntimeWarning, stacklevel=3)
    return n


def drive(engine, config, *, solver, niter, block, gamma=1.0, adapt=True, reweight_goal=None, ignore=None,
          print_level=-1, run_iteration=None, first_iter=0, debug=False, fused=False):
    """The `for iter in 1:niter` loop of main.jl:142-207 on top of an engine.  Engine-agnostic so the
    tests can run the CPU oracle through the same host logic.  fused=True (CudaEngine only) runs the whole loop —
    iterations, doReweight!, train! — as one mci_run_iterations call with no host round trip in between."""
    if ignore is None:
        ignore = 1 if adapt else 0
    run = run_iteration or engine.run_iteration
    history = []
    neval_total = 0
    chain = solver in ("mcmc", "vegasmc")
    fused_stats = None
    if fused and run_iteration is None and hasattr(engine, "run_iterations") and not debug:
        fused_stats = engine.run_iterations(first_iter, niter, gamma=gamma, goal=reweight_goal, do_reweight=chain,
                                            do_refine=bool(adapt))
    for it in range(niter):
        stats = fused_stats[it] if fused_stats is not None else run(first_iter + it)
        if debug:
            _warn_nonfinite(engine, solver, it)
        neval_total += stats.neval
        config.neval = stats.neval
        config.visited = stats.visited
        config.propose, config.accept = stats.propose, stats.accept
        config.normalization = float(stats.block_norm.sum())
        if fused_stats is None:
            if chain:
                engine.reweight(gamma, reweight_goal)  # main.jl:183
            if adapt:
                engine.refine()  # main.jl:193-198
        mean, std = block_statistics(stats, config, block)
        history.append((mean, std, {"neval": stats.neval, "hist": stats.hist}))
    maps_to_config(config, *engine.get_maps())  # `configs[1]` keeps the trained grid (main.jl:204)
    result = Result.from_history(history, ignore, config, neval_total)
    if print_level >= 0:
        report(result)
    return result


def integrate(integrand, *, solver="vegasmc", config=None, neval=1e4, niter=10, block=16, verbose=-1, gamma=1.0,
              adapt=True, debug=False, reweight_goal=None, ignore=None, measure=None, measurefreq=1,
              thermal_ratio=0.1, inplace=False, parallel="nothread", print=-1, walkers=None, device=0,
              hist_policy=0, comm=None, jit=0, chain_burnin=None, idx=1, fused=None, **kwargs):
    """Drop-in for MCIntegration.integrate (main.jl:71-90) on one B200.

    `integrand` is a built-in device integrand (mcintegration_jl_b200.integrands) or CUDA C++ source
    (integrands.user_integrand, compiled by NVRTC into a kernel specialised to the problem) instead of a closure;
    `jit` = 0 auto / 1 always specialise with NVRTC / 2 ahead-of-time kernels only;
    `chain_burnin` = extra unmeasured steps at the start of every chain (0 = the reference; default 1024 when a
    block is cut into several chains);
    `measure` is None, "bin_by_discrete" or integrands.user_measure(source).  `walkers` = concurrent chains per block for :vegasmc/:mcmc (default: chosen by the library to fill the GPU; the
    choice depends on the number of GPUs, so pass it explicitly for results that are identical across GPU counts).
    `comm` = (rank, world, unique_id_bytes) attaches an NCCL communicator (one process per GPU);
    `idx` = the integrand every :mcmc chain starts in (the kwarg of MCMC.montecarlo, mcmc/montecarlo.jl:76);
    `debug=True` warns, like the reference's loops, when the kernels met non-finite integrand values;
    `fused` = run all iterations device-side in one library call (the default when debug is off).
    Everything else keeps the reference's meaning; `neval` is per iteration.  Like the reference (main.jl:136) every
    call draws the seed of its sample streams from config.rng, so repeated calls on one Configuration — a warm restart
    with config=result.config — use fresh, independent samples.
    """
    from .engine import CudaEngine

    solver = str(solver).lstrip(":")
    if solver not in _SOLVERS:
        raise ValueError(f"Solver {solver} is not supported!")  # main.jl:263
    print_level = max(print, verbose)
    f = resolve(integrand)
    from .integrands import UserMeasure
    measure_source = None
    if isinstance(measure, UserMeasure):  # measure(x, obs, relw, config) as CUDA C++ source
        measure_source, measure = measure.source, "user"
    if config is None:
        config = Configuration(**kwargs)
    for i, m in enumerate(config.maxdof):
        assert m + 2 <= config.var[i].poolsize(), "maxdof should be less than the length of var"  # main.jl:99-101
    world = comm[1] if comm else 1
    nevalperblock, block = _standardize_block(int(neval), int(block), world)
    assert block % world == 0  # main.jl:122
    if solver == "vegas":
        W = 1
    elif walkers:
        W = int(walkers)
        nevalperblock = (nevalperblock // W) * W
    else:
        W = 0  # the library fills the GPU with whole CTAs of chains of >= MIN_CHAIN_STEPS steps (engine.chain_info())
    spec = problem_from_config(config, f, solver=_SOLVERS[solver], measure_kind=_MEASURES[measure], n_block=block,
                               neval_per_block=nevalperblock, measurefreq=measurefreq, walkers_per_block=W,
                               thermal_ratio=thermal_ratio, device=device, hist_policy=hist_policy, jit=jit, chain_burnin=default_chain_burnin(W if W else DEFAULT_WALKERS, chain_burnin),
                               measure_source=measure_source, seed=config.next_seed(), mcmc_start_idx=int(idx))
    t0 = time.perf_counter()
    engine = CudaEngine(spec)
    try:
        engine.set_maps(*maps_from_config(config))  # warm restart: config=result.config (main.jl:95-97)
        t1 = time.perf_counter()
        if comm:
            engine.comm_init(comm[2], comm[0], comm[1])
        t2 = time.perf_counter()
        if fused is None:
            fused = True  # also with a communicator: its all-reduce is stream-ordered like the kernels around it
        res = drive(engine, config, solver=solver, niter=niter, block=block, gamma=gamma, adapt=adapt,
                    reweight_goal=reweight_goal, ignore=ignore, print_level=print_level, debug=bool(debug), fused=fused)
        # where the wall time of this call went: building the handle (incl. an NVRTC build on the first use of a problem),
        # creating the NCCL communicator, and the niter iterations themselves (fill + all-reduce + train!)
        res.timing = dict(create_s=t1 - t0, comm_init_s=t2 - t1, iterations_s=time.perf_counter() - t2)
        return res
    finally:
        engine.close()

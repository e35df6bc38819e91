#!/usr/bin/env python
"""bench.py — MC samples/sec (fp64) of the VEGAS fill path on BASELINE.json config C2:
10-D unit Gaussian, Continuous(-5,5), dof=[[10]], solver=:vegas, fp64 — and, in the same JSON line under
"bubble", the second workload the north star names: example/bubble.jl under :vegasmc (BASELINE C4).

One "step" = one iteration: 16 logical blocks per GPU x 2^26 samples = 2^30 samples per GPU
(fill kernel -> finalize -> [NCCL all-reduce of the packed statistics when N>1] -> device refine).
Weak scaling: every GPU keeps 2^30 samples/step, n_block = 16*N, blocks b mod N == rank.

  value          device-resident loop, timed with CUDA events on the library's stream, max over ranks
  e2e            the same step driven through the public host API with HOST buffers every step:
                 mci_set_maps (H2D from pinned memory) -> mci_run_iteration (D2H statistics) ->
                 host block statistics -> mci_refine -> mci_get_maps (D2H)
  roofline       fill kernel: algorithmic fp64 flop/sample x samples / CUDA-event time vs the DFMA
                 peak measured in this run (MEASURED_PEAKS.json has no fp64 figure); HBM is reported
                 beside it to show the path is not bandwidth-bound.  `traffic` / `limiters_ncu` come from the
                 committed ncu capture and are printed only while the kernel sources still hash to the
                 capture's `csrc_sha256` (tools/csrc_hash.py): a kernel change cannot leave them stale
  cpu_baseline   the CPU oracle (a port of the reference's Julia loop; Julia is not in the image), bounded
                 sample, on all host cores AND on one thread (the reference publishes ~4.1 M evals/s on
                 one thread for a lighter 3-D problem, example/benchmark/cuba/benchmark.jl:146-147)
  parity         before the timed region: a small fixed problem of each benched workload run (a) partitioned
                 over the ranks with the in-library NCCL all-reduce and (b) unpartitioned on rank 0; the sums
                 must agree to 1e-10 and the refined maps must be bit-identical on every rank, or the bench
                 exits non-zero.  At N = 1 the partition is emulated (two handles, blocks b mod 2).

`--impl reference` times that CPU port alone (rank 0 only) on the same config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

BLOCKS_PER_GPU = 16


class Workload:
    """what is benched: C2 (the metric BASELINE.json names) or the bubble (north star's second target)"""

    def __init__(self, name, npb):
        self.name, self.npb = name, npb
        if name == "gauss10":
            self.metric, self.unit = "MC samples/sec (fp64) for 10-D Gaussian", "samples/s"
            self.text = "C2: 10-D unit Gaussian, Continuous(-5,5), dof=[[10]], solver=vegas, 16 blocks/GPU x 2^26 samples"
            if npb != 1 << 26:
                self.text = f"C2: 10-D unit Gaussian, Continuous(-5,5), dof=[[10]], solver=vegas, 16 blocks/GPU x {npb} samples"
            # 8*10 (map) + 5 + 10 (hist) + 62 (integrand)
            self.flop, self.chain, self.ncu = 157.0, False, "r2_fill_ncu.json"
        else:
            self.metric, self.unit = "MC steps/sec (fp64) for example/bubble.jl under vegasmc", "steps/s"
            self.text = (f"C4: example/bubble.jl, (R,theta,phi,T) Continuous alpha=3 x Ext Discrete(1,4), 4-bin observable, "
                         f"solver=vegasmc, 16 blocks/GPU x {npb} steps")
            # algorithmic flop per VegasMC step (SURVEY 8d convention: FMA 2, add/mul 1, div/sqrt 8, exp/sincos 40 per
            # value): integrand ~400 (2 sincos, 4 exp, 6 div, ~60 mul/add) + shift 8 + padding products 10 + ratio 4 +
            # histogram weights of 4 adaptive leaves 22 + measurement 20
            self.flop, self.chain, self.ncu = 464.0, True, "r2_bubble_ncu.json"

    def spec(self, n_gpus, npb=None, device=0, seed=1234, cpu=False, n_block=None, walkers=None):
        import mcintegration_jl_b200 as mci
        from mcintegration_jl_b200 import _abi, integrands

        npb = self.npb if npb is None else npb
        nb = BLOCKS_PER_GPU * n_gpus if n_block is None else n_block
        if self.chain:
            from cases import case
            b = case("bubble")
            cfg = mci.Configuration(seed=seed, **b["cfg"]())
            # the CPU port runs the reference's decomposition (one chain per block); the GPU lets the library fill the SMs
            w = (1 if cpu else 0) if walkers is None else walkers
            spec = mci.problem_from_config(cfg, b["f"], solver=_abi.SOLVER_VEGASMC, measure_kind=_abi.MEASURE_BIN_BY_DISCRETE,
                                           n_block=nb, neval_per_block=npb, device=device, walkers_per_block=w,
                                           chain_burnin=0 if cpu else 1024)
            return spec, cfg
        cfg = mci.Configuration(var=mci.Continuous(-5.0, 5.0), dof=[[10]], seed=seed)
        spec = mci.problem_from_config(cfg, integrands.gauss_nd, solver=_abi.SOLVER_VEGAS, measure_kind=0, n_block=nb,
                                       neval_per_block=npb, device=device)
        return spec, cfg

    def exact(self):
        if self.chain:
            from mcintegration_jl_b200 import integrands as I
            return I.lindhard_finite_temperature(I.BubblePara(), 0.0)  # first momentum bin, exact finite-T value
        return math.erf(5 / math.sqrt(2)) ** 10


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(wl: Workload, target_s=10.0, threads=None):
    """Time the CPU oracle (port of the reference's montecarlo loop of this workload) for ~target_s seconds on
    `threads` host threads (default: all cores; blocks are the parallel unit, parallel=:thread, main.jl:153-158)."""
    from oracle_lib import OracleEngine

    cores = threads or (os.cpu_count() or 1)
    spec, _ = wl.spec(1, 1 << 12, cpu=True)
    spec.n_block = max(BLOCKS_PER_GPU if cores > 1 else 1, cores)
    eng = OracleEngine(spec, nthreads=cores)
    t0 = time.perf_counter(); eng.run_iteration(0); dt = time.perf_counter() - t0
    rate = spec.n_block * spec.neval_per_block / dt
    eng.close()
    npb = max(1 << 12, int(rate * target_s / spec.n_block))
    spec2, _ = wl.spec(1, npb, cpu=True)
    spec2.n_block = spec.n_block
    eng = OracleEngine(spec2, nthreads=cores)
    eng.run_iteration(0)
    if wl.chain:
        eng.reweight(1.0)
    eng.refine()  # adapt once so the timed pass runs on a trained map, like the GPU steps
    t0 = time.perf_counter(); eng.run_iteration(1); dt = time.perf_counter() - t0
    eng.close()
    n = spec2.n_block * npb
    return {"value": n / dt, "unit": wl.unit, "cores": cores, "kind": "port",
            "sample": f"{spec2.n_block} blocks x {npb} {'steps' if wl.chain else 'samples'} = {n} of the same workload in "
                      f"{dt:.2f} s (oracle/mci_oracle.c, {'OpenMP over blocks' if cores > 1 else 'one thread'})"}, n, dt


def cpu_baseline(wl: Workload, target_s):
    """all host cores + one thread (BASELINE.md 3 i/ii)"""
    cb, _, _ = cpu_reference_rate(wl, target_s=target_s)
    one, _, _ = cpu_reference_rate(wl, target_s=max(2.0, target_s / 3), threads=1)
    cb["single_thread"] = {"value": one["value"], "unit": one["unit"], "cores": 1, "sample": one["sample"],
                           "reference_published": "~4.1 M evals/s (vegas) / ~2.0 M evals/s (vegasmc) on one thread, 3-D, 11 integrands "
                                                  "(example/benchmark/cuba/benchmark.jl:119-120,146-147; hardware not stated)"}
    return cb


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, args.neval_per_block)
    vals, last = [], None
    budget = max(1.0, min(8.0, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        cb, n, dt = cpu_reference_rate(wl, target_s=budget)
        if i >= args.warmup:
            vals.append((n, dt)); last = cb
    n_tot = sum(v[0] for v in vals); t_tot = sum(v[1] for v in vals)
    v = n_tot / t_tot
    last["value"] = v
    one, _, _ = cpu_reference_rate(wl, target_s=3.0, threads=1)
    last["single_thread"] = {"value": one["value"], "unit": one["unit"], "cores": 1, "sample": one["sample"]}
    print(json.dumps({"impl": "reference", "metric": wl.metric, "value": v, "unit": wl.unit, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / len(vals),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic",
                      # the SAME workload as the GPU arm; each timed step is a bounded sample of it (see cpu_baseline.sample)
                      "config": {"workload": wl.text, "step_is": "a bounded CPU sample of the workload, see cpu_baseline.sample"},
                      "cpu_baseline": last,
                      "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------------------------------
def _raw(st, nb):
    """statistics without the additive floors of clearStatistics!/addConfig! ((B+1) x floor), so partitions can be summed"""
    B1 = nb + 1
    return [st.block_obs, st.block_norm - 1e-10, st.hist - B1 * 1e-10, st.visited - B1 * 1e-8, st.propose - B1 * 1e-8,
            st.accept - B1 * 1e-10]


def _rel(a, b):
    import numpy as np
    scale = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max() / scale) if a.size else 0.0


_JOB_NCCL_ID = None


def job_nccl_id():
    """The 128-byte NCCL unique id of this job: rank 0 creates it, torch.distributed carries the bytes.  The library owns
    the communicator and keeps it for the life of the process, so every engine of the job joins the same one."""
    global _JOB_NCCL_ID
    if _JOB_NCCL_ID is None:
        import torch
        import torch.distributed as dist

        import mcintegration_jl_b200 as mci

        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if dist.get_rank() == 0:
            idt.copy_(torch.frombuffer(bytearray(mci.CudaEngine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        _JOB_NCCL_ID = bytes(idt.cpu().numpy().tobytes())
    return _JOB_NCCL_ID


def parity_check(wl: Workload, rank, world, local, dist):
    """A small fixed problem of this workload, (a) partitioned over the ranks (in-library NCCL all-reduce; at N = 1 two
    handles with blocks b mod 2, summed on the host) and (b) unpartitioned: statistics of two iterations with a refine in
    between must agree to 1e-10, the refined maps must be bit-identical on every rank (main.jl:177-188 semantics, pinned
    in the reference by test/mpi_test.jl:73-146)."""
    import numpy as np
    import torch

    import mcintegration_jl_b200 as mci

    parts = max(world, 2)
    nb = 8 * parts
    npb = 64 * 256 if wl.chain else 1 << 14
    kw = dict(n_block=nb, npb=npb, walkers=64 if wl.chain else None)
    spec, _ = wl.spec(1, device=local, **kw)
    out = {"workload": wl.name, "n_block": nb, "neval_per_block": npb, "iterations": 2,
           "mode": "nccl" if world > 1 else "partition-emulated (2 handles, blocks b mod 2)"}

    def run(engines, reduce_host):
        stats = []
        for it in range(2):
            sts = [e.run_iteration(it) for e in engines]
            if reduce_host:
                raw = [sum(x) for x in zip(*[_raw(s, nb) for s in sts])]
                # the emulated ranks have no collective: give every handle the summed statistics by hand
                for e in engines:
                    ptr, n = e.stats_device_ptr()
                    packed = np.concatenate([raw[2], raw[0].ravel(), raw[1], raw[3], raw[4].ravel(), raw[5].ravel(),
                                             [float(sum(s.neval for s in sts))]])
                    assert packed.size == n
                    _h2d(ptr, packed)
            else:
                raw = _raw(sts[0], nb)
            stats.append(raw)
            for e in engines:
                if wl.chain:
                    e.reweight(1.0)
                e.refine()
        return stats, [e.get_maps() for e in engines]

    if world > 1:
        eng = mci.CudaEngine(spec)
        eng.comm_init(job_nccl_id(), rank, world)
        got, maps = run([eng], False)
        eng.close()
        # maps identical across ranks: compare every rank's grid bytes with rank 0's
        g = torch.from_numpy(np.concatenate([np.ravel(m) for m in maps[0]])).cuda()
        g0 = g.clone()
        dist.broadcast(g0, 0)
        same = torch.tensor([1 if torch.equal(g.view(torch.int64), g0.view(torch.int64)) else 0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["maps_identical_across_ranks"] = bool(same.item())
    else:
        engines = [mci.CudaEngine(spec) for _ in range(2)]
        for r, e in enumerate(engines):
            e.set_partition(r, 2)
        got, maps = run(engines, True)
        for e in engines:
            e.close()
        out["maps_identical_across_ranks"] = all(np.array_equal(a, b) for a, b in zip(maps[0], maps[1]))
    if rank == 0:
        one = mci.CudaEngine(spec)
        want, wmaps = run([one], False)
        one.close()
        out["max_rel_obs"] = max(max(_rel(g[0], w[0]), _rel(g[1], w[1])) for g, w in zip(got, want))
        out["max_rel_hist"] = max(_rel(g[2], w[2]) for g, w in zip(got, want))
        out["max_rel_counters"] = max(max(_rel(g[k], w[k]) for k in (3, 4, 5)) for g, w in zip(got, want))
        out["max_rel_maps_vs_unpartitioned"] = max(_rel(np.ravel(a), np.ravel(b)) for a, b in zip(maps[0], wmaps[0]) if a.size)
        out["ok"] = bool(out["maps_identical_across_ranks"] and out["max_rel_obs"] <= 1e-10 and out["max_rel_hist"] <= 1e-10
                         and out["max_rel_counters"] <= 1e-10)
    return out


def _h2d(dst, arr):
    import ctypes as C
    import numpy as np
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    assert rt.cudaMemcpy(C.c_void_p(dst), arr.ctypes.data_as(C.c_void_p), arr.nbytes, 1) == 0
    assert rt.cudaDeviceSynchronize() == 0


def ncu_capture(wl: Workload):
    """(traffic, limiters) of the committed ncu capture of this workload's fill kernel — only while the kernel sources
    are the ones that were profiled"""
    try:
        from csrc_hash import csrc_sha256
        rec = json.load(open(os.path.join(ROOT, "profiles", wl.ncu)))
        now = csrc_sha256()
        if rec.get("csrc_sha256") != now:
            return None, {"stale": f"profiles/{wl.ncu} was captured from csrc {str(rec.get('csrc_sha256'))[:12]}, the sources are now "
                                   f"{now[:12]}: not printed"}
        lim = {k: rec[k] for k in ("issue_active_pct", "smem_pipe_busy_frac", "warp_instructions_per_sample_warp", "fp64_pipe_pct",
                                   "warps_active_pct", "threads_per_inst", "source", "csrc_sha256", "git_commit") if k in rec}
        return rec.get("dram_bytes_per_launch"), lim
    except Exception as e:  # no capture committed for this workload
        return None, {"missing": f"profiles/{wl.ncu}: {type(e).__name__}"}


def bench_workload(wl: Workload, args, rank, world, local, dist, want_cpu):
    import numpy as np
    import torch

    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200.main import block_statistics

    spec, cfg = wl.spec(world, device=local)
    eng = mci.CudaEngine(spec)
    if world > 1:
        eng.comm_init(job_nccl_id(), rank, world)
    npb = wl.npb
    units_per_step = world * BLOCKS_PER_GPU * npb

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def refine():
        if wl.chain:  # chain solver: doReweight! belongs to the iteration (main.jl:183)
            eng.reweight(1.0)
        eng.refine()

    # ---------------- device-resident leg ("value") ----------------
    eng.set_async(True)
    it = 0
    for _ in range(args.warmup):
        eng.launch_iteration(it); refine(); it += 1
    eng.fetch_iteration()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    eng.timer_start()
    for _ in range(args.steps):
        eng.launch_iteration(it); refine(); it += 1
    ms = eng.timer_stop()
    st = eng.fetch_iteration()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = maxr(ms)
    value = units_per_step * args.steps / (ms * 1e-3)
    # per-launch fill-kernel time of THE TIMED STEPS: CUDA events recorded around every fill launch on the
    # library's own stream (mci_fill_history keeps the last 64 pairs)
    fills = eng.fill_history(min(64, args.steps))
    fill_ms = maxr(sum(fills) / len(fills))
    eng.set_async(False)
    mean, std = block_statistics(st, cfg, spec.n_block)

    # ---------------- end-to-end leg ("e2e"): host buffers every step ----------------
    grids, ddist, daccum, rw = eng.get_maps()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    grids_h, rw_h = pin(grids), pin(rw)
    h2d = grids_h.nbytes + rw_h.nbytes
    d2h = 8 * (eng.hist_len + spec.n_block * (eng.obs_total + 1) + eng.nd + 2 * eng.pa_len + 1) + grids_h.nbytes + rw_h.nbytes
    for _ in range(max(1, args.warmup // 2)):
        eng.set_maps(grids=grids_h, reweight=rw_h); s2 = eng.run_iteration(it); it += 1
        block_statistics(s2, cfg, spec.n_block); refine(); g2 = eng.get_maps(); grids_h[:] = g2[0]; rw_h[:] = g2[3]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_maps(grids=grids_h, reweight=rw_h)
        s2 = eng.run_iteration(it); it += 1
        m2, e2 = block_statistics(s2, cfg, spec.n_block)
        refine()
        g2 = eng.get_maps(); grids_h[:] = g2[0]; rw_h[:] = g2[3]
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e = units_per_step * args.steps / e2e_s

    out = None
    if rank == 0:
        peak = mci.fp64_peak_tflops(local)
        info = eng.launch_info()
        fill_rate = BLOCKS_PER_GPU * npb / (fill_ms * 1e-3)  # per-GPU units/s inside the fill kernel
        achieved = fill_rate * wl.flop / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes per launch (DESIGN.md "Roofline"): read the staged map image once per CTA (L2 hits after
        # the first), write one histogram partial per CTA and the per-tile / per-CTA partial statistics
        alg_bytes = 8.0 * (2 * eng.hist_len + eng.hist_len * info["grid"] + max(1, info["tiles_per_block"]) * BLOCKS_PER_GPU * 2)
        traffic, ncu = ncu_capture(wl)
        cb = cpu_baseline(wl, 8.0 if wl.name == "gauss10" else 5.0) if want_cpu else None
        chain = eng.chain_info() if wl.chain else None
        out = {"metric": wl.metric, "value": value, "unit": wl.unit, "ms_per_step": ms / args.steps,
               "config": {"workload": wl.text, ("steps_per_step" if wl.chain else "samples_per_step"): units_per_step,
                          "n_block": spec.n_block, "neval_per_block": npb,
                          "l2": "inputs are generated on chip (Philox counters); the working set (map tables + histogram per CTA) "
                                "lives in shared memory; every step draws fresh samples, so there is no L2-resident input to flush",
                          "kernel": info["kernel"], "grid": info["grid"], "threads": info["threads"],
                          "smem_bytes": info["smem_bytes"], "hist_mode": info["hist_mode"], "chains": chain},
               "result": {"integral": float(np.ravel(mean[0])[0]), "sigma": float(np.ravel(std[0])[0]), "exact": wl.exact()},
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if peak else None, "traffic": traffic,
                            "note": f"{wl.flop:.0f} algorithmic flop per {'step' if wl.chain else 'sample'} (SURVEY 8d) x "
                                    f"{BLOCKS_PER_GPU * npb} per launch / {fill_ms:.3f} ms fill kernel (CUDA events, library "
                                    "stream); peak = DFMA probe measured in this run (mci_fp64_peak; MEASURED_PEAKS.json has no "
                                    "fp64 figure; theoretical 64 DFMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s)",
                            "fill_ms": fill_ms, "fill_units_per_s": fill_rate,
                            "limiters_ncu": ncu,  # what actually binds (committed ncu capture of the same kernel sources)
                            "hbm": {"achieved_gbs": alg_bytes / (fill_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                    "frac": alg_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured" if peaks else "fallback",
                                    "algorithmic_bytes_per_launch": alg_bytes}},
               "cpu_baseline": cb,
               "e2e": {"value": e2e, "unit": wl.unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(launches), "clocks": clocks}
    eng.close()
    return out


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever native libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    main_wl = Workload(args.workload, args.neval_per_block)
    other = None
    if args.workload == "gauss10" and not args.no_bubble:
        other = Workload("bubble", 1 << 25)  # ~3600 measured steps per chain on top of the 1024 burn-in steps
    # ---- correctness first: partitioned (NCCL) against unpartitioned, both workloads ----
    parity = [parity_check(w, rank, world, local, dist) for w in ([main_wl] + ([other] if other else []))]
    bad = rank == 0 and not all(p.get("ok") for p in parity)

    out = bench_workload(main_wl, args, rank, world, local, dist, want_cpu=(world == 1 and not args.no_cpu))
    sub = None
    if other is not None:
        sub_args = argparse.Namespace(**vars(args))
        sub_args.steps, sub_args.warmup = min(args.steps, 10), min(args.warmup, 3)
        sub = bench_workload(other, sub_args, rank, world, local, dist, want_cpu=(world == 1 and not args.no_cpu))
    if dist is not None:
        dist.barrier()
        import mcintegration_jl_b200 as mci

        mci.CudaEngine.comm_finalize()  # the job's NCCL communicator (every engine joined it: job_nccl_id)
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    if rank == 0:
        line = {"metric": out["metric"], "value": out["value"], "unit": out["unit"], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": out["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update({k: out[k] for k in ("config", "result", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks")})
        line["parity"] = parity
        if sub is not None:
            sub.update({"steps": min(args.steps, 10), "warmup": min(args.warmup, 3), "n_gpus": world, "higher_is_better": True,
                        "scaling": "weak", "dtype": "f64"})
            line["bubble"] = sub
            line["gpu_launches"] = int(out["gpu_launches"]) + int(sub["gpu_launches"])
        print(json.dumps(line), flush=True)
    if bad:
        sys.stderr.write("parity check FAILED: " + json.dumps(parity) + "\n")
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--neval-per-block", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-bubble", action="store_true", help="skip the second workload (the \"bubble\" object of the line)")
    ap.add_argument("--workload", default="gauss10", choices=["gauss10", "bubble"],
                    help="gauss10 = BASELINE C2 (the metric; the bubble rides along as a sub-object); bubble = example/bubble.jl "
                         "under vegasmc (C4) alone, steps/s")
    args = ap.parse_args()
    if args.neval_per_block is None:
        args.neval_per_block = 1 << 26 if args.workload == "gauss10" else 1 << 25
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

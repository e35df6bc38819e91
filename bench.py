#!/usr/bin/env python
"""bench.py — MC samples/sec (fp64) of the VEGAS fill path on BASELINE.json config C2:
10-D unit Gaussian, Continuous(-5,5), dof=[[10]], solver=:vegas, fp64.

One "step" = one VEGAS iteration: 16 logical blocks per GPU x 2^26 samples = 2^30 samples per GPU
(fill kernel -> finalize -> [NCCL all-reduce of the packed statistics when N>1] -> device refine).
Weak scaling: every GPU keeps 2^30 samples/step, n_block = 16*N, blocks b mod N == rank.

  value          device-resident loop, timed with CUDA events on the library's stream, max over ranks
  e2e            the same step driven through the public host API with HOST buffers every step:
                 mci_set_maps (H2D from pinned memory) -> mci_run_iteration (D2H statistics) ->
                 host block statistics -> mci_refine -> mci_get_maps (D2H)
  roofline       fill kernel: algorithmic fp64 flop/sample x samples / CUDA-event time vs the DFMA
                 peak measured in this run (MEASURED_PEAKS.json has no fp64 figure); HBM is reported
                 beside it to show the path is not bandwidth-bound
  cpu_baseline   the CPU oracle (a port of the reference's Julia loop; Julia is not in the image) on
                 all host cores, bounded sample

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

METRIC = "MC samples/sec (fp64) for 10-D Gaussian"
UNIT = "samples/s"
WORKLOAD = "C2: 10-D unit Gaussian, Continuous(-5,5), dof=[[10]], solver=vegas, 16 blocks/GPU x 2^26 samples"
FLOP_PER_SAMPLE = 157.0  # 8*10 (map) + 5 + 10 (hist) + 62 (integrand)
BLOCKS_PER_GPU = 16
BUBBLE = False  # --workload bubble: the second workload the north star names (example/bubble.jl, solver=vegasmc)
# algorithmic flop per VegasMC step on bubble.jl (SURVEY 8d convention: FMA 2, add/mul 1, div/sqrt 8, exp/sincos 40 per
# value): integrand ~400 (2 sincos, 4 exp, 6 div, ~60 mul/add) + shift 8 + padding products 10 + ratio 4 + histogram
# weights of 4 adaptive leaves 22 + measurement 20
BUBBLE_FLOP_PER_STEP = 464.0


def _spec(n_gpus, neval_per_block, device=0, seed=1234, cpu=False):
    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200 import _abi, integrands

    if BUBBLE:
        from cases import case
        b = case("bubble")
        cfg = mci.Configuration(seed=seed, **b["cfg"]())
        # the CPU port runs the reference's decomposition (one chain per block); the GPU lets the library fill the SMs
        spec = mci.problem_from_config(cfg, b["f"], solver=_abi.SOLVER_VEGASMC, measure_kind=_abi.MEASURE_BIN_BY_DISCRETE,
                                       n_block=BLOCKS_PER_GPU * n_gpus, neval_per_block=neval_per_block, device=device,
                                       walkers_per_block=1 if cpu else 0, chain_burnin=0 if cpu else 1024)
        return spec, cfg
    cfg = mci.Configuration(var=mci.Continuous(-5.0, 5.0), dof=[[10]], seed=seed)
    spec = mci.problem_from_config(cfg, integrands.gauss_nd, solver=_abi.SOLVER_VEGAS, measure_kind=0,
                                   n_block=BLOCKS_PER_GPU * n_gpus, neval_per_block=neval_per_block, device=device)
    return spec, cfg


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


def cpu_reference_rate(n_blocks=BLOCKS_PER_GPU, target_s=12.0, iters_after=1):
    """Time the CPU oracle (port of vegas/montecarlo.jl:117-187) on all host cores for ~target_s seconds."""
    from oracle_lib import OracleEngine

    cores = os.cpu_count() or 1
    spec, _ = _spec(1, 1 << 14, cpu=True)
    spec.n_block = max(n_blocks, cores)  # blocks are the parallel unit (parallel=:thread, main.jl:153-158)
    eng = OracleEngine(spec, nthreads=cores)
    t0 = time.perf_counter(); eng.run_iteration(0); dt = time.perf_counter() - t0
    rate = spec.n_block * spec.neval_per_block / dt
    eng.refine()
    eng.close()
    npb = max(1 << 14, int(rate * target_s / spec.n_block))
    spec2, _ = _spec(1, npb, cpu=True)
    spec2.n_block = spec.n_block
    eng = OracleEngine(spec2, nthreads=cores)
    eng.run_iteration(0); eng.refine()  # adapt once so the timed pass runs on a trained map, like the GPU steps
    t0 = time.perf_counter(); eng.run_iteration(1); dt = time.perf_counter() - t0
    eng.close()
    n = spec2.n_block * npb
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{spec2.n_block} blocks x {npb} samples = {n} samples of the same workload in {dt:.2f} s "
                      f"(oracle/mci_oracle.c, OpenMP over blocks)"}, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        cb, n, dt = cpu_reference_rate(target_s=max(1.0, min(8.0, 60.0 / max(1, args.steps + args.warmup))))
        if i >= args.warmup:
            vals.append((n, dt)); last = cb
    n_tot = sum(v[0] for v in vals); t_tot = sum(v[1] for v in vals)
    v = n_tot / t_tot
    last["value"] = v
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / len(vals),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": {"workload": WORKLOAD + " (bounded CPU sample per step)"},
                      "cpu_baseline": last,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_ours(args):
    import numpy as np
    import torch

    import mcintegration_jl_b200 as mci
    from mcintegration_jl_b200.main import block_statistics

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

    npb = args.neval_per_block
    spec, cfg = _spec(world, npb, device=local)
    eng = mci.CudaEngine(spec)
    if world > 1:
        # the library owns its communicator: rank 0 creates the id, torch.distributed carries the 128 bytes
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mci.CudaEngine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    samples_per_step = world * BLOCKS_PER_GPU * npb

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

    if BUBBLE:  # chain solver: doReweight! belongs to the iteration (main.jl:183)
        _refine = eng.refine
        eng.refine = lambda: (eng.reweight(1.0), _refine())
    # ---------------- device-resident leg ("value") ----------------
    eng.set_async(True)
    it = 0
    for _ in range(args.warmup):
        eng.launch_iteration(it); eng.refine(); it += 1
    eng.fetch_iteration()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    fills = []
    eng.timer_start()
    for _ in range(args.steps):
        eng.launch_iteration(it); eng.refine(); it += 1
    ms = eng.timer_stop()
    st = eng.fetch_iteration()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = maxr(ms)
    value = samples_per_step * args.steps / (ms * 1e-3)
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
        block_statistics(s2, cfg, spec.n_block); eng.refine(); g2 = eng.get_maps(); grids_h[:] = g2[0]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_maps(grids=grids_h, reweight=rw_h)
        s2 = eng.run_iteration(it); it += 1
        m2, e2 = block_statistics(s2, cfg, spec.n_block)
        eng.refine()
        g2 = eng.get_maps(); grids_h[:] = g2[0]
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    e2e = samples_per_step * args.steps / e2e_s

    out = None
    if rank == 0:
        peak = mci.fp64_peak_tflops(local)
        info = eng.launch_info()
        fill_rate = BLOCKS_PER_GPU * npb / (fill_ms * 1e-3)  # per-GPU samples/s inside the fill kernel
        achieved = fill_rate * FLOP_PER_SAMPLE / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes per launch (DESIGN.md "Roofline"): read the staged map image once per CTA (L2 hits after
        # the first), write one histogram partial per CTA and (obs, count) per tile
        alg_bytes = 8.0 * (2 * eng.hist_len + eng.hist_len * info["grid"] + info["tiles_per_block"] * BLOCKS_PER_GPU * 2)
        traffic, ncu = None, None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel (profiles/)
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r1_fill_ncu.json")))
            traffic = ncu["dram_bytes_per_launch"]
            ncu = {k: ncu[k] for k in ("issue_active_pct", "smem_pipe_busy_frac", "warp_instructions_per_sample_warp", "source") if k in ncu}
        except Exception:
            pass
        if BUBBLE:
            from mcintegration_jl_b200 import integrands as _I
            exact_value = _I.lindhard_finite_temperature(_I.BubblePara(), 0.0)  # first momentum bin, exact finite-T value
            traffic, ncu = None, None  # the ncu capture in profiles/r1_fill_ncu.json is of the C2 kernel
        else:
            exact_value = math.erf(5 / math.sqrt(2)) ** 10
        cb = None
        if world == 1 and not args.no_cpu:
            cb, _, _ = cpu_reference_rate()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": WORKLOAD, "samples_per_step": samples_per_step, "n_block": spec.n_block,
                          "neval_per_block": npb, "l2": "inputs are generated on chip (Philox counters); the working set (16 KB map + 128 KB "
                          "striped histogram per CTA) lives in shared memory; every step draws fresh samples, so there is "
                          "no L2-resident input to flush",
                          "kernel": info["kernel"], "grid": info["grid"], "threads": info["threads"],
                          "smem_bytes": info["smem_bytes"], "hist_mode": info["hist_mode"]},
               "result": {"integral": float(np.ravel(mean[0])[0]), "sigma": float(np.ravel(std[0])[0]),
                          "exact": exact_value},
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if peak else None, "traffic": traffic,
                            "note": f"{FLOP_PER_SAMPLE:.0f} algorithmic flop/sample (SURVEY 8d) x {BLOCKS_PER_GPU * npb} "
                                    f"samples per launch / {fill_ms:.3f} ms fill kernel (CUDA events, library stream); peak = "
                                    "DFMA probe measured in this run (mci_fp64_peak; MEASURED_PEAKS.json has no fp64 figure)",
                            "fill_ms": fill_ms, "fill_samples_per_s": fill_rate,
                            "limiters_ncu": ncu,  # what actually binds (offline ncu capture of the same launch): issue slots + shared-memory wavefronts
                            "hbm": {"achieved_gbs": alg_bytes / (fill_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                    "frac": alg_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak, "of": "measured" if peaks else "fallback",
                                    "algorithmic_bytes_per_launch": alg_bytes}},
               "cpu_baseline": cb,
               "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
               "gpu_launches": int(launches), "clocks": clocks}
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--neval-per-block", type=int, default=1 << 26)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="gauss10", choices=["gauss10", "bubble"],
                    help="gauss10 = BASELINE C2 (the metric); bubble = example/bubble.jl under vegasmc (C4), steps/s")
    args = ap.parse_args()
    if args.workload == "bubble":
        global BUBBLE, WORKLOAD, FLOP_PER_SAMPLE, METRIC, UNIT
        BUBBLE = True
        if args.neval_per_block == 1 << 26:
            args.neval_per_block = 1 << 25  # ~3600 measured steps per chain on top of the 1024 burn-in steps
        WORKLOAD = (f"C4: example/bubble.jl, (R,theta,phi,T) Continuous alpha=3 x Ext Discrete(1,4), 4-bin observable, solver=vegasmc, "
                    f"16 blocks/GPU x {args.neval_per_block} steps")
        FLOP_PER_SAMPLE = BUBBLE_FLOP_PER_STEP
        METRIC = "MC steps/sec (fp64) for example/bubble.jl under vegasmc"
        UNIT = "steps/s"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

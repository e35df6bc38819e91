#!/usr/bin/env python
"""BASELINE C4 / C5 on several GPUs through the public API: one process per GPU (torchrun), blocks partitioned over the
ranks, ONE NCCL all-reduce of the packed statistics (histograms + block sums) per iteration inside libmcib200.
  python -m torch.distributed.run --nproc-per-node 8 tools/run_configs_multi.py c5
  python -m torch.distributed.run --nproc-per-node 4 tools/run_configs_multi.py c4"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200 import integrands as I
from cases import case

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()


def fresh_comm():
    """(rank, world, NCCL unique id): the library keeps one communicator per id for the life of the process"""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(mci.CudaEngine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return (rank, world, bytes(idt.cpu().numpy().tobytes()))


block = 16 * world
def problem():
    """fresh variables for every call: like the reference's, a Configuration resizes the pool of the variable it is
    given (configuration.jl:156-160) and refuses to do that twice"""
    if which == "c5":
        kw = dict(solver="mcmc", neval=1e9, niter=10, block=block, var=mci.Continuous(0.0, 1.0), dof=[[20]] * 4)
        return kw, I.product4(), [[1.0]] * 4, "C5 20-D 4-component product, mcmc, neval=1e9/iter"
    b = case("bubble")
    para = b["para"]
    kw = dict(solver="vegasmc", neval=1e8, niter=10, block=block, measure="bin_by_discrete", **b["cfg"]())
    return kw, b["f"], [[I.lindhard_finite_temperature(para, q) for q in para.extQ]], "C4 example/bubble.jl, vegasmc, neval=1e8/iter"


# the first call of a process pays for the CUDA context, the NVRTC build of the problem's kernel and NCCL's topology
# discovery: run one iteration of the same problem first, then time the real call (which joins the same communicator)
kw, f, exact, tag = problem()
comm = fresh_comm()
mci.integrate(f, seed=1, device=local, comm=comm, **dict(kw, niter=1))  # same size: the launch plan, hence the NVRTC program, depends on it
kw, f, exact, tag = problem()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
res = mci.integrate(f, seed=1234, device=local, comm=comm, **kw)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    mean = [np.ravel(m).tolist() for m in res.mean]
    std = [np.ravel(s).tolist() for s in res.stdev]
    dev = [[(m - e) / s for m, s, e in zip(mm, ss, ee)] for mm, ss, ee in zip(mean, std, exact)]
    print(json.dumps(dict(config=tag, n_gpus=world, block=block, mean=mean, sigma=std, dev_sigma=dev, wall_s=dt,
                          timing=res.timing, steps_per_s=kw["neval"] * kw["niter"] / res.timing["iterations_s"],
                          note="steps_per_s over the iterations only; wall_s also holds the creation of the handle")),
          flush=True)
dist.barrier()
mci.CudaEngine.comm_finalize()
dist.destroy_process_group()

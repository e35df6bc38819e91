"""Worker of tests/test_gpu_multi.py: one process per GPU; libmcib200's own NCCL all-reduce of the packed statistics
(mci_comm_init) with the 128-byte unique id carried by torch.distributed, then the ordinary host driver."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import mcintegration_jl_b200 as mci  # noqa: E402
from mcintegration_jl_b200.main import drive  # noqa: E402
from cases import make_spec  # noqa: E402


def main():
    name, solver, out = sys.argv[1], sys.argv[2], sys.argv[3]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    kw = dict(n_block=8, neval_per_block=20000) if solver == "vegas" else dict(n_block=8, neval_per_block=64 * 300, walkers=64)
    spec, cfg = make_spec(name, solver, **kw)
    spec.device = local
    eng = mci.CudaEngine(spec)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(mci.CudaEngine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    uid = bytes(idt.cpu().numpy().tobytes())
    eng.comm_init(uid, rank, world)
    res = drive(eng, cfg, solver=solver, niter=4, block=spec.n_block)
    maps = eng.get_maps()
    eng.close()
    # a second handle joins the communicator of the same id and runs the same job as ONE fused library call
    # (mci_run_iterations: the all-reduce is stream-ordered like everything else, no host round trip per iteration)
    spec2, cfg2 = make_spec(name, solver, **kw)
    spec2.device = local
    eng2 = mci.CudaEngine(spec2)
    eng2.comm_init(uid, rank, world)
    res2 = drive(eng2, cfg2, solver=solver, niter=4, block=spec2.n_block, fused=True)
    maps2 = eng2.get_maps()
    eng2.close()
    with open(f"{out}.{rank}", "wb") as f:
        pickle.dump(dict(mean=[np.asarray(m) for m in res.mean], std=[np.asarray(s) for s in res.stdev], maps=maps,
                         neval=res.neval, fused_mean=[np.asarray(m) for m in res2.mean],
                         fused_std=[np.asarray(s) for s in res2.stdev], fused_maps=maps2, fused_neval=res2.neval), f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

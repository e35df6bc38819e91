"""Worker of tests/test_dist_gloo.py: one process per rank (gloo, CPU).  Each rank runs its share of the blocks
(b mod world == rank, main.jl:122,152-166) on the ORACLE engine, the packed statistics are summed with ONE
all_reduce — the collective libmcib200 issues through NCCL on GPUs (replacing MPIreduce + MPIreduceConfig! +
MPIbcastConfig!, main.jl:177-188) — and every rank then refines / reweights identically through the product's
own host driver (mcintegration_jl_b200.main.drive)."""
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
from cases import case, make_spec  # noqa: E402
from oracle_lib import OracleEngine  # noqa: E402


def main():
    name, solver, out = sys.argv[1], sys.argv[2], sys.argv[3]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    kw = dict(n_block=8, neval_per_block=2000) if solver == "vegas" else dict(n_block=8, neval_per_block=16 * 200, walkers=16)
    spec, cfg = make_spec(name, solver, **kw)
    eng = OracleEngine(spec, nthreads=1)
    eng.set_partition(rank, world)

    def run_iteration(it):
        eng.run_local(it)
        t = torch.from_numpy(eng.packed())  # shares memory with the engine's packed buffer
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return eng.fetch_iteration()

    res = drive(eng, cfg, solver=solver, niter=4, block=spec.n_block, run_iteration=run_iteration)
    maps = eng.get_maps()
    with open(f"{out}.{rank}", "wb") as f:
        pickle.dump(dict(mean=[np.asarray(m) for m in res.mean], std=[np.asarray(s) for s in res.stdev], maps=maps,
                         last_hist=res.iterations[-1][2]["hist"], neval=res.neval), f)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

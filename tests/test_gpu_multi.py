"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): blocks partitioned b mod world == rank, ONE NCCL
all-reduce of the packed statistics per iteration inside libmcib200 (replaces MPIreduce/MPIreduceConfig!/
MPIbcastConfig!, src/main.jl:177-188), identical device refine on every rank — against the single-GPU run."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200.main import drive
from cases import make_spec

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name,solver", [("gauss10", "vegas"), ("bubble", "vegasmc"), ("sphere2", "mcmc")])
def test_nccl_ranks_match_single_gpu(tmp_path, name, solver):
    n = min(_ngpu(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    out = str(tmp_path / "rank")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200), os.path.join(HERE, "nccl_worker.py"), name, solver, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ranks = [pickle.load(open(f"{out}.{k}", "rb")) for k in range(world)]
    for k in range(1, world):  # identical statistics on every rank -> bit-identical maps and results
        for a, b in zip(ranks[0]["maps"], ranks[k]["maps"]):
            assert np.array_equal(a, b)
        for a, b in zip(ranks[0]["mean"], ranks[k]["mean"]):
            assert np.array_equal(a, b)
    # the fused multi-iteration call over the shared communicator = the iteration-by-iteration driver
    for k in range(world):
        for a, b in zip(ranks[k]["maps"], ranks[k]["fused_maps"]):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-12)
        for a, b in zip(ranks[k]["mean"], ranks[k]["fused_mean"]):
            np.testing.assert_allclose(a, b, rtol=1e-8)
        for a, b in zip(ranks[k]["std"], ranks[k]["fused_std"]):
            np.testing.assert_allclose(a, b, rtol=1e-5)
        assert ranks[k]["neval"] == ranks[k]["fused_neval"]
        for a, b in zip(ranks[0]["fused_maps"], ranks[k]["fused_maps"]):
            assert np.array_equal(a, b)
    kw = dict(n_block=8, neval_per_block=20000) if solver == "vegas" else dict(n_block=8, neval_per_block=64 * 300, walkers=64)
    spec, cfg = make_spec(name, solver, **kw)
    eng = mci.CudaEngine(spec)
    res = drive(eng, cfg, solver=solver, niter=4, block=spec.n_block)
    maps = eng.get_maps()
    eng.close()
    for a, b in zip(ranks[0]["maps"], maps):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-12)
    for k in range(len(res.mean)):
        np.testing.assert_allclose(ranks[0]["mean"][k], res.mean[k], rtol=1e-8)
        np.testing.assert_allclose(ranks[0]["std"][k], res.stdev[k], rtol=1e-5)
    assert ranks[0]["neval"] == res.neval

"""CPU tests of the N>1 path with world_size 2 over gloo (tests/dist_worker.py): block partition + one all-reduce
of the packed statistics + identical refinement on every rank must reproduce the single-rank run."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import mcintegration_jl_b200 as mci
from mcintegration_jl_b200.main import drive
from cases import make_spec
from oracle_lib import OracleEngine

HERE = os.path.dirname(os.path.abspath(__file__))


def _single(name, solver):
    kw = dict(n_block=8, neval_per_block=2000) if solver == "vegas" else dict(n_block=8, neval_per_block=16 * 200, walkers=16)
    spec, cfg = make_spec(name, solver, **kw)
    eng = OracleEngine(spec, nthreads=1)
    res = drive(eng, cfg, solver=solver, niter=4, block=spec.n_block)
    maps = eng.get_maps()
    eng.close()
    return res, maps


@pytest.mark.parametrize("name,solver", [("gauss4", "vegas"), ("bubble", "vegasmc"), ("sphere2", "mcmc")])
def test_two_ranks_match_one(tmp_path, name, solver):
    out = str(tmp_path / "rank")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py"), name, solver, out]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ranks = [pickle.load(open(f"{out}.{k}", "rb")) for k in range(2)]
    # every rank holds identical statistics after the all-reduce, so it refines to bit-identical maps
    for a, b in zip(ranks[0]["maps"], ranks[1]["maps"]):
        assert np.array_equal(a, b)
    for a, b in zip(ranks[0]["mean"], ranks[1]["mean"]):
        assert np.array_equal(a, b)
    # and the union of the two partitions is the single-rank run up to summation order
    res, maps = _single(name, solver)
    for a, b in zip(ranks[0]["maps"], maps):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-13)
    for k in range(len(res.mean)):
        np.testing.assert_allclose(ranks[0]["mean"][k], res.mean[k], rtol=1e-9)
        np.testing.assert_allclose(ranks[0]["std"][k], res.stdev[k], rtol=1e-6)
    assert ranks[0]["neval"] == res.neval

"""GPU (-m gpu): the NCCL communicator behind the C ABI (nlp_comm_init): nlp_predict merges across
the ranks itself -- global cutoff from all-reduced select histograms, one all-gather, final sort --
and every rank ends with the single-GPU result, bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run_world(world, tmp_path):
    import comm_worker as W
    uid = str(tmp_path / ("uid%d" % world))
    outs = [str(tmp_path / ("out%d_%d.npz" % (world, r))) for r in range(world)]
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "comm_worker.py"), str(r), str(world), uid, outs[r]], env=env)
             for r in range(world)]
    for pr in procs:
        assert pr.wait(timeout=600) == 0
    return W, [np.load(o) for o in outs]


def _check(W, results, oracle):
    off, keys = W.graph()
    for path in (0, 3, 1):
        for i, (m, D, K) in enumerate(W.CASES):
            if D == 0 and path != 0:
                continue
            want = oracle.oracle_predict(off, keys, m, D, max_edges=K)[:3]
            for r, z in enumerate(results):
                got = (z["u_%d_%d" % (path, i)], z["v_%d_%d" % (path, i)], z["s_%d_%d" % (path, i)])
                err = parity.compare(got, want, "world %d rank %d path %d %s D=%d K=%d" % (len(results), r, path, m, D, K))
                assert err is None, err


def test_comm_single_rank(oracle, tmp_path):
    """One rank: the same NCCL code path (all-reduce / all-gather over a communicator of one)."""
    W, results = _run_world(1, tmp_path)
    _check(W, results, oracle)
    assert int(results[0]["bytes"][0]) > 0


def test_comm_two_gpus(oracle, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    W, results = _run_world(2, tmp_path)
    _check(W, results, oracle)


def test_comm_four_gpus(oracle, tmp_path):
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs four GPUs")
    W, results = _run_world(4, tmp_path)
    _check(W, results, oracle)

"""CPU, world_size 2 over gloo: the multi-GPU plumbing (variable-length candidate all-gather and
the global merge) with per-rank candidate lists cut from the oracle the way the GPU ranks cut
theirs (source vertices owned in blocks of 32 ids, round-robin)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, K, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nlp_b200 as N
    from oracle import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, keys = N.graphs.to_numpy(*N.graphs.rmat(10, 8, 1))
    u, v, s, _ = O.oracle_predict(off, keys, "AA", 0)                 # every candidate, canonical
    mine = N.distributed.owner_of_vertex(u.astype(np.int64), world) == rank
    lu, lv, ls = u[mine][:K], v[mine][:K], s[mine][:K]                   # local top-K of this rank
    gu, gv, gs = N.distributed.gather_candidates(torch.from_numpy(lu.view(np.int32).copy()),
                                                 torch.from_numpy(lv.view(np.int32).copy()),
                                                 torch.from_numpy(ls.copy()))
    gu, gv, gs = gu.numpy().view(np.uint32), gv.numpy().view(np.uint32), gs.numpy()
    o = O.canonical_order(gu, gv, gs)[:K]
    ok = (np.array_equal(gu[o], u[:K]) and np.array_equal(gv[o], v[:K]) and
          np.array_equal(gs[o].view(np.uint32), s[:K].view(np.uint32)))
    q.put((rank, bool(ok), int(mine.sum())))
    dist.destroy_process_group()


@pytest.mark.parametrize("K", [50, 5000, 10 ** 7])
def test_gather_and_merge_world2(oracle, K):
    world = 2
    port = 29500 + (os.getpid() + K) % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in out), out
    assert all(n > 0 for _, _, n in out)

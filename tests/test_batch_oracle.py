"""CPU: the random edge removal that precedes every prediction (inc/batch.hxx:99-112, 200-208;
SURVEY.md section 8f-3, not on the GPU yet).  The plain-C restatement (oracle/batch_oracle.c) must
reproduce the compiled reference draw for draw -- same std::default_random_engine seed, same
removed edges -- and the slot-parallel formulation (tests/batch_parallel.py, the algorithm for a
GPU kernel) must reproduce both."""
import os

import numpy as np
import pytest

import batch_parallel
import parity

HAVE_REF = os.path.isdir("/root/reference") or os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libnlpref_batch.so"))


def graphs(nlp):
    g = nlp.graphs
    return {
        "rmat12": g.to_numpy(*g.rmat(12, 16, 31)),             # many isolated vertices: retries
        "road60": g.to_numpy(*g.road_lattice(60, 0.6, 33)),
        "web20k": g.to_numpy(*g.web_crawl(20000, 10, window=500, seed=36)),
        "empty": (np.zeros(6, np.uint64), np.empty(0, np.uint32)),      # every draw fails five times
    }


SEEDS = (0, 1, 12345, 2147483647, 4000000000)


@pytest.mark.skipif(not HAVE_REF, reason="compiled reference (oracle/_ref/libnlpref_batch.so) not available")
def test_oracle_matches_reference_batch_generator(nlp, oracle):
    if not oracle.ref_batch_available():
        pytest.skip("oracle/_ref/libnlpref_batch.so missing")
    for name, (off, keys) in graphs(nlp).items():
        for seed in SEEDS:
            for B in (0, 1, 17, max(2, len(keys) // 20)):
                u, v, words = oracle.oracle_edge_deletions(off, keys, seed, B)
                ru, rv = oracle.ref_edge_deletions(off, keys, seed, B)
                assert len(u) == len(ru) and (u == ru).all() and (v == rv).all(), (name, seed, B)


def test_slot_parallel_formulation_matches_oracle(nlp, oracle):
    for name, (off, keys) in graphs(nlp).items():
        for seed in SEEDS:
            for B in (1, 2, 17, 1000, max(2, len(keys) // 20)):
                u, v, words = oracle.oracle_edge_deletions(off, keys, seed, B)
                pu, pv, pwords = batch_parallel.edge_deletions_parallel(off, keys, seed, B)
                assert len(u) == len(pu) and (u == pu).all() and (v == pv).all(), (name, seed, B)
                assert words == pwords, (name, seed, B, words, pwords)


def test_slot_parallel_formulation_large_batch(nlp, oracle):
    """More deletions than the graph has edges (18 pointer-doubling rounds, heavy duplication)."""
    off, keys = graphs(nlp)["web20k"]
    B = 200_000
    u, v, words = oracle.oracle_edge_deletions(off, keys, 99, B)
    pu, pv, pwords = batch_parallel.edge_deletions_parallel(off, keys, 99, B)
    assert len(u) == len(pu) and (u == pu).all() and (v == pv).all() and words == pwords
    assert len(u) < 2 * B                                  # duplicates were collapsed


def test_removed_edges_exist_and_are_symmetric(nlp, oracle):
    off, keys = graphs(nlp)["rmat12"]
    u, v, _ = oracle.oracle_edge_deletions(off, keys, 7, len(keys) // 10)
    S = len(off) - 1
    src = np.repeat(np.arange(S, dtype=np.int64), np.diff(off.astype(np.int64)))
    have = set((src * S + keys.astype(np.int64)).tolist())
    comp = u.astype(np.int64) * S + v.astype(np.int64)
    assert all(c in have for c in comp.tolist())
    assert set(comp.tolist()) == set((v.astype(np.int64) * S + u.astype(np.int64)).tolist())
    assert (np.diff(comp) > 0).all()          # sorted, unique


@pytest.mark.gpu
def test_device_batch_generation_matches_oracle(nlp, oracle):
    """nlp_generate_deletions (csrc/batch.cuh) against the sequential oracle: the same removed edges
    and the same stream position for the same std::default_random_engine seed."""
    pred = nlp.Predictor(0)
    try:
        for name, (off, keys) in graphs(nlp).items():
            if name == "web20k":
                continue
            pred.set_graph(off, keys)
            for seed in (0, 12345, 4000000000):
                for B in (0, 1, 2, 17, max(3, len(keys) // 20)):
                    u, v, words = oracle.oracle_edge_deletions(off, keys, seed, B)
                    gu, gv, gwords = pred.generate_deletions(seed, B)
                    assert len(gu) == len(u) and (gu == u).all() and (gv == v).all(), (name, seed, B, len(gu), len(u))
                    assert gwords == words, (name, seed, B, gwords, words)
        # the generated list is main.cxx's sorted deletions0: it can be handed to nlp_set_truth as it lies
        off, keys = graphs(nlp)["rmat12"]
        pred.set_graph(off, keys)
        n, _ = pred.generate_deletions(7, 500, fetch=False)
        du, dv, dn = pred.deletions_device()
        assert dn == n
        pred.set_truth_pointers(du, dv, dn)
        r = pred.predict("JC", 0, max_edges=n // 2)
        ev = pred.evaluate()
        assert ev["truth"] == n and ev["predicted"] == 2 * r["count"] and ev["common"] == 0   # existing edges are never predicted
    finally:
        pred.close()


@pytest.mark.gpu
def test_device_apply_deletions_matches_reference(nlp, oracle):
    """nlp_apply_deletions (mark by binary search, rescan offsets, compact keys) against the
    reference's own applyBatchUpdateOmpU (oracle/_ref, simple graphs) and against the torch
    restatement (also multiset rows: one stored copy goes per request); the rebuilt graph then
    predicts like a freshly uploaded one.  Graph lent with nlp_set_graph_device stays untouched."""
    import torch
    g = nlp.graphs
    pred = nlp.Predictor(0)
    try:
        cases = dict(graphs(nlp))
        o, k = g.duplicate_symmetric(*g.planted_partition(3000, 40, 8, 2, 77), every=3, copies=2)
        cases["pp3k_symdup"] = g.to_numpy(o, k)
        for name, (off, keys) in cases.items():
            B = max(3, len(keys) // 20)
            for mode in ("host", "device"):
                if mode == "host":
                    pred.set_graph(off, keys)
                else:
                    d_off = torch.from_numpy(off.astype(np.int64)).cuda(); d_keys = torch.from_numpy(keys.astype(np.int32)).cuda()
                    pred.set_graph_pointers(d_off.data_ptr(), d_keys.data_ptr(), len(off) - 1, device=True, keep=(d_off, d_keys))
                du, dv, _ = pred.generate_deletions(12345, B)
                if mode == "host":
                    pred.apply_deletions(du, dv)
                else:
                    pred.apply_deletions()                      # the batch still on the GPU
                o2, k2 = pred.fetch_graph()
                to, tk = g.apply_deletions(torch.from_numpy(off.astype(np.int64)), torch.from_numpy(keys.astype(np.int32)),
                                           torch.from_numpy(du.astype(np.int64)), torch.from_numpy(dv.astype(np.int64)))
                to, tk = g.to_numpy(to, tk)
                assert np.array_equal(o2, to) and np.array_equal(k2, tk), (name, mode)
                if name != "pp3k_symdup" and oracle.ref_batch_available():
                    ro, rk = oracle.ref_apply_deletions(off, keys, du, dv)
                    assert np.array_equal(o2, ro) and np.array_equal(k2, rk), (name, mode, "reference applyBatchUpdateOmpU")
                if mode == "device":
                    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), off) and np.array_equal(d_keys.cpu().numpy().view(np.uint32), keys)
                K = max(1, len(du) // 2)
                for m, D in (("JC", 4), ("AA", 16), ("CN", 0)):
                    r = pred.predict(m, D, max_edges=K)
                    got = pred.fetch(r["count"])
                    want = oracle.oracle_predict(o2, k2, m, D, max_edges=K)[:3]
                    assert parity.compare(got, want, "after apply %s %s %s D=%d" % (name, mode, m, D)) is None
        # a second batch on top of the first (the handle's own graph is the source now), repeated requests, absent pairs
        off, keys = cases["rmat12"]
        pred.set_graph(off, keys)
        du, dv, _ = pred.generate_deletions(1, 400)
        pred.apply_deletions(du, dv)
        o2, k2 = pred.fetch_graph()
        du2, dv2, _ = pred.generate_deletions(2, 400)
        du2 = np.concatenate([du2, du2[:10], np.array([1, 5], np.uint32)]); dv2 = np.concatenate([dv2, dv2[:10], np.array([1, 4000], np.uint32)])
        pred.apply_deletions(du2, dv2)
        o3, k3 = pred.fetch_graph()
        to, tk = g.apply_deletions(torch.from_numpy(o2.astype(np.int64)), torch.from_numpy(k2.astype(np.int32)),
                                   torch.from_numpy(du2.astype(np.int64)), torch.from_numpy(dv2.astype(np.int64)))
        to, tk = g.to_numpy(to, tk)
        assert np.array_equal(o3, to) and np.array_equal(k3, tk)
        # the batch loop of main.cxx:163-169: checkpoint the base, per batch roll back and apply
        for mode in ("host", "device"):
            if mode == "host":
                pred.set_graph(off, keys)
            else:
                d_off = torch.from_numpy(off.astype(np.int64)).cuda(); d_keys = torch.from_numpy(keys.astype(np.int32)).cuda()
                pred.set_graph_pointers(d_off.data_ptr(), d_keys.data_ptr(), len(off) - 1, device=True, keep=(d_off, d_keys))
            pred.graph_checkpoint()
            for seed in (5, 6, 7):
                pred.graph_rollback()
                o0, k0 = pred.fetch_graph()
                assert np.array_equal(o0, off) and np.array_equal(k0, keys), (mode, seed, "rollback")
                du, dv, _ = pred.generate_deletions(seed, 300)
                pred.apply_deletions(du, dv)
                if seed == 6:                       # two batches in a row (BATCH_LENGTH = 2), still from this base
                    du_b, dv_b, _ = pred.generate_deletions(seed + 100, 200)
                    pred.apply_deletions(du_b, dv_b)
                o1, k1 = pred.fetch_graph()
                to, tk = g.apply_deletions(torch.from_numpy(off.astype(np.int64)), torch.from_numpy(keys.astype(np.int32)),
                                           torch.from_numpy(du.astype(np.int64)), torch.from_numpy(dv.astype(np.int64)))
                if seed == 6:
                    to, tk = g.apply_deletions(to, tk, torch.from_numpy(du_b.astype(np.int64)), torch.from_numpy(dv_b.astype(np.int64)))
                to, tk = g.to_numpy(to, tk)
                assert np.array_equal(o1, to) and np.array_equal(k1, tk), (mode, seed)
                r = pred.predict("JC", 4, max_edges=500)
                want = oracle.oracle_predict(o1, k1, "JC", 4, max_edges=500)[:3]
                assert parity.compare(pred.fetch(r["count"]), want, "batch loop %s %d" % (mode, seed)) is None
    finally:
        pred.close()

"""Host logic of the synthetic workloads (CPU): generators are deterministic, give the graph shape
the reference's loader leaves (symmetric, loop-free, sorted unique rows, vertex 0 empty), the
multiset variants behave as documented, and the .mtx writer round-trips."""
import numpy as np
import torch


def rows(off, keys):
    return [keys[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def is_symmetric_multiset(off, keys):
    S = len(off) - 1
    src = np.repeat(np.arange(S, dtype=np.int64), np.diff(off).astype(np.int64))
    a = np.sort(src * S + keys.astype(np.int64))
    b = np.sort(keys.astype(np.int64) * S + src)
    return bool((a == b).all())


def test_generators_are_deterministic_and_well_formed(nlp):
    g = nlp.graphs
    for make in (lambda: g.rmat(10, 8, 5), lambda: g.rmat(10, 8, 5, permute=True), lambda: g.road_lattice(30, 0.6, 6),
                 lambda: g.web_crawl(3000, 8, window=100, seed=7), lambda: g.planted_partition(1500, 30, 8, 2, 8)):
        o1, k1 = g.to_numpy(*make())
        o2, k2 = g.to_numpy(*make())
        assert (o1 == o2).all() and (k1 == k2).all()
        assert o1[0] == 0 and o1[1] == 0                      # vertex 0 is empty (ids are 1-based, mtx.hxx:240)
        for u, r in enumerate(rows(o1, k1)):
            assert (np.diff(r.astype(np.int64)) > 0).all()    # sorted, no duplicates
            assert not (r == u).any()                         # no self-loops
        assert is_symmetric_multiset(o1, k1)


def test_multiset_variants(nlp):
    g = nlp.graphs
    base = g.planted_partition(1500, 30, 8, 2, 8)
    o, k = g.to_numpy(*g.duplicate_symmetric(*base, every=4, copies=3))
    assert len(k) > base[1].numel() and is_symmetric_multiset(o, k)
    for r in rows(o, k):
        assert (np.diff(r.astype(np.int64)) >= 0).all()       # still sorted
    o, k = g.to_numpy(*g.duplicate_some_entries(*base, every=4))
    assert not is_symmetric_multiset(o, k)                    # the asymmetric kind


def test_remove_edges_and_mtx_round_trip(nlp, tmp_path):
    g = nlp.graphs
    off, keys = g.rmat(9, 8, 3)
    o2, k2, lo, hi = g.remove_edges(off, keys, 0.1, 11)
    assert k2.numel() + 2 * lo.numel() == keys.numel()
    assert (lo < hi).all()
    n = g.write_mtx(str(tmp_path / "g.mtx"), o2, k2)
    assert 2 * n == k2.numel()
    lines = open(tmp_path / "g.mtx").read().splitlines()
    assert lines[0].startswith("%%MatrixMarket matrix coordinate pattern symmetric")
    r, c, m = map(int, lines[1].split())
    assert r == c == off.numel() - 2 and m == n == len(lines) - 2
    e = np.array([[int(x) for x in l.split()] for l in lines[2:]], dtype=np.int64)
    back = g.csr_from_pairs(torch.from_numpy(e[:, 0]), torch.from_numpy(e[:, 1]), r)
    assert torch.equal(back[0], o2) and torch.equal(back[1], k2)


def test_partition_owner_matches_kernel_rule(nlp):
    # blocks of 32 consecutive ids dealt round-robin (owns_row_block in csrc/frontier.cuh)
    d = nlp.distributed
    assert [d.owner_of_vertex(u, 4) for u in (0, 31, 32, 63, 64, 127, 128)] == [0, 0, 1, 1, 2, 3, 0]

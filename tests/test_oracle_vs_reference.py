"""CPU: the C oracle against the compiled, unmodified reference (oracle/_ref/libnlpref.so) on
seeded random graphs -- sequential templates for exhaustive lists, OpenMP templates for top-K
(score multiset + timing fields).  Skipped where the prebuilt reference library is absent."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.skipif(
    not __import__("oracle.oracle_py", fromlist=["x"]).ref_available(),
    reason="oracle/_ref/libnlpref.so not built (needs /root/reference)")


def _graphs():
    import nlp_b200 as N
    g = N.graphs
    return {
        "rmat9": g.to_numpy(*g.rmat(9, 8, 21)),
        "road20": g.to_numpy(*g.road_lattice(20, 0.7, 22)),
        "pp500": g.to_numpy(*g.planted_partition(500, 20, 6, 2, 23)),
        "web3k": g.to_numpy(*g.web_crawl(3000, 8, window=200, seed=24)),
        "pp500_multiset": g.to_numpy(*g.duplicate_some_entries(*g.planted_partition(500, 20, 6, 2, 25), every=3)),
    }


@pytest.mark.parametrize("name", ["rmat9", "road20", "pp500", "web3k", "pp500_multiset"])
def test_exhaustive_lists_equal_reference(oracle, name):
    off, keys = _graphs()[name]
    R = oracle.RefGraph(off, keys)
    for m in oracle.MEASURES:
        for D in (0, 2, 8, 64, 1024):
            want = R.predict(m, D, omp=False)[:3]
            got = oracle.oracle_predict(off, keys, m, D)[:3]
            assert parity.compare(got, want, "%s %s D=%d" % (name, m, D)) is None


@pytest.mark.parametrize("measure", ["CN", "JC", "SC", "AA", "RA"])
def test_topk_scores_equal_reference_openmp(oracle, measure):
    """Ties are broken differently (reference: heap accident), so compare the score multiset
    of the K best and that every reference edge has the oracle's score."""
    off, keys = _graphs()["rmat9"]
    R = oracle.RefGraph(off, keys)
    full = oracle.oracle_predict(off, keys, measure, 0)
    K = len(full[0]) // 10
    assert K > 10
    ru, rv, rs, t, ts = R.predict(measure, 0, max_edges=K, omp=True, threads=4)
    ou, ov, os_, _ = oracle.oracle_predict(off, keys, measure, 0, max_edges=K)
    assert len(ru) == K == len(ou)
    assert np.array_equal(np.sort(rs.view(np.uint32)), np.sort(os_.view(np.uint32)))
    lookup = {(int(a), int(b)): int(c) for a, b, c in zip(full[0], full[1], full[2].view(np.uint32))}
    for a, b, c in zip(ru, rv, rs.view(np.uint32)):
        assert lookup[(int(a), int(b))] == int(c)
    assert t >= ts >= 0.0


def test_tie_straddling_cutoff_bounds_f1_difference(oracle):
    """INTEGRATION.md, "Ties": with integer-valued common-neighbour scores the K-th score sits in
    a large tie class; the reference keeps whichever tied pairs its heap saw last
    (inc/predict.hxx:332), the canonical order keeps the smallest (u, v).  Pin what that can change:
    both lists agree on every pair scored above the cutoff, hold the same number of pairs AT the
    cutoff score, and so precision / recall (main.cxx:201-202) differ by at most the number of tied
    pairs in the list over the list sizes."""
    import nlp_b200 as N
    g = N.graphs
    off0, keys0 = g.planted_partition(1500, 30, 8, 2, 91)
    off, keys, lo, hi = g.remove_edges(off0, keys0, 0.1, 92)
    offn, keysn = g.to_numpy(off, keys)
    lo, hi = lo.numpy(), hi.numpy()
    K = len(lo)
    S = len(offn) - 1
    R = oracle.RefGraph(offn, keysn)
    ru, rv, rs, _, _ = R.predict("CN", 0, max_edges=K, omp=False, canonical=False)
    cu, cv, cs, _ = oracle.oracle_predict(offn, keysn, "CN", 0, max_edges=K)
    assert len(ru) == len(cu) == K
    cut = float(cs[-1])
    assert float(rs.min()) == cut
    tied = int((cs == cut).sum())
    assert tied > 1 and int((rs == cut).sum()) == tied          # the cutoff really straddles a tie class
    above_r = set(zip(ru[rs > cut].tolist(), rv[rs > cut].tolist()))
    above_c = set(zip(cu[cs > cut].tolist(), cv[cs > cut].tolist()))
    assert above_r == above_c
    truth = set(zip(lo.tolist(), hi.tolist()))
    hit_r = sum((a, b) in truth for a, b in zip(ru.tolist(), rv.tolist()))
    hit_c = sum((a, b) in truth for a, b in zip(cu.tolist(), cv.tolist()))
    assert abs(hit_r - hit_c) <= tied                             # precision = recall = hits / K here (K predictions, K removed)

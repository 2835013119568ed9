"""F1 parity (BASELINE.json metric 4): remove edges, predict exactly that many, score the
prediction against the removed edges the way main.cxx:48-57,199-206 does (both directions,
sort, unique, set_intersection -> precision, recall; F1 from them).  The GPU path and the
reference-pinned oracle must give the SAME F1, exactly, because the predicted edge sets are
identical under the canonical (score desc, u, v) order."""
import numpy as np
import pytest


def f1_of(u, v, removed_lo, removed_hi, span):
    pred = np.unique(np.minimum(u, v).astype(np.int64) * span + np.maximum(u, v).astype(np.int64))
    truth = np.unique(removed_lo.astype(np.int64) * span + removed_hi.astype(np.int64))
    common = np.intersect1d(pred, truth, assume_unique=True).size
    precision = common / max(pred.size, 1)
    recall = common / max(truth.size, 1)
    f1 = 0.0 if common == 0 else 2 * precision * recall / (precision + recall)
    return precision, recall, f1


def workload(nlp, kind):
    g = nlp.graphs
    if kind == "clustered":      # planted partition: link prediction has a non-trivial F1 here
        off, keys = g.planted_partition(4000, 200, 10, 1, 61)
    else:
        off, keys = g.rmat(12, 16, 62)
    off2, keys2, lo, hi = g.remove_edges(off, keys, 0.1, 63)
    o, k = g.to_numpy(off2, keys2)
    return o, k, lo.numpy(), hi.numpy()


def test_oracle_f1_is_nontrivial_on_clustered_graph(nlp, oracle):
    o, k, lo, hi = workload(nlp, "clustered")
    u, v, s, st = oracle.oracle_predict(o, k, "JC", 0, max_edges=len(lo))
    p, r, f1 = f1_of(u, v, lo, hi, len(o))
    assert f1 > 0.1, f1          # R-MAT gives ~1e-4; this graph has structure


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["clustered", "rmat"])
def test_f1_matches_exactly(nlp, oracle, kind):
    o, k, lo, hi = workload(nlp, kind)
    pred = nlp.Predictor(0)
    try:
        pred.set_graph(o, k)
        for measure in nlp.MEASURES:
            for D in (0, 4, 32):
                r = pred.predict(measure, D, max_edges=len(lo))
                gu, gv, gs = pred.fetch(r["count"])
                wu, wv, ws, st = oracle.oracle_predict(o, k, measure, D, max_edges=len(lo))
                got = f1_of(gu, gv, lo, hi, len(o))
                want = f1_of(wu, wv, lo, hi, len(o))
                assert got == want, (kind, measure, D, got, want)
    finally:
        pred.close()

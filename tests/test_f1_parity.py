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


def directed_truth(lo, hi):
    """main.cxx:206-207: both directions of every removed edge, sorted by (u, v)."""
    tu = np.concatenate([lo, hi]).astype(np.uint32)
    tv = np.concatenate([hi, lo]).astype(np.uint32)
    o = np.lexsort((tv, tu))
    return tu[o], tv[o]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["clustered", "rmat"])
def test_device_evaluation_matches_host(nlp, kind):
    """nlp_set_truth + nlp_evaluate (precision / recall on the device, main.cxx:48-57,201-202)
    against the host computation on the fetched edges."""
    o, k, lo, hi = workload(nlp, kind)
    tu, tv = directed_truth(lo, hi)
    pred = nlp.Predictor(0)
    try:
        pred.set_graph(o, k)
        with pytest.raises(nlp.NlpError) as e:
            pred.predict("JC", 4, max_edges=10)
            pred.evaluate()
        assert e.value.code == 6                       # NLP_ERR_NO_TRUTH
        with pytest.raises(nlp.NlpError) as e:
            pred.set_truth(tu[::-1], tv[::-1])
        assert e.value.code == 1                       # not sorted
        pred.set_truth(tu, tv)
        for measure in ("CN", "JC", "AA"):
            for D in (0, 4, 32):
                for K in (len(lo), len(lo) // 3, 1):
                    r = pred.predict(measure, D, max_edges=K)
                    ev = pred.evaluate()
                    gu, gv, gs = pred.fetch(r["count"])
                    p, rc, f1 = f1_of(gu, gv, lo, hi, len(o))
                    assert ev["predicted"] == 2 * r["count"] and ev["truth"] == 2 * len(lo)
                    assert ev["precision"] == p and ev["recall"] == rc, (kind, measure, D, K, ev, p, rc)
        # a truth list with a repeated entry and ids outside the graph still matches once per edge
        tu2 = np.concatenate([tu[:1], tu, [np.uint32(len(o) + 5)]]).astype(np.uint32)
        tv2 = np.concatenate([tv[:1], tv, [np.uint32(1)]]).astype(np.uint32)
        pred.set_truth(tu2, tv2)
        r = pred.predict("JC", 0, max_edges=len(lo))
        ev2 = pred.evaluate()
        gu, gv, gs = pred.fetch(r["count"])
        p, rc, f1 = f1_of(gu, gv, lo, hi, len(o))
        assert ev2["truth"] == len(tu2) and ev2["common"] == round(p * 2 * r["count"])
        # empty truth / empty result
        pred.set_truth(np.empty(0, np.uint32), np.empty(0, np.uint32))
        ev3 = pred.evaluate()
        assert ev3["common"] == 0 and ev3["recall"] == 0.0 and ev3["precision"] == 0.0
    finally:
        pred.close()

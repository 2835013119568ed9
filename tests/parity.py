"""Shared helpers of the parity tests: build small graphs, run GPU path + oracle, compare."""
import numpy as np


def kat_graph():
    """The 7-vertex known-answer graph of SURVEY.md section 8c."""
    edges = [(1, 3), (2, 3), (1, 4), (2, 4), (2, 5), (5, 6), (6, 7), (4, 7)]
    return csr_from_edge_list(edges, 8)


def csr_from_edge_list(edges, span, symmetric=True):
    adj = [[] for _ in range(span)]
    for a, b in edges:
        adj[a].append(b)
        if symmetric:
            adj[b].append(a)
    off = np.zeros(span + 1, np.uint64)
    keys = []
    for u in range(span):
        adj[u].sort()
        keys += adj[u]
        off[u + 1] = len(keys)
    return off, np.array(keys, np.uint32)


def compare(got, want, what=""):
    """Bit-exact comparison of (u, v, score) triples; returns an error string or None."""
    gu, gv, gs = got
    wu, wv, ws = want
    if len(gu) != len(wu):
        return "%s: count %d != %d" % (what, len(gu), len(wu))
    if len(gu) == 0:
        return None
    bad = (gu != wu) | (gv != wv) | (gs.view(np.uint32) != ws.view(np.uint32))
    if bad.any():
        i = int(np.argmax(bad))
        return "%s: %d/%d rows differ, first at %d: got (%d,%d,%r/0x%08x) want (%d,%d,%r/0x%08x)" % (
            what, int(bad.sum()), len(gu), i, gu[i], gv[i], float(gs[i]), gs.view(np.uint32)[i],
            wu[i], wv[i], float(ws[i]), ws.view(np.uint32)[i])
    return None


def check_case(pred, oracle_py, off, keys, measure, D, max_edges, min_score=0.0, max_factor2=0, tag=""):
    """Run one (measure, D, K) case on the GPU through the C ABI and on the oracle."""
    r = pred.predict(measure, D, max_edges=max_edges, min_score=min_score, max_factor2=max_factor2)
    got = pred.fetch(r["count"])
    wu, wv, ws, st = oracle_py.oracle_predict(off, keys, measure, D, max_edges=max_edges,
                                              min_score=min_score, max_factor2=max_factor2)
    what = "%s %s D=%d K=%s" % (tag, measure, D, "all" if max_edges >= (1 << 63) else max_edges)
    err = compare(got, (wu, wv, ws), what)
    if err is None:
        for k in ("first_hop", "eligible_first_hop", "wedges", "candidates", "kept"):
            if r[k] != st[k]:
                err = "%s: counter %s = %d, oracle %d" % (what, k, r[k], st[k])
                break
    return err, r, st

"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle and the compiled reference.

* ``oracle_predict``  -> oracle/liboracle.so   (plain-C restatement, oracle/nlp_oracle.c)
* ``RefGraph``        -> oracle/_ref/libnlpref.so (UNMODIFIED reference templates compiled
                         from /root/reference by oracle/Makefile; see oracle/ref_driver.cxx)
* ``oracle_edge_deletions`` / ``ref_edge_deletions`` -> the random edge removal that precedes the
                         prediction (inc/batch.hxx), restated in oracle/batch_oracle.c and wrapped
                         from the reference in oracle/ref_batch_driver.cxx

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MEASURES = ["CN", "JC", "SI", "SC", "HP", "HD", "LHN", "AA", "RA"]   # main.cxx:212-220 order
REF_DEGREES = [0, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]          # main.cxx:67-80
UNBOUNDED = (1 << 64) - 1


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("first_hop", "eligible_first_hop", "wedges", "wedges_vgtu", "candidates", "kept")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(reference=True):
    """Compile the oracle (and, where /root/reference exists, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if reference:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_oracle = None
_ref = None


def _load_oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(reference=False)
        lib = C.CDLL(path)
        lib.nlp_oracle_predict.restype = C.c_int
        lib.nlp_oracle_predict.argtypes = [
            C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64,
            C.c_float, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
            C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(OracleStats)]
        lib.nlp_oracle_free.argtypes = [C.c_void_p]
        lib.nlp_oracle_edge_deletions.restype = C.c_int
        lib.nlp_oracle_edge_deletions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64,
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.nlp_oracle_batch_free.argtypes = [C.c_void_p]
        _oracle = lib
    return _oracle


def _copy_out(ptr, n, dtype):
    if n == 0:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def oracle_predict(offsets, keys, measure, min_degree1=4, max_edges=UNBOUNDED, min_score=0.0,
                   max_factor2=0, threads=0):
    """Run the C oracle.  Returns (u, v, score, stats-dict) in canonical (score desc, u, v) order."""
    lib = _load_oracle()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    span = offsets.shape[0] - 1
    if isinstance(measure, str):
        measure = MEASURES.index(measure)
    pu, pv, ps = C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = C.c_uint64(0)
    st = OracleStats()
    kp = keys.ctypes.data if keys.size else None
    rc = lib.nlp_oracle_predict(offsets.ctypes.data, kp, span, measure, min_degree1, max_factor2,
                                C.c_uint64(max_edges), C.c_float(min_score), threads,
                                C.byref(pu), C.byref(pv), C.byref(ps), C.byref(n), C.byref(st))
    if rc != 0:
        raise MemoryError("nlp_oracle_predict failed")
    k = int(n.value)
    u = _copy_out(pu.value, k, np.uint32)
    v = _copy_out(pv.value, k, np.uint32)
    s = _copy_out(ps.value, k, np.float32)
    for p in (pu, pv, ps):
        lib.nlp_oracle_free(p)
    return u, v, s, st.as_dict()


def oracle_edge_deletions(offsets, keys, seed, batch_size):
    """The C restatement of generateEdgeDeletions + tidyBatchUpdateU (inc/batch.hxx:99-112, 200-208)
    with std::default_random_engine(seed).  Returns (u, v, engine_words_consumed): the sorted unique
    directed list of removed edges."""
    lib = _load_oracle()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    pu, pv = C.c_void_p(), C.c_void_p()
    n, words = C.c_uint64(0), C.c_uint64(0)
    rc = lib.nlp_oracle_edge_deletions(offsets.ctypes.data, keys.ctypes.data if keys.size else None,
                                       offsets.shape[0] - 1, seed, batch_size,
                                       C.byref(pu), C.byref(pv), C.byref(n), C.byref(words))
    if rc != 0:
        raise MemoryError("nlp_oracle_edge_deletions failed")
    u = _copy_out(pu.value, int(n.value), np.uint32)
    v = _copy_out(pv.value, int(n.value), np.uint32)
    lib.nlp_oracle_batch_free(pu); lib.nlp_oracle_batch_free(pv)
    return u, v, int(words.value)


def ref_batch_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libnlpref_batch.so"))


def ref_edge_deletions(offsets, keys, seed, batch_size):
    """The reference's own generateEdgeDeletions + tidyBatchUpdateU on a DiGraph built from the CSR."""
    lib = C.CDLL(os.path.join(_HERE, "_ref", "libnlpref_batch.so"))
    lib.nlpref_edge_deletions.restype = C.c_int64
    lib.nlpref_edge_deletions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64,
                                          C.c_void_p, C.c_void_p, C.c_uint64]
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    cap = 2 * batch_size + 2
    u = np.empty(cap, np.uint32); v = np.empty(cap, np.uint32)
    n = lib.nlpref_edge_deletions(offsets.ctypes.data, keys.ctypes.data if keys.size else None,
                                  offsets.shape[0] - 1, seed, batch_size, u.ctypes.data, v.ctypes.data, cap)
    if n < 0:
        raise ValueError("nlpref_edge_deletions: capacity")
    return u[:n].copy(), v[:n].copy()


def ref_apply_deletions(offsets, keys, del_u, del_v):
    """The reference's own applyBatchUpdateOmpU (inc/batch.hxx:239-247) on a DiGraph built from the
    CSR; returns the CSR of the graph afterwards."""
    lib = C.CDLL(os.path.join(_HERE, "_ref", "libnlpref_batch.so"))
    lib.nlpref_apply_deletions.restype = C.c_int64
    lib.nlpref_apply_deletions.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64,
                                           C.c_void_p, C.c_void_p]
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    du = np.ascontiguousarray(del_u, dtype=np.uint32); dv = np.ascontiguousarray(del_v, dtype=np.uint32)
    o2 = np.empty_like(offsets); k2 = np.empty(max(keys.size, 1), np.uint32)
    n = lib.nlpref_apply_deletions(offsets.ctypes.data, keys.ctypes.data if keys.size else None, offsets.shape[0] - 1,
                                   du.ctypes.data if du.size else None, dv.ctypes.data if dv.size else None, du.size,
                                   o2.ctypes.data, k2.ctypes.data)
    return o2, k2[:n].copy()


def ref_read_mtx(path, symmetric=False, drop_self_loops=True):
    """The reference's own ingest (main.cxx:243-245: readMtxOmpW, symmetrizeOmp unless `symmetric`,
    removeSelfLoopsOmpU) of the Matrix Market file at `path`; returns its graph as CSR."""
    lib = C.CDLL(os.path.join(_HERE, "_ref", "libnlpref_batch.so"))
    lib.nlpref_read_mtx.restype = C.c_int64
    lib.nlpref_read_mtx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    span = C.c_uint32(0)
    m = lib.nlpref_read_mtx(path.encode(), int(symmetric), int(drop_self_loops), C.byref(span), None, None)
    off = np.empty(span.value + 1, np.uint64); keys = np.empty(max(m, 1), np.uint32)
    lib.nlpref_read_mtx(path.encode(), int(symmetric), int(drop_self_loops), C.byref(span), off.ctypes.data, keys.ctypes.data)
    return off, keys[:m].copy()


def _set_union_last(x, y):
    """Literal port of the reference's set_union_last_inplace (inc/_algorithm.hxx:177-221) for sorted
    lists of ints: y merged into x.  NOT a set union -- once the routine has left its first loop, an
    element that is in both lists is written twice (the copy waiting in the deque follows the one
    taken from y)."""
    if not y:
        return list(x)
    if not x:
        return sorted(set(y))
    x = list(x); y = list(y)
    xb = 0; yb = 0
    while True:
        while x[xb] < y[yb]:
            xb += 1
            if xb == len(x):
                return x + sorted(set(y[yb:]))
        if x[xb] != y[yb]:
            break
        x[xb] = y[yb]; yb += 1
        if yb == len(y):
            return x
    out = x[:xb]                       # everything before the insertion point stays
    q = [x[xb]]; xb += 1               # deque of displaced x elements
    out.append(y[yb]); yb += 1
    while yb < len(y):
        if out[-1] == y[yb]:
            out[-1] = y[yb]; yb += 1
        else:
            if xb < len(x):
                q.append(x[xb]); xb += 1
            if q and q[0] < y[yb]:
                out.append(q.pop(0))
            else:
                out.append(y[yb]); yb += 1
    while True:
        if xb < len(x):
            q.append(x[xb]); xb += 1
        if not q:
            break
        out.append(q.pop(0))
    return out


def mtx_to_csr(text, symmetrize=True, drop_self_loops=True):
    """Restatement of the reference's ingest for Matrix Market coordinate TEXT (bytes): header as
    inc/mtx.hxx:38-55, body lines "u v [w]" (inc/mtx.hxx:173-179; both directions for a symmetric
    banner), rows sorted and made unique by the first update, then -- with `symmetrize`
    (main.cxx:244, inc/symmetrize.hxx:71-82) -- the reverse of every stored edge merged in by the
    reference's own set_union_last_inplace (which leaves some common entries twice, see
    _set_union_last), then one copy of every self-loop removed (inc/selfLoop.hxx:117-124,
    set_difference_inplace).  Returns (offsets, keys) with span = max(rows, cols) + 1."""
    lines = text.decode().split("\n")
    sym = False
    i = 0
    while True:
        ln = lines[i]; i += 1
        if not ln.startswith("%"):
            break
        if ln.startswith("%%"):
            h = ln.split()
            assert h[1] == "matrix" and h[2] == "coordinate"
            sym = len(h) > 4 and h[4] in ("symmetric", "skew-symmetric")
    rows, cols, _ = (int(x) for x in ln.split()[:3])
    n = max(rows, cols)
    stored = [set() for _ in range(n + 1)]
    for ln in lines[i:]:
        t = ln.split()
        if len(t) < 2 or ln.lstrip().startswith("%"):
            continue
        u, v = int(t[0]), int(t[1])
        stored[u].add(v)
        if sym:
            stored[v].add(u)
    x = [sorted(r) for r in stored]
    if symmetrize:
        rev = [[] for _ in range(n + 1)]
        for u in range(n + 1):
            for v in x[u]:
                rev[v].append(u)          # ascending u: already sorted
        x = [_set_union_last(x[u], rev[u]) for u in range(n + 1)]
    if drop_self_loops:
        for u in range(n + 1):
            if u in x[u]:
                x[u].remove(u)            # one copy
    off = np.zeros(n + 2, np.uint64)
    np.cumsum([len(r) for r in x], out=off[1:])
    keys = np.array([v for r in x for v in r], np.uint32)
    return off, keys


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libnlpref.so"))


def _load_ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(_HERE, "_ref", "libnlpref.so"))
        lib.nlpref_graph_create.restype = C.c_void_p
        lib.nlpref_graph_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        lib.nlpref_graph_destroy.argtypes = [C.c_void_p]
        lib.nlpref_predict.restype = C.c_int64
        lib.nlpref_predict.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                       C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.nlpref_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.nlpref_max_threads.restype = C.c_int
        _ref = lib
    return _ref


def canonical_order(u, v, s):
    """Indices that sort (u, v, score) into the canonical (score desc, u asc, v asc) order."""
    return np.lexsort((v, u, -s.astype(np.float64)))


class RefGraph:
    """The reference's own DiGraphCsr + predictLinks* templates (compiled, unmodified)."""

    def __init__(self, offsets, keys):
        self.lib = _load_ref()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        self.h = self.lib.nlpref_graph_create(offsets.ctypes.data,
                                              keys.ctypes.data if keys.size else None,
                                              offsets.shape[0] - 1)

    def max_threads(self):
        return int(self.lib.nlpref_max_threads())

    def predict(self, measure, min_degree1, max_edges=UNBOUNDED, min_score=0.0, repeat=1,
                omp=False, threads=0, canonical=True):
        """Returns (u, v, score, time_ms, scoring_ms).  ``omp=True`` must only be used when
        #candidates >= max_edges (reference UB otherwise, inc/predict.hxx:424,452-453)."""
        if isinstance(measure, str):
            measure = MEASURES.index(measure)
        t, ts = C.c_float(0), C.c_float(0)
        n = self.lib.nlpref_predict(self.h, measure, min_degree1, int(omp), threads,
                                    C.c_uint64(max_edges), C.c_float(min_score), repeat,
                                    C.byref(t), C.byref(ts))
        if n < 0:
            raise ValueError("reference has no instantiation for measure=%r D=%r" % (measure, min_degree1))
        u = np.empty(n, np.uint32); v = np.empty(n, np.uint32); s = np.empty(n, np.float32)
        if n:
            self.lib.nlpref_fetch(self.h, u.ctypes.data, v.ctypes.data, s.ctypes.data)
        if canonical and n:
            o = canonical_order(u, v, s)
            u, v, s = u[o], v[o], s[o]
        return u, v, s, float(t.value), float(ts.value)

    def close(self):
        if self.h:
            self.lib.nlpref_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

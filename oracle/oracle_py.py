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

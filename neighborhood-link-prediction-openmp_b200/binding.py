"""ctypes binding of the C ABI in include/nlp_b200.h (the product path).

This module never imports anything under oracle/ and has no CPU fallback: if the CUDA library
is missing or no B200 is present it raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MEASURES = ["CN", "JC", "SI", "SC", "HP", "HD", "LHN", "AA", "RA"]   # main.cxx:212-220 order
UNBOUNDED = (1 << 64) - 1
STATUS = {0: "NLP_OK", 1: "NLP_ERR_ARG", 2: "NLP_ERR_CUDA", 3: "NLP_ERR_NO_GRAPH",
          4: "NLP_ERR_CAPACITY", 5: "NLP_ERR_NO_RESULT", 6: "NLP_ERR_NO_TRUTH", 7: "NLP_ERR_COMM"}

EXPORTS = ["nlp_create", "nlp_destroy", "nlp_set_graph", "nlp_set_graph_device", "nlp_set_partition",
           "nlp_set_scratch_limit", "nlp_set_path", "nlp_fetch_async", "nlp_fetch_wait", "nlp_set_reuse", "nlp_predict", "nlp_fetch", "nlp_result_device", "nlp_merge",
           "nlp_comm_unique_id", "nlp_comm_init", "nlp_comm_destroy", "nlp_comm_bytes",
           "nlp_set_truth", "nlp_evaluate", "nlp_generate_deletions", "nlp_fetch_deletions", "nlp_deletions_device", "nlp_apply_deletions", "nlp_graph_checkpoint", "nlp_graph_rollback", "nlp_graph_size", "nlp_fetch_graph", "nlp_ingest_mtx", "nlp_launch_count", "nlp_stream", "nlp_last_error", "nlp_version"]


class Options(C.Structure):
    _fields_ = [("measure", C.c_int32), ("min_degree1", C.c_uint32), ("max_factor2", C.c_uint32),
                ("repeat", C.c_int32), ("max_edges", C.c_uint64), ("min_score", C.c_float)]


class Result(C.Structure):
    _fields_ = [("count", C.c_uint64), ("time_ms", C.c_float), ("scoring_ms", C.c_float),
                ("select_ms", C.c_float), ("frontier_ms", C.c_float), ("first_hop", C.c_uint64),
                ("eligible_first_hop", C.c_uint64), ("wedges", C.c_uint64), ("candidates", C.c_uint64),
                ("kept", C.c_uint64), ("emitted", C.c_uint64), ("frontier_sources", C.c_uint64),
                ("bin_sources", C.c_uint64 * 8), ("passes", C.c_uint32), ("path", C.c_uint32),
                ("phase_ms", C.c_float * 8), ("pair_records", C.c_uint64)]

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            x = getattr(self, n)
            d[n] = list(x) if n in ("bin_sources", "phase_ms") else (float(x) if n.endswith("_ms") else int(x))
        return d


class Evaluation(C.Structure):
    _fields_ = [("predicted", C.c_uint64), ("truth", C.c_uint64), ("common", C.c_uint64),
                ("precision", C.c_double), ("recall", C.c_double), ("ms", C.c_float)]

    def as_dict(self):
        return {"predicted": int(self.predicted), "truth": int(self.truth), "common": int(self.common),
                "precision": float(self.precision), "recall": float(self.recall), "ms": float(self.ms)}


class NlpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS.get(code, code), msg))
        self.code = code


_lib = None


def load_library(build_if_missing=True):
    """dlopen libnlp_b200.so (building it with nvcc first if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(path)
        _build.build()
    lib = C.CDLL(path)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.nlp_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.nlp_destroy.argtypes = [vp]
    lib.nlp_set_graph.argtypes = [vp, vp, vp, u32]
    lib.nlp_set_graph_device.argtypes = [vp, vp, vp, u32]
    lib.nlp_set_partition.argtypes = [vp, C.c_int, C.c_int]
    lib.nlp_set_scratch_limit.argtypes = [vp, u64]
    lib.nlp_comm_unique_id.argtypes = [vp]
    lib.nlp_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.nlp_comm_destroy.argtypes = [vp]
    lib.nlp_comm_bytes.argtypes = [vp]
    lib.nlp_comm_bytes.restype = u64
    lib.nlp_set_path.argtypes = [vp, C.c_int]
    lib.nlp_set_reuse.argtypes = [vp, C.c_int]
    lib.nlp_predict.argtypes = [vp, C.POINTER(Options), C.POINTER(Result)]
    lib.nlp_fetch.argtypes = [vp, vp, vp, vp, u64]
    lib.nlp_fetch_async.argtypes = [vp, vp, vp, vp, u64]
    lib.nlp_fetch_wait.argtypes = [vp]
    lib.nlp_result_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    lib.nlp_merge.argtypes = [vp, vp, vp, vp, u64, u64, C.POINTER(C.c_float)]
    lib.nlp_set_truth.argtypes = [vp, vp, vp, u64]
    lib.nlp_evaluate.argtypes = [vp, C.POINTER(Evaluation)]
    lib.nlp_generate_deletions.argtypes = [vp, u32, u64, C.POINTER(u64), C.POINTER(u64)]
    lib.nlp_fetch_deletions.argtypes = [vp, vp, vp, u64]
    lib.nlp_deletions_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    lib.nlp_apply_deletions.argtypes = [vp, vp, vp, u64]
    lib.nlp_graph_checkpoint.argtypes = [vp]
    lib.nlp_graph_rollback.argtypes = [vp]
    lib.nlp_graph_size.argtypes = [vp, C.POINTER(u32), C.POINTER(u64)]
    lib.nlp_fetch_graph.argtypes = [vp, vp, vp]
    lib.nlp_ingest_mtx.argtypes = [vp, C.c_char_p, u64, u32, C.POINTER(u32), C.POINTER(u64)]
    lib.nlp_launch_count.argtypes = [vp]
    lib.nlp_launch_count.restype = u64
    lib.nlp_stream.argtypes = [vp]
    lib.nlp_stream.restype = vp
    lib.nlp_last_error.argtypes = [vp]
    lib.nlp_last_error.restype = C.c_char_p
    lib.nlp_version.restype = C.c_char_p
    _lib = lib
    return lib


class Predictor:
    """One handle = one GPU.  Thin Python mirror of include/nlp_b200.h for tests and bench.py."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.nlp_create(C.byref(h), device)
        if rc != 0:
            raise NlpError(rc, self.lib.nlp_last_error(None).decode())
        self.h = h
        self._keep = None

    def _check(self, rc):
        if rc != 0:
            raise NlpError(rc, self.lib.nlp_last_error(self.h).decode())

    def set_graph(self, offsets, keys):
        """Host CSR arrays (numpy uint64 offsets[S+1], uint32 keys[M]); copied to the GPU."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        self._check(self.lib.nlp_set_graph(self.h, offsets.ctypes.data,
                                           keys.ctypes.data if keys.size else None,
                                           offsets.shape[0] - 1))

    def set_graph_pointers(self, offsets_ptr, keys_ptr, span, device=False, keep=None):
        """Raw pointers (host, e.g. pinned torch tensors, or device when ``device=True``)."""
        fn = self.lib.nlp_set_graph_device if device else self.lib.nlp_set_graph
        self._check(fn(self.h, offsets_ptr, keys_ptr, span))
        self._keep = keep

    def set_partition(self, rank, world):
        self._check(self.lib.nlp_set_partition(self.h, rank, world))

    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id (rank 0 creates it; hand it to the other ranks yourself)."""
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.nlp_comm_unique_id(buf)
        if rc != 0:
            raise NlpError(rc, lib.nlp_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid, rank, world):
        """Collective: join the NCCL communicator of `world` ranks; nlp_predict then merges across
        the ranks by itself and every rank holds the full result."""
        self._check(self.lib.nlp_comm_init(self.h, C.c_char_p(bytes(uid)), rank, world))

    def comm_destroy(self):
        self._check(self.lib.nlp_comm_destroy(self.h))

    def comm_bytes(self):
        return int(self.lib.nlp_comm_bytes(self.h))

    def set_path(self, path):
        """0 = auto, 1 = source-centric kernels, 2 = LHub pair path (when admissible)."""
        self._check(self.lib.nlp_set_path(self.h, path))

    def set_reuse(self, on):
        """Keep sorted wedge records per threshold for later measures (empties the store)."""
        self._check(self.lib.nlp_set_reuse(self.h, 1 if on else 0))

    def set_scratch_limit(self, nbytes):
        self._check(self.lib.nlp_set_scratch_limit(self.h, nbytes))

    def predict(self, measure, min_degree1=4, max_edges=UNBOUNDED, min_score=0.0, repeat=1, max_factor2=0):
        if isinstance(measure, str):
            measure = MEASURES.index(measure)
        o = Options(measure, min_degree1, max_factor2, repeat, max_edges, min_score)
        r = Result()
        self._check(self.lib.nlp_predict(self.h, C.byref(o), C.byref(r)))
        return r.as_dict()

    def fetch(self, count):
        u = np.empty(count, np.uint32); v = np.empty(count, np.uint32); s = np.empty(count, np.float32)
        if count:
            self._check(self.lib.nlp_fetch(self.h, u.ctypes.data, v.ctypes.data, s.ctypes.data, count))
        return u, v, s

    def fetch_into(self, u_ptr, v_ptr, s_ptr, capacity):
        """Copy the result into caller memory (host or device pointers)."""
        self._check(self.lib.nlp_fetch(self.h, u_ptr, v_ptr, s_ptr, capacity))

    def fetch_async(self, u_ptr, v_ptr, s_ptr, capacity):
        """Non-blocking fetch (see nlp_fetch_async); finish with fetch_wait()."""
        self._check(self.lib.nlp_fetch_async(self.h, u_ptr, v_ptr, s_ptr, capacity))

    def fetch_wait(self):
        self._check(self.lib.nlp_fetch_wait(self.h))

    def result_device(self):
        pu, pv, ps, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.nlp_result_device(self.h, C.byref(pu), C.byref(pv), C.byref(ps), C.byref(n)))
        return pu.value, pv.value, ps.value, int(n.value)

    def merge(self, u_ptr, v_ptr, s_ptr, n, max_edges):
        ms = C.c_float(0)
        self._check(self.lib.nlp_merge(self.h, u_ptr, v_ptr, s_ptr, n, max_edges, C.byref(ms)))
        return float(ms.value)

    def set_truth(self, u, v):
        """The held-back edges as main.cxx holds them: directed (u, v) pairs (both directions of
        every removed edge), sorted ascending by (u, v).  numpy arrays; copied to the GPU."""
        u = np.ascontiguousarray(u, dtype=np.uint32); v = np.ascontiguousarray(v, dtype=np.uint32)
        assert u.shape == v.shape
        self._check(self.lib.nlp_set_truth(self.h, u.ctypes.data if u.size else None,
                                           v.ctypes.data if v.size else None, u.size))

    def set_truth_pointers(self, u_ptr, v_ptr, n):
        self._check(self.lib.nlp_set_truth(self.h, u_ptr, v_ptr, n))

    def evaluate(self):
        """Precision / recall of the last result against the held-back edges (main.cxx:48-57,
        201-202), computed on the device."""
        e = Evaluation()
        self._check(self.lib.nlp_evaluate(self.h, C.byref(e)))
        return e.as_dict()

    def generate_deletions(self, seed, batch_size, fetch=True):
        """The reference's random edge removal (inc/batch.hxx:99-112, 200-208) on the resident graph
        for std::default_random_engine(seed).  Returns (u, v, engine_words) -- the sorted unique
        directed pairs -- or (count, engine_words) with ``fetch=False`` (pairs stay on the GPU)."""
        n, words = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.nlp_generate_deletions(self.h, seed, batch_size, C.byref(n), C.byref(words)))
        n = int(n.value)
        if not fetch:
            return n, int(words.value)
        u = np.empty(n, np.uint32); v = np.empty(n, np.uint32)
        if n:
            self._check(self.lib.nlp_fetch_deletions(self.h, u.ctypes.data, v.ctypes.data, n))
        return u, v, int(words.value)

    def fetch_deletions_into(self, u_ptr, v_ptr, capacity):
        """Copy the generated deletions into caller memory (host or device pointers)."""
        self._check(self.lib.nlp_fetch_deletions(self.h, u_ptr, v_ptr, capacity))

    def deletions_device(self):
        pu, pv, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.lib.nlp_deletions_device(self.h, C.byref(pu), C.byref(pv), C.byref(n)))
        return pu.value, pv.value, int(n.value)

    def apply_deletions(self, u=None, v=None, pointers=None):
        """Remove directed pairs from the resident graph on the device (nlp_apply_deletions).
        numpy arrays, or ``pointers=(u_ptr, v_ptr, n)`` (host or device); with neither, the batch
        nlp_generate_deletions left on the GPU is applied."""
        if pointers is None and u is None:
            pointers = self.deletions_device()
        if pointers is not None:
            self._check(self.lib.nlp_apply_deletions(self.h, pointers[0], pointers[1], pointers[2]))
            return
        u = np.ascontiguousarray(u, dtype=np.uint32); v = np.ascontiguousarray(v, dtype=np.uint32)
        self._check(self.lib.nlp_apply_deletions(self.h, u.ctypes.data if u.size else None, v.ctypes.data if v.size else None, u.size))

    def graph_checkpoint(self):
        """Mark the resident graph as the base of a batch loop (main.cxx:164)."""
        self._check(self.lib.nlp_graph_checkpoint(self.h))

    def graph_rollback(self):
        """Make the checkpointed base graph the resident graph again."""
        self._check(self.lib.nlp_graph_rollback(self.h))

    def graph_size(self):
        s, m = C.c_uint32(0), C.c_uint64(0)
        self._check(self.lib.nlp_graph_size(self.h, C.byref(s), C.byref(m)))
        return int(s.value), int(m.value)

    def fetch_graph(self):
        """Host copy of the resident CSR (numpy uint64 offsets, uint32 keys)."""
        S, M = self.graph_size()
        off = np.empty(S + 1, np.uint64); keys = np.empty(M, np.uint32)
        self._check(self.lib.nlp_fetch_graph(self.h, off.ctypes.data, keys.ctypes.data if M else None))
        return off, keys

    def ingest_mtx(self, text, symmetrize=True, drop_self_loops=True):
        """Matrix Market coordinate text (bytes) -> resident CSR, built on the GPU (main.cxx:243-245).
        Returns (span, entries)."""
        s, m = C.c_uint32(0), C.c_uint64(0)
        flags = (1 if symmetrize else 0) | (2 if drop_self_loops else 0)
        self._check(self.lib.nlp_ingest_mtx(self.h, text, len(text), flags, C.byref(s), C.byref(m)))
        self._keep = None
        return int(s.value), int(m.value)

    def launch_count(self):
        return int(self.lib.nlp_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.nlp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Graph ingest: nlp_ingest_mtx on the GPU against the reference's own reader + symmetrize +
self-loop removal (main.cxx:243-245, oracle/_ref) on the host cores, same Matrix Market text.

    python tools/ingest_bench.py [rmat scale] [edge factor]
"""
import ctypes as C
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import pandas as pd     # noqa: E402
import torch            # noqa: E402
import nlp_b200 as N    # noqa: E402
from oracle import oracle_py as O   # noqa: E402


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ef = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    off, keys = N.graphs.rmat(scale, ef, 47, permute=True, device="cuda")
    S = int(off.numel() - 1)
    rows = torch.repeat_interleave(torch.arange(S, device="cuda"), (off[1:] - off[:-1]))
    k = keys.to(torch.int64)
    up = rows < k                                        # one line per undirected edge ("general" file, main.cxx symmetrizes)
    u = rows[up].cpu().numpy(); v = k[up].cpu().numpy()
    rng = np.random.default_rng(1)
    flip = rng.random(u.size) < 0.5                      # either direction, and 5 % of the edges listed both ways
    a = np.where(flip, v, u); b = np.where(flip, u, v)
    both = rng.random(u.size) < 0.05
    a = np.concatenate([a, b[both]]); b = np.concatenate([b, a[:u.size][both]])
    t0 = time.time()
    buf = io.StringIO()
    pd.DataFrame({"u": a, "v": b, "w": 1}).to_csv(buf, sep=" ", header=False, index=False)
    text = ("%%%%MatrixMarket matrix coordinate integer general\n%d %d %d\n" % (S - 1, S - 1, a.size)).encode() + buf.getvalue().encode()
    del buf
    t_text = time.time() - t0
    p = N.Predictor(0)
    p.ingest_mtx(text)                                   # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    span, entries = p.ingest_mtx(text)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    goff, gkeys = p.fetch_graph()
    path = "/dev/shm/ingest_bench.mtx" if os.path.isdir("/dev/shm") else "/tmp/ingest_bench.mtx"
    with open(path, "wb") as f:
        f.write(text)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libnlpref_batch.so"))
    lib.nlpref_read_mtx.restype = C.c_int64
    lib.nlpref_read_mtx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    sp = C.c_uint32(0)
    roff = np.empty(span + 1, np.uint64); rkeys = np.empty(entries + 16, np.uint32)
    t0 = time.perf_counter()
    m = lib.nlpref_read_mtx(path.encode(), 0, 1, C.byref(sp), roff.ctypes.data, rkeys.ctypes.data)
    t_ref = time.perf_counter() - t0
    os.remove(path)
    same = bool(m == entries and np.array_equal(roff, goff) and np.array_equal(rkeys[:m], gkeys))
    dup = int((gkeys[1:] == gkeys[:-1]).sum())
    print(json.dumps({"workload": "R-MAT %d ef %d as a general Matrix Market file, 5%% of the edges listed in both directions" % (scale, ef),
                      "text_bytes": len(text), "lines": int(a.size), "span": span, "entries": entries,
                      "entries_stored_twice_by_the_reference_merge": dup,
                      "gpu_ingest_s": round(t_gpu, 4), "gpu_text_gbps": round(len(text) / t_gpu / 1e9, 2),
                      "reference_ingest_s": round(t_ref, 3), "reference_threads": os.cpu_count(),
                      "speedup": round(t_ref / t_gpu, 1), "identical_to_reference": same, "text_build_s": round(t_text, 1)}))


if __name__ == "__main__":
    main()

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libnlpref.so).

Run in the build container only (needs /root/reference to have been compiled by
``make -C oracle ref``):   python tests/golden/make_golden.py

For every fixture graph and every (measure, D) the reference's *sequential* entry point
(inc/predict.hxx:502-831, the non-Omp twins) is called with maxEdges = size_t(-1), which returns
every candidate; the list is put into the canonical (score desc, u asc, v asc) order.
Small graphs store the full lists; larger ones store the first 4096 rows, the row count and a
SHA-256 over the full canonical (u, v, score-bits) arrays.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nlp_b200 as N                   # noqa: E402
from oracle import oracle_py as O      # noqa: E402
import parity                          # noqa: E402

DEGREES = [0, 2, 4, 16]
FULL_LIMIT = 20000
TOP = 4096


def digest(u, v, s):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(u, np.uint32).tobytes())
    h.update(np.ascontiguousarray(v, np.uint32).tobytes())
    h.update(np.ascontiguousarray(s, np.float32).view(np.uint32).tobytes())
    return h.hexdigest()


def fixture_graphs():
    g = N.graphs
    return {
        "kat7": parity.kat_graph(),
        "rmat7": g.to_numpy(*g.rmat(7, 6, 11)),
        "road12": g.to_numpy(*g.road_lattice(12, 0.6, 12)),
        "pp300_multiset": g.to_numpy(*g.duplicate_some_entries(*g.planted_partition(300, 10, 6, 2, 13), every=5)),
        "rmat10": g.to_numpy(*g.rmat(10, 8, 1)),
    }


def main():
    assert O.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    for name, (off, keys) in fixture_graphs().items():
        R = O.RefGraph(off, keys)
        out = {"offsets": off, "keys": keys}
        for m in O.MEASURES:
            for D in DEGREES:
                u, v, s, _, _ = R.predict(m, D, omp=False)
                tag = "%s_%d" % (m, D)
                out[tag + "_count"] = np.array([len(u)], np.uint64)
                out[tag + "_sha"] = np.array([digest(u, v, s)])
                n = len(u) if len(u) <= FULL_LIMIT else TOP
                out[tag + "_u"] = u[:n]; out[tag + "_v"] = v[:n]; out[tag + "_s"] = s[:n].view(np.uint32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "span", len(off) - 1, "entries", len(keys),
              "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))
        R.close()


if __name__ == "__main__":
    main()

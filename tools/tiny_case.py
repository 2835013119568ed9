"""Smallest end-to-end case (for compute-sanitizer / debugging): python tools/tiny_case.py [measure] [D] [K] [path]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nlp_b200 as N   # noqa: E402

m = sys.argv[1] if len(sys.argv) > 1 else "JC"
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 500
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
off, keys = N.graphs.to_numpy(*N.graphs.rmat(10, 8, 7))
p = N.Predictor(0)
p.set_graph(off, keys)
p.set_path(path)
r = p.predict(m, D, max_edges=K)
print(r)
print(p.fetch(min(5, r["count"])))

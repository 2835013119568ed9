"""Run a matrix of (graph, measure, D, K) parity cases on the GPU and report every failure.

Usage (on a GPU box):  python tests/gpu_sweep.py [--quick]
Used during bring-up; the pytest -m gpu tests cover the same ground case by case.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nlp_b200 as N                     # noqa: E402
from oracle import oracle_py as O        # noqa: E402
import parity                            # noqa: E402


def graphs(quick):
    g = N.graphs
    out = [("kat", parity.kat_graph())]
    out.append(("rmat10", g.to_numpy(*g.rmat(10, 8, 1))))
    out.append(("road40", g.to_numpy(*g.road_lattice(40, 0.6, 2))))
    out.append(("pp2k", g.to_numpy(*g.planted_partition(2000, 40, 8, 2, 3))))
    out.append(("pp2k-dup", g.to_numpy(*g.duplicate_some_entries(*g.planted_partition(2000, 40, 8, 2, 3), every=5))))
    out.append(("pp2k-symdup", g.to_numpy(*g.duplicate_symmetric(*g.planted_partition(2000, 40, 8, 2, 3), every=5, copies=3))))
    out.append(("rmat13", g.to_numpy(*g.rmat(13, 16, 4))))
    if not quick:
        out.append(("web50k", g.to_numpy(*g.web_crawl(50000, 12, seed=5))))
        out.append(("rmat16", g.to_numpy(*g.rmat(16, 16, 6))))
    only = os.environ.get("SWEEP_GRAPHS")
    if only:
        out = [g for g in out if g[0] in only.split(",")]
    return out


def main():
    quick = "--quick" in sys.argv
    pred = N.Predictor(0)
    fails = 0
    total = 0
    t0 = time.time()
    for name, (off, keys) in graphs(quick):
        pred.set_graph(off, keys)
        M = len(keys)
        for measure in N.MEASURES:
            for D in (0, 2, 4, 16, 1024):
                for K in (N.UNBOUNDED, max(1, M // 20), 3):
                    if name == "rmat16" and D in (0, 1024) and (measure not in ("CN", "JC", "AA") or K == N.UNBOUNDED):
                        continue   # the oracle needs minutes for 1.8e8 unbounded candidates
                    for path in ((1, 2) if D else (1,)):   # source-centric kernels; LHub pair path
                        total += 1
                        pred.set_path(path)
                        try:
                            err, r, st = parity.check_case(pred, O, off, keys, measure, D, K, tag="%s path%d" % (name, path))
                        except Exception as e:   # noqa: BLE001
                            err = "%s %s D=%d path%d: EXCEPTION %r" % (name, measure, D, path, e)
                        if err:
                            fails += 1
                            print("FAIL", err, flush=True)
        print("graph %s done (%d entries) t=%.1fs fails=%d" % (name, M, time.time() - t0, fails), flush=True)
    print("SWEEP total=%d fails=%d" % (total, fails))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())

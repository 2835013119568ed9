"""Small driver for ncu: build one workload on the GPU and run a few predictions.

    python tools/profile_one.py [workload] [degree] [measures] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch            # noqa: E402
import bench            # noqa: E402
import nlp_b200 as N    # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "rmat20"
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    measures = sys.argv[3].split(",") if len(sys.argv) > 3 else ["JC", "AA"]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    path = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    pred = N.Predictor(0)
    off, keys, K, info, _, _ = bench.build_workload(wl, "cuda:0", pred=pred)
    pred.set_graph_pointers(off.data_ptr(), keys.data_ptr(), off.numel() - 1, device=True, keep=(off, keys))
    pred.set_path(path)
    for _ in range(reps):
        for m in measures:
            r = pred.predict(m, D, max_edges=K)
            print(m, D, {k: r[k] for k in ("count", "time_ms", "scoring_ms", "select_ms", "frontier_ms", "wedges",
                                           "candidates", "kept", "emitted", "passes", "path", "pair_records")},
                  "bins", r["bin_sources"][:6], "phase", [round(x, 3) for x in r["phase_ms"]], flush=True)


if __name__ == "__main__":
    main()

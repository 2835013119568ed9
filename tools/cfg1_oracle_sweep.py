"""BASELINE configs[0] in main.cxx's own order (nine measures x eleven thresholds = 99 predictions,
main.cxx:67-80, 212-220) on R-MAT 18 with 10^-2 |E| removed by the reference sampler
(default_random_engine(12345)): every prediction on the GPU through the C ABI AND on the C oracle
(canonical tie order) on the same CSR.  Prints one JSON object per prediction: full-list parity,
precision / recall from nlp_evaluate and from the host formula of main.cxx:48-57, 201-202, device
time.  This is the line-by-line explanation of where a reference run (ties kept by heap accident,
inc/predict.hxx:332) may differ from the GPU run: the canonical oracle never does.

    python tools/cfg1_oracle_sweep.py [workload] > profiles/r02_cfg1_sweep_gpu_vs_oracle.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import bench            # noqa: E402
import nlp_b200 as N    # noqa: E402
import parity           # noqa: E402
from oracle import oracle_py as O   # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "rmat18"
    degrees = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else O.REF_DEGREES
    pred = N.Predictor(0)
    off, keys, K, info, (du, dv), _ = bench.build_workload(wl, "cuda:0", pred=pred)
    S = int(off.numel() - 1)
    pred.set_graph_pointers(off.data_ptr(), keys.data_ptr(), S, device=True, keep=(off, keys))
    pred.set_truth_pointers(du.data_ptr(), dv.data_ptr(), int(du.numel()))
    offn, keysn = N.graphs.to_numpy(off, keys)
    dun = du.cpu().numpy().astype(np.int64); dvn = dv.cpu().numpy().astype(np.int64)
    truth = dun * S + dvn
    threads = os.cpu_count() or 1
    total_ms, bad = 0.0, 0
    print(json.dumps({"config": info, "threads_for_oracle": threads}), flush=True)
    for m in N.MEASURES:
        for D in degrees:
            r = pred.predict(m, D, max_edges=K)
            ev = pred.evaluate()
            u, v, s = pred.fetch(r["count"])
            t0 = time.time()
            wu, wv, ws, st = O.oracle_predict(offn, keysn, m, D, max_edges=K, threads=threads)
            osec = time.time() - t0
            err = parity.compare((u, v, s), (wu, wv, ws), "%s D=%d" % (m, D))
            a = np.unique(np.concatenate([wu, wv]).astype(np.int64) * S + np.concatenate([wv, wu]).astype(np.int64))
            common = np.intersect1d(a, truth, assume_unique=True).size
            op, orc = common / max(a.size, 1), common / max(truth.size, 1)
            same = err is None and ev["precision"] == op and ev["recall"] == orc
            bad += 0 if same else 1
            total_ms += r["time_ms"]
            print(json.dumps({"measure": m, "D": D, "edges": r["count"], "gpu_time_ms": r["time_ms"], "gpu_scoring_ms": r["scoring_ms"],
                              "path": r["path"], "precision": ev["precision"], "recall": ev["recall"],
                              "oracle_precision": op, "oracle_recall": orc, "list_parity": "bit-exact" if err is None else err,
                              "oracle_seconds": round(osec, 2)}), flush=True)
    print(json.dumps({"predictions": len(N.MEASURES) * len(degrees), "mismatches": bad, "gpu_prediction_ms_total": total_ms}), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

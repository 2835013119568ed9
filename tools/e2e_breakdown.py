"""Where the end-to-end step goes (bench.py's `e2e` leg): host timings, with a device
synchronize on both sides, of each C-ABI call of one step on the default workload.

    python tools/e2e_breakdown.py [workload] [D]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402

import bench   # noqa: E402
import nlp_b200 as N   # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "rmat22"
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    pred = N.Predictor(0)
    off, keys, K, info, _, _ = bench.build_workload(name, "cuda:0", pred=pred)
    S = int(off.numel() - 1)
    h_off = off.cpu().pin_memory(); h_keys = keys.cpu().pin_memory()
    out = [torch.empty(K, dtype=torch.int32).pin_memory() for _ in range(3)]

    def t(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r

    rows = []
    for rep in range(3):
        ms_set, _ = t(lambda: pred.set_graph_pointers(h_off.data_ptr(), h_keys.data_ptr(), S, device=False, keep=(h_off, h_keys)))
        ms_first, r = t(lambda: pred.predict("CN", D, max_edges=K))
        ms_second, r2 = t(lambda: pred.predict("JC", D, max_edges=K))
        ms_fetch, _ = t(lambda: pred.fetch_into(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), r2["count"]))
        per = {}
        for m in ("SI", "SC", "HP", "HD", "LHN", "AA", "RA"):
            ms_m, rm = t(lambda: pred.predict(m, D, max_edges=K))
            per[m] = [round(ms_m, 3), round(rm["time_ms"], 3), rm["path"]]
        rows.append({"set_graph_ms": ms_set, "first_predict_ms": ms_first, "first_predict_device_ms": r["time_ms"],
                     "second_predict_ms": ms_second, "second_predict_device_ms": r2["time_ms"], "fetch_ms": ms_fetch,
                     "rest_host_ms_device_ms_path": per, "first_path": r["path"],
                     "h2d_GBps": ((S + 1) * 8 + keys.numel() * 4) / ms_set / 1e6, "d2h_GBps": r2["count"] * 12 / ms_fetch / 1e6})
    print(json.dumps({"workload": name, "D": D, "K": K, "reps": rows}))


if __name__ == "__main__":
    main()

"""Where the end-to-end step goes (bench.py's `e2e` leg): host timings, with a device synchronize
on both sides, of each C-ABI call of the batch loop on the default workload, for several steps.

    python tools/e2e_breakdown.py [workload] [D] [steps]
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
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    pred = N.Predictor(0)
    off, keys, K, info, (du, dv), _ = bench.build_workload(name, "cuda:0", pred=pred)
    base = bench.BASE_GRAPH["graph"]
    S = int(off.numel() - 1)
    h_du = du.cpu().pin_memory(); h_dv = dv.cpu().pin_memory()
    out = [[torch.empty(K, dtype=torch.int32).pin_memory() for _ in range(2)] + [torch.empty(K, dtype=torch.float32).pin_memory()] for _ in range(2)]
    pred.set_graph_pointers(base[0].data_ptr(), base[1].data_ptr(), S, device=True, keep=base)
    pred.graph_checkpoint()

    def t(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        return round((time.perf_counter() - t0) * 1e3, 3), r

    rows = []
    for rep in range(steps):
        row = {}
        row["rollback_ms"], _ = t(pred.graph_rollback)
        row["apply_ms"], _ = t(lambda: pred.apply_deletions(pointers=(h_du.data_ptr(), h_dv.data_ptr(), int(h_du.numel()))))
        per = {}
        for i, m in enumerate(N.MEASURES):
            ms, r = t(lambda: pred.predict(m, D, max_edges=K))
            o = out[i % 2]
            fms, _ = t(lambda: pred.fetch_into(o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), r["count"]))
            per[m] = [ms, round(r["time_ms"], 3), fms]
        row["predict_host_ms__device_ms__fetch_ms"] = per
        row["sum_ms"] = round(row["rollback_ms"] + row["apply_ms"] + sum(v[0] + v[2] for v in per.values()), 2)
        rows.append(row)
    print(json.dumps({"workload": name, "D": D, "K": K, "deletions": int(h_du.numel()), "steps": rows}))


if __name__ == "__main__":
    main()

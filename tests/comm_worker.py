"""Worker of tests/test_comm.py: one process per GPU, NCCL communicator inside the C ABI.
argv: rank world uid_file out_file"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [("JC", 16, 20000), ("AA", 16, 20000), ("CN", 4, 5000), ("CN", 16, 10**9), ("JC", 0, 20000), ("AA", 0, 3000), ("RA", 1024, 20000)]


def graph():
    import nlp_b200 as N
    return N.graphs.to_numpy(*N.graphs.rmat(14, 16, 91, permute=True))


def main():
    rank, world, uid_file, out_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    import nlp_b200 as N
    p = N.Predictor(rank)
    if rank == 0:
        uid = N.Predictor.comm_unique_id()
        with open(uid_file + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uid_file + ".tmp", uid_file)
    else:
        t0 = time.time()
        while not os.path.exists(uid_file):
            time.sleep(0.05)
            if time.time() - t0 > 120:
                raise SystemExit("no uid file")
        uid = open(uid_file, "rb").read()
    p.comm_init(uid, rank, world)
    off, keys = graph()
    p.set_graph(off, keys)
    out = {}
    for path in (0, 3, 1):
        p.set_path(path)
        for i, (m, D, K) in enumerate(CASES):
            if D == 0 and path != 0:
                continue
            r = p.predict(m, D, max_edges=K)
            u, v, s = p.fetch(r["count"])
            out["u_%d_%d" % (path, i)] = u; out["v_%d_%d" % (path, i)] = v; out["s_%d_%d" % (path, i)] = s
            out["path_%d_%d" % (path, i)] = np.array([r["path"]])
    out["bytes"] = np.array([p.comm_bytes()])
    np.savez(out_file, **out)
    p.comm_destroy()
    p.close()


if __name__ == "__main__":
    main()

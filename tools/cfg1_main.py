"""BASELINE configs[0]: the reference's own experiment driver (main.cxx) on an R-MAT scale-18
.mtx file with 10^-2 |E| edges removed -- once on the reference's host-OpenMP predict.hxx
(oracle/_ref/ref_main) and once on the B200 path (oracle/_ref/dropin_main, the same main.cxx
compiled against include/predict_b200.hxx).  Both binaries use the same fixed RNG seed, so they
remove the same edges.

    python tools/cfg1_main.py ref    [scale] > profiles/...   # CPU only (run in the build container)
    python tools/cfg1_main.py dropin [scale]                  # needs a B200

Prints one JSON object per log line: technique, time_ms, scoring_ms, precision, recall.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LINE = re.compile(r"\{-(\S+)/\+(\S+) batchf, (\d+) threads\} -> \{(\S+)ms, (\S+)ms scoring, (\S+) precision, (\S+) recall\} (\S+)")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "dropin"
    scale = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    import nlp_b200 as N
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_main" if which == "ref" else "dropin_main")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "rmat%d.mtx" % scale)
        off, keys = N.graphs.rmat(scale, 16, 42)
        n = N.graphs.write_mtx(path, off, keys)
        sys.stderr.write("wrote %s: %d undirected edges\n" % (path, n))
        t0 = time.time()
        p = subprocess.run("ulimit -s unlimited; exec %s %s 1 0" % (exe, path), shell=True, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        wall = time.time() - t0
    rows = []
    for line in p.stdout.splitlines():
        m = LINE.search(line)
        if m:
            rows.append({"technique": m.group(8), "threads": int(m.group(3)), "time_ms": float(m.group(4)),
                         "scoring_ms": float(m.group(5)), "precision": float(m.group(6)), "recall": float(m.group(7))})
        else:
            sys.stderr.write(line + "\n")
    for r in rows:
        print(json.dumps(r))
    sys.stderr.write("%s: rc=%d, %d result lines, %.1f s wall\n" % (which, p.returncode, len(rows), wall))
    return 0 if p.returncode == 0 and rows else 1


if __name__ == "__main__":
    sys.exit(main())

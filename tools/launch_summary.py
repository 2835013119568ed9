"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and
share per kernel (torch's generator kernels are left out).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_<what>.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        if "nlp::" not in name and not name.startswith(("void k_", "k_", "void nlp", "nlp")):
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1e3 if unit in ("ns", "nsecond") else val * 1e3 if unit in ("ms", "msecond") else val
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("nlp::", "")
        rows.append((short, us))
    tot = defaultdict(float); cnt = defaultdict(int)
    for n, us in rows:
        tot[n] += us; cnt[n] += 1
    total = sum(tot.values())
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for n in sorted(tot, key=lambda k: -tot[k]):
        print("| `%s` | %d | %.0f | %.1f%% |" % (n, cnt[n], tot[n], 100 * tot[n] / total))
    print("| **all nlp kernels** | %d | %.0f | 100%% |" % (len(rows), total))


if __name__ == "__main__":
    main()

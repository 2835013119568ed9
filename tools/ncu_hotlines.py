"""Hot CUDA source lines of one kernel in an .ncu-rep (stall samples and executed instructions per
line; needs -lineinfo and --import-source on).  Read here with `ncu -i`, no GPU needed.

    python tools/ncu_hotlines.py report.ncu-rep kernel-regex [top] [launches to skip]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"      # launches of that kernel to skip in the report
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    acc, fname = [], None
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0].isdigit():
            try:
                acc.append((int(r[4]), int(r[7]), fname, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    ts = sum(a[0] for a in acc) or 1
    ti = sum(a[1] for a in acc) or 1
    print("| samples | instructions | line | source |")
    print("|---|---|---|---|")
    for a in sorted(acc, reverse=True)[:top]:
        print("| %.1f%% | %.1f%% | `%s:%d` | `%s` |" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2], a[3], a[4].replace("|", "\\|")))


if __name__ == "__main__":
    main()

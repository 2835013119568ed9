"""Markdown table of the metrics the round notes quote from an `ncu --page raw --csv` export (made on
the GPU box by tools/ncu_all_kernels.sh; the .ncu-rep files are too large to bring back).

    python tools/ncu_raw_table.py gpurun_out/r02_ncu_X_raw.csv [> profiles/...]
"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed", "smem atom % of peak"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]


def fmt(x):
    try:
        v = float(x.replace(",", ""))
    except ValueError:
        return x
    if v == 0:
        return "0"
    if abs(v) >= 1e6:
        return "%.3g" % v
    return ("%.2f" % v).rstrip("0").rstrip(".")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in KEYS if k in hdr]
    print("| # | kernel | " + " | ".join("%s%s" % (n, " (%s)" % u if u and u not in ("%", "inst") else "") for _, n, u in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for i, r in enumerate(data):
        name = r[kn].split("(")[0].replace("void ", "").replace("nlp::", "")
        print("| %d | `%s` | %s |" % (i + 1, name, " | ".join(fmt(r[c]) for c, _, _ in cols)))


if __name__ == "__main__":
    main()

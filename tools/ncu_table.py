"""Markdown table of the metrics the round notes quote, from one or more .ncu-rep files
(read here with `ncu -i`, no GPU needed).

    python tools/ncu_table.py a.ncu-rep [b.ncu-rep ...]
"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    cols, units, names = [], {}, []
    for p in sys.argv[1:]:
        c, u = load(p)
        for d in c:
            cols.append(d); names.append(d.get("Kernel Name", "")[:40])
        units.update(u)
    print("| metric | " + " | ".join(names) + " | unit |")
    print("|---|" + "---|" * (len(cols) + 1))
    for k in KEYS:
        print("| %s | %s | %s |" % (k, " | ".join(c.get(k, "") for c in cols), units.get(k, "")))


if __name__ == "__main__":
    main()

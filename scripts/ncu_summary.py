#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: per kernel, the median capture of the metrics the judge reads.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv
    python scripts/ncu_summary.py gpurun_out/X_raw.csv "title line" > profiles/X_ncu_summary.txt
"""
import csv
import re
import statistics
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_selected",
]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = {}
    for m in METRICS:
        for j, h in enumerate(hdr):
            if h == m or h.endswith("." + m) or h.endswith(m):
                cols[m] = j
                break
    per = defaultdict(list)
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        per[name].append(r)
    print("# " + title)
    for name, rs in per.items():
        print("\n== %s  (%d captures, median shown)" % (name, len(rs)))
        for m in METRICS:
            if m not in cols:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[cols[m]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                print("   %-72s %14.4f %s" % (m, statistics.median(vals), units[cols[m]]))


if __name__ == "__main__":
    main()

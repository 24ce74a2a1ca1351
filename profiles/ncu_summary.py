#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch): duration, DRAM bytes, issue-slot
utilisation, instructions, occupancy and the top stall reasons.  usage: ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")].split("(")[0])
    for w in WANT:
        if w in hdr:
            print(f"    {w} {units[hdr.index(w)]} {r[hdr.index(w)]}")
    st = [(float(r[i]), hdr[i]) for i in range(len(hdr))
          if "smsp__average_warps_issue_stalled" in hdr[i] and hdr[i].endswith("_per_issue_active.ratio")
          and r[i] not in ("", "n/a")]
    top = ", ".join("%s=%.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                    for v, n in sorted(st, reverse=True)[:6])
    print("    top stalls (warps per issue): " + top)

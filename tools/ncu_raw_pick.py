#!/usr/bin/env python3
"""Prints the handful of raw ncu metrics the roofline discussion needs from `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {units[i]:10s} {r[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.05:
            print(f"{h:75s} {r[i]}")

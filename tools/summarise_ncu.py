#!/usr/bin/env python3
"""Text summary of one kernel from an `ncu --set full` report: duration, pipe utilisation, occupancy, stall reasons,
DRAM/L2 traffic.  usage: summarise_ncu.py <file.ncu-rep> [kernel-substring] > profiles/rNN_ncu_full_<kernel>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "smsp__inst_executed_op_shared_atom.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    if want and want not in name:
        continue
    print(f"# ncu --set full --clock-control none: {name.split('(')[0]}   (source: {rep.split('/')[-1]})")
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print(f"{k:80s} {d[k]:>18s} {u.get(k, '')}")
    print("# stall reasons (warps stalled per issue-active cycle)")
    st = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d.get(k)]
    for v, k in sorted(st, reverse=True)[:9]:
        print(f"{k:80s} {v:18.2f}")
    print()

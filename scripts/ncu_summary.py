#!/usr/bin/env python
"""Key per-kernel metrics from an ncu report (details page).  python scripts/ncu_summary.py rep [kernel-substr]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
idi, ki, mi, ui, vi = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
want = ("Duration", "Elapsed Cycles", "SM Active Cycles", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread",
        "Achieved Occupancy", "Theoretical Occupancy", "Executed Instructions", "Avg. Active Threads Per Warp",
        "Waves Per SM", "Warp Cycles Per Issued Instruction", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
        "Grid Size", "Block Size", "DRAM Throughput", "Memory Throughput", "L2 Hit Rate", "Dynamic Shared Memory Per Block",
        "Local Load Instructions", "Local Store Instructions")
seen = set()
for r in rows[1:]:
    if len(r) > vi and r[mi] in want and sub in r[ki]:
        k = (r[idi], r[mi])
        if k in seen:
            continue
        seen.add(k)
        print(f"[{r[idi]}] {r[ki][:40]:40s} {r[mi]:38s} {r[vi]:>14s} {r[ui]}")

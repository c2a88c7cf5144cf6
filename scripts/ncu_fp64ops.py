#!/usr/bin/env python
"""profiles/fp64_ops.json from an ncu --metrics CSV (thread-level DADD/DMUL/DFMA counts, FP64-pipe warp instructions and
all warp instructions per launch, averaged over the captured launches of each kernel).

    python scripts/ncu_fp64ops.py gpurun_out/x_fp64ops.csv <workload> [source-note]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, workload = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else path
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
h = rows[hdr]
ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
names = {"smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul",
         "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma", "smsp__inst_executed_pipe_fp64.sum": "fp64_warp_inst",
         "smsp__inst_executed.sum": "warp_inst", "gpu__time_duration.sum": "ncu_ns"}
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] not in names:
        continue
    k = re.sub(r"^.*?(\w+_kernel).*$", r"\1", r[ki])
    per.setdefault(k, {}).setdefault(r[idi], {})[names[r[mi]]] = float(r[vi].replace(",", ""))
out_path = os.path.join(ROOT, "profiles", "fp64_ops.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {"_comment": "per-launch instruction counts from ncu (--metrics smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on.sum, "
                       "smsp__inst_executed_pipe_fp64.sum, smsp__inst_executed.sum) over bench.py's launches, averaged per kernel; "
                       "read by bench.py for roofline.fp64"}
w = out.setdefault(workload, {})
for k, launches in per.items():
    n = len(launches)
    w[k] = {m: sum(l.get(m, 0.0) for l in launches.values()) / n for m in ("dadd", "dmul", "dfma", "fp64_warp_inst", "warp_inst", "ncu_ns")}
    w[k]["launches_averaged"] = n
w["source"] = note
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(w, indent=1))

#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample table from an ncu report (cuda,sass source page).

    python scripts/ncu_lines.py gpurun_out/x.ncu-rep <kernel-regex> [min_pct]
"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = {}
hdr = None
first_kernel = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if first_kernel is None:
            first_kernel = r[1]
        elif r[1] != first_kernel:
            cur_file = None  # only the first matching kernel launch
        continue
    if r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(i, c[len("stall_"):]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or cur_file is None or len(r) <= ie:
        continue
    if r[2] == "-" and r[ie].isdigit():  # a CUDA source line (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        e, s = int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0
        a = agg.setdefault(key, [0, 0, r[1], {}])
        a[0] += e
        a[1] += s
        for i, nm in stall_cols:
            if i < len(r) and r[i].isdigit() and int(r[i]):
                a[3][nm] = a[3].get(nm, 0) + int(r[i])
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f"kernel {first_kernel}\ntotal warp-instructions {tot}  samples {tots}")
allst = {}
for (f, ln), (e, s, src, st) in sorted(agg.items()):
    for k, v in st.items():
        allst[k] = allst.get(k, 0) + v
    if 100 * e / tot >= minpct or 100 * s / tots >= minpct:
        top = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"{e:>11} {100*e/tot:5.1f}%i {100*s/tots:5.1f}%s  {f}:{ln}: {src.strip()[:90]}   [{top}]")
ts = sum(allst.values()) or 1
print("stall samples by reason: " + ", ".join(f"{k} {100*v/ts:.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))

#!/usr/bin/env python
"""Write the judged summary of an ncu report into profiles/: key raw metrics per kernel launch (CSV) and the
per-source-line instruction/stall table of each kernel.

    python scripts/ncu_export.py gpurun_out/x.ncu-rep profiles/r01_qp_diag_n8
"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]
idx = [(k, h.index(k)) for k in keep if k in h]
with open(out + "_ncu_raw.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow([k for k, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])
traffic = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    short = "admm_fwd_kernel" if "admm_fwd" in name else ("qcqp_bwd_kernel" if "qcqp_bwd" in name else ("qp_bwd_kernel" if "qp_bwd" in name else name))
    def val(k):
        v, u = float(r[h.index(k)]), units[h.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic[short] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
print(json.dumps(traffic))
kernels = sorted(set(r[h.index("Kernel Name")] for r in rows[2:]))
with open(out + "_ncu_lines.txt", "w") as fh:
    for k in ("admm_fwd", "qp_bwd", "qcqp_bwd"):
        if any(k in n for n in kernels):
            fh.write(subprocess.run([sys.executable, "scripts/ncu_lines.py", rep, k, "0.8"], capture_output=True, text=True).stdout)
            fh.write("\n")

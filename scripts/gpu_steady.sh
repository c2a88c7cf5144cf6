#!/bin/bash
tag=${1:-ss}
mkdir -p gpurun_out
M=smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 python scripts/steady_state.py > gpurun_out/${tag}_plain.txt 2>&1; cat gpurun_out/${tag}_plain.txt
timeout 900 ncu --replay-mode app-range --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/${tag}_range.csv \
    python scripts/steady_state.py > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log; head -c 3000 gpurun_out/${tag}_range.csv
for wl in qcqp_n24 qcqp_n16; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd|_bwd' -s 4 -c 2 -o gpurun_out/${tag}_${wl}_prof -f \
    python bench.py --workload $wl --batch 65536 --steps 4 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_${wl}_ncu.log 2>&1
done
ls -la gpurun_out | grep ${tag}

#!/bin/bash
# Round-2 GPU visit: all parity tests with recorded statistics, smoke, both bench arms, ncu launch list + FP64 op counts +
# full capture of the hot kernels.   bash scripts/gpu_r2b.sh <tag>
tag=${1:-r02b}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_parity.txt
DQ_PARITY_LOG=gpurun_out/${tag}_parity.txt timeout 1800 python -m pytest tests -q -m gpu -rA 2>&1 | tail -110 > gpurun_out/${tag}_pytest_gpu.txt
tail -4 gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/${tag}_smoke.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/${tag}_bench_reference.json
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/${tag}_bench.json
cat gpurun_out/${tag}_bench.json | cut -c1-600
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -s 6 -c 16 --csv --log-file gpurun_out/${tag}_fp64ops.csv \
    python bench.py --steps 10 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_fp64ops.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 10 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd|_bwd' -s 6 -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 10 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_full.log 2>&1
# steady state: 40 forward launches on 4 streams as ONE profiled range (kernel replay would serialise them)
M2=smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 120 python scripts/steady_state.py > gpurun_out/${tag}_steady_plain.txt 2>&1
timeout 900 ncu --replay-mode app-range --clock-control none --metrics $M2 --csv --log-file gpurun_out/${tag}_steady_range.csv \
    python scripts/steady_state.py > gpurun_out/${tag}_steady_ncu.log 2>&1
tail -2 gpurun_out/${tag}_steady_ncu.log
ls -la gpurun_out | grep ${tag}

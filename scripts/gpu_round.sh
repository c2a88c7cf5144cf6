#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch list + full capture of the hot kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [workload]
tag=${1:-r01}
wlname=${2:-qp_diag_n8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_env.txt; nproc >> gpurun_out/${tag}_env.txt
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/${tag}_smoke.txt
timeout 600 python bench.py --impl reference --workload $wlname --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${tag}_bench_reference.json
timeout 600 python bench.py --workload $wlname 2>&1 | tail -1 | tee gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --workload $wlname --steps 10 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd|_bwd' -s 6 -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --workload $wlname --steps 10 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12

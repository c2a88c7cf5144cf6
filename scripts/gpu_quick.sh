#!/bin/bash
# Quick GPU visit: parity tests (no -x, summary only) + headline bench + ncu full capture of the hot kernels.
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -2 > gpurun_out/${tag}_bench.json
if [ "$2" != "noprof" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd|qp_bwd' -s 6 -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_ncu_full.log 2>&1
fi
tail -5 gpurun_out/${tag}_pytest_gpu.txt; cat gpurun_out/${tag}_bench.json

#!/bin/bash
tag=${1:-v2p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "bit_identical" 2>&1 | grep -E "^E  .*Assert|passed|failed" | cut -c1-400 | tee gpurun_out/${tag}_bit.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd_diag8' -s 12 -c 1 -o gpurun_out/${tag}_prof -f \
    python scripts/fwd_ab.py qp_diag > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log

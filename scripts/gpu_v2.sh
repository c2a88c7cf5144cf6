#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "bit_identical" 2>&1 | grep -E "^E  .*Assert|passed|failed" | cut -c1-400 | tee gpurun_out/${tag}_bit.txt
timeout 300 python scripts/fwd_ab.py qp_diag qp_dense 2>&1 | tail -12 | tee gpurun_out/${tag}_ab.txt
for v in scripts/variants/lib_*.so; do
  echo "== $v" | tee -a gpurun_out/${tag}_ab.txt
  DQ_LIB_PATH=$v timeout 300 python scripts/fwd_ab.py qp_diag qp_dense 2>&1 | grep persistent | tee -a gpurun_out/${tag}_ab.txt
done

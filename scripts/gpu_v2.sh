#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
for v in "" scripts/variants/lib_wps28.so; do
  echo "== [$v]" | tee -a gpurun_out/${tag}_ab.txt
  DQ_LIB_PATH=$v timeout 300 python scripts/fwd_ab.py qp_diag qp_dense 2>&1 | grep persistent | tee -a gpurun_out/${tag}_ab.txt
done

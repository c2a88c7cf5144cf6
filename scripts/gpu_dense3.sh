#!/bin/bash
# dense workloads + headline on the in-tree build, then the GPU test suite.  gpu_dense3.sh <tag>
tag=${1:-dn}
mkdir -p gpurun_out
bash scripts/gpu_dense2.sh ${tag} "qcqp_n16:65536 qcqp_n24:0 qcqp_n32:0 qp_dense_n32:0 qcqp_n8:0 qp_dense_n8:0 qp_diag_n8:0 qcqp_diag_n8:0" | awk 'NR<=8'
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt

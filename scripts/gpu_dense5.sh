#!/bin/bash
tag=${1:-dn}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
bash scripts/gpu_dense2.sh ${tag} "qcqp_n16:65536 qcqp_n16:262144 qcqp_n8:0 qp_dense_n8:0"

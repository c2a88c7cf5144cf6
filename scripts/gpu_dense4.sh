#!/bin/bash
tag=${1:-dn}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
bash scripts/gpu_dense2.sh ${tag} "qcqp_diag_n8:0 qcqp_n8:0 qcqp_n16:65536 qcqp_n24:0"
for lib in "" scripts/variants/lib_fastprox0.so; do DQ_LIB_PATH=$lib timeout 300 python bench.py --workload qcqp_diag_n8 --steps 1000 --warmup 10 --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$lib] long run', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline'].get('kernel_ms_sustained'), d['detail'])"; done

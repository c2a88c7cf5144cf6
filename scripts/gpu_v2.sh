#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "qcqp or cfg" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/${tag}_pytest_gpu.txt
for v in "" scripts/variants/lib_minb5.so scripts/variants/lib_minb6.so; do
DQ_LIB_PATH=$v python bench.py --workload qcqp_n24 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', d['config']['name'], d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a gpurun_out/${tag}_bench.txt
done

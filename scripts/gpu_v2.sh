#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
for v in "" scripts/variants/lib_f16w24.so scripts/variants/lib_f16w32.so; do
DQ_LIB_PATH=$v python bench.py --workload qcqp_n16 --batch 65536 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', d['config']['name'], d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a gpurun_out/${tag}_bench.txt
done

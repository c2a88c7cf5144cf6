#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
for v in "" scripts/variants/lib_wps20.so scripts/variants/lib_wps24.so scripts/variants/lib_fw2.so; do
for w in qcqp_n24; do
DQ_LIB_PATH=$v python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', d['config']['name'], d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a gpurun_out/${tag}_bench.txt
done; done

#!/bin/bash
tag=${1:-q8b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
for wl in qcqp_diag_n8 qcqp_n8; do
  timeout 600 python bench.py --workload $wl --steps 300 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${tag}_bench_$wl.json
  python -c "
import json,sys
l=json.loads(open('gpurun_out/${tag}_bench_$wl.json').read()); print('$wl', 'value', l['value'], 'ms', l['ms_per_step'], l['detail'], l['roofline']['kernel_ms'], l['roofline']['kernel_ms_sustained'], 'e2e', l['e2e']['ms_per_step'])"
done

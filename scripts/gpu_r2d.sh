#!/bin/bash
# dense-hint check + the other workloads' bench lines + the default line with other_configs
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/${tag}_pytest.txt; cat gpurun_out/${tag}_pytest.txt
for wl in qp_dense_n8 qcqp_n8 qcqp_diag_n8; do
  timeout 600 python bench.py --workload $wl --steps 300 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${tag}_bench_$wl.json
  python -c "
import json,sys
l=json.loads(open('gpurun_out/${tag}_bench_$wl.json').read()); print('$wl', 'value', l['value'], 'ms', l['ms_per_step'], l['detail'], l['roofline']['kernel_ms'], 'e2e', l['e2e']['ms_per_step'])"
done
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/${tag}_bench.json
python -c "
import json
l=json.loads(open('gpurun_out/${tag}_bench.json').read()); print('default value', l['value'], 'ms', l['ms_per_step']); print(json.dumps(l['other_configs'])[:3000])"

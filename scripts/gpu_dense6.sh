#!/bin/bash
tag=${1:-dn}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.txt
for lib in "" $(ls scripts/variants/lib_*.so 2>/dev/null); do
  for wl in qcqp_n16:65536 qcqp_n24:0 qp_dense_n8:0; do
    w=${wl%%:*}; b=${wl##*:}
    DQ_LIB_PATH=$lib timeout 300 python bench.py --workload $w --batch $b --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | \
      python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$lib]', d['config']['name'], 'ms/step', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])" | tee -a gpurun_out/${tag}_ab.txt
  done
done

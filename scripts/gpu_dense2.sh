#!/bin/bash
# A/B of library variants on chosen workloads: gpu_dense2.sh <tag> "<wl:batch> ..."
tag=${1:-dn}; wls=${2:-"qcqp_n16:65536"}
mkdir -p gpurun_out
out=gpurun_out/${tag}_ab.txt; : > $out
for rep in 1 2; do
for lib in "" $(ls scripts/variants/lib_*.so 2>/dev/null); do
  for wl in $wls; do
    w=${wl%%:*}; b=${wl##*:}
    DQ_LIB_PATH=$lib timeout 600 python bench.py --workload $w --batch $b --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs 2>/dev/null | tail -1 | \
      python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$lib]', d['config']['name'], 'B', d['config']['B_per_gpu'], 'ms/step', round(d['ms_per_step'],4), d['roofline']['kernel_ms'])" | tee -a $out
  done
done
done

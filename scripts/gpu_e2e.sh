#!/bin/bash
tag=${1:-e2e}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "pipelined or host_entry or autograd or headline" 2>&1 | tail -4 > $out
for z in 0 1 2; do
  echo "== DQ_HOST_ZEROCOPY=$z" >> $out
  DQ_HOST_ZEROCOPY=$z timeout 600 python bench.py --steps 300 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('e2e', l['e2e']['ms_per_step'], 'autograd', l['e2e_autograd'], 'value', l['value'], 'fp64', l['roofline']['fp64'])" >> $out
done
for wl in qcqp_n8 qp_dense_n8; do
for z in 0 1 2; do
  echo "== $wl DQ_HOST_ZEROCOPY=$z" >> $out
  DQ_HOST_ZEROCOPY=$z timeout 600 python bench.py --workload $wl --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('e2e', l['e2e']['ms_per_step'], 'autograd', l['e2e_autograd'], 'value', l['value'])" >> $out
done
done
cat $out

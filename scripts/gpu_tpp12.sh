#!/bin/bash
tag=${1:-t12}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "bit_identical or handoff or headline" 2>&1 | tail -5 > $out
cat $out
for v in trace; do
  echo "== $v" >> $out
  tw=1; [ $v = notile ] && tw=0
  TPP_TILE_WARPS=$tw DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 120 python scripts/tpp_trace.py 48 4 1 1000 8 2>&1 | grep -E "launch|warps|thread loop done|end  |d thread loop|trips per warp|cycles per trip|d tile|parked" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_ab.py --paths 3 --elems 8 --caps 32,48,64 2>&1 | grep "path 3" >> $out
done
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --steps 500 > gpurun_out/${tag}_bench.txt 2>&1
tail -1 gpurun_out/${tag}_bench.txt | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('value', l['value'], 'ms', l['ms_per_step'], l['detail'], l['roofline']['kernel_ms'], l['roofline']['kernel_ms_sustained'])" >> $out
cat $out

#!/bin/bash
tag=${1:-t6}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline" 2>&1 | tail -12 > $out
for e in 8 4 2; do
  echo "== trace E=$e" >> $out
  DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 $e 2>&1 | grep -E "launch|warps|d P read|d setup|d sort|d thread loop|trips per warp|d tile|d wait|parked" >> $out
done
timeout 600 python scripts/tpp_ab.py --paths 2,3 --elems 8,4,2 --caps 24,32,48,64 2>&1 >> $out
cat $out

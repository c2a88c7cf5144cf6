#!/bin/bash
tag=${1:-t10}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline" 2>&1 | tail -5 > $out
for v in trace noslow r4; do
  echo "== $v E=8" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 8 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip|update section" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_ab.py --paths 3 --elems 8 --caps 48 2>&1 | grep "path 3" >> $out
done
cat $out

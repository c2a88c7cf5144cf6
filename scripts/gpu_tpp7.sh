#!/bin/bash
tag=${1:-t7}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
: > $out
for v in trace nocoop; do
for e in 8 4; do
  echo "== $v E=$e" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 $e 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip" >> $out
done
done
echo "== trace E=8 no adaptive, max_iter 48" >> $out
DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py 0 4 0 48 8 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip" >> $out
cat $out

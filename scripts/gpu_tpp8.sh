#!/bin/bash
tag=${1:-t8}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
: > $out
for e in 8 4; do
  echo "== trace E=$e" >> $out
  DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 $e 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip|update section" >> $out
done
cat $out

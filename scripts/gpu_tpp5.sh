#!/bin/bash
tag=${1:-t5}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
: > $out
for v in trace coop; do
  for args in "0 4 0 48" "0 4 1 48" "48 4 1 1000"; do
  echo "== variant $v args(cap W adaptive max_iter) $args" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_trace.py $args 2>&1 | grep -E "launch|d thread loop|trips per warp" >> $out
  done
done
cat $out

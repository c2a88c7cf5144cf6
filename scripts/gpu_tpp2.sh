#!/bin/bash
tag=${1:-t2}
mkdir -p gpurun_out
for v in trace w2 c4; do
  W=4; [ $v = w2 ] && W=2
  for cap in 48 0; do
    echo "== variant $v cap $cap" >> gpurun_out/${tag}_trace.txt
    DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_trace.py $cap $W >> gpurun_out/${tag}_trace.txt 2>&1
  done
  echo "== variant $v A/B" >> gpurun_out/${tag}_trace.txt
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_ab.py --paths 3 --caps 32,48,64 >> gpurun_out/${tag}_trace.txt 2>&1
done
cat gpurun_out/${tag}_trace.txt

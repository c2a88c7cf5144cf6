#!/bin/bash
tag=${1:-t9}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline" 2>&1 | tail -12 > $out
for e in 8 4; do
  echo "== trace E=$e" >> $out
  DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 $e 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip|d tile|d setup|d P read" >> $out
done
timeout 600 python scripts/tpp_ab.py --paths 3 --elems 8,4 --caps 32,48,64 2>&1 >> $out
cat $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp8 -s 3 -c 1 -o gpurun_out/${tag}_prof -f \
    python scripts/tpp_ab.py --caps 48 --paths 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

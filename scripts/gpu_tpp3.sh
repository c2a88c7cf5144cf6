#!/bin/bash
tag=${1:-t3}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline or fast_sqrt" 2>&1 | tail -15 > $out
for cap in 48 32; do
  echo "== trace cap $cap" >> $out
  DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py $cap 4 >> $out 2>&1
done
timeout 600 python scripts/tpp_ab.py --caps 24,32,40,48,64 >> $out 2>&1
cat $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp8 -s 3 -c 1 -o gpurun_out/${tag}_prof -f \
    python scripts/tpp_ab.py --caps 48 --paths 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

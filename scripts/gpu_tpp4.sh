#!/bin/bash
tag=${1:-t4}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline or fast_sqrt" 2>&1 | tail -12 > $out
for v in trace coop r1 r4; do
  echo "== variant $v" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_trace.py 48 4 2>&1 | grep -E "launch|d thread loop|trips per warp|d tile|d setup" >> $out
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_ab.py --paths 3 --caps 32,48 2>&1 | grep path >> $out
done
cat $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp8 -s 3 -c 1 -o gpurun_out/${tag}_prof -f \
    python scripts/tpp_ab.py --caps 48 --paths 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log

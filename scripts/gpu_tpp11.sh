#!/bin/bash
tag=${1:-t11}
mkdir -p gpurun_out
out=gpurun_out/${tag}_out.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -s -k "bit_identical or handoff or headline" 2>&1 | tail -5 > $out
DQ_LIB_PATH=scripts/variants/lib_trace.so timeout 300 python scripts/tpp_trace.py 48 4 1 1000 8 2>&1 | grep -E "launch|d thread loop|trips per warp|cycles per trip|update section" >> $out
timeout 600 python scripts/tpp_ab.py --paths 2,3 --elems 8 --caps 40,48,56 --batches 16384,32768,65536,131072,262144 2>&1 >> $out
timeout 600 python scripts/tpp_ab.py --paths 3 --elems 4 --caps 48 --batches 65536 2>&1 | grep "path 3" >> $out
cat $out

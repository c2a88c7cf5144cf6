#!/bin/bash
tag=${1:-san2}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_summary.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "$tool: $(grep -c '^ok' gpurun_out/${tag}_$tool.log) paths ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_$tool.log | tail -1)" | tee -a gpurun_out/${tag}_summary.txt
done
for v in p5 p6 p9; do
  echo "== PPS4 variant $v" | tee -a gpurun_out/${tag}_pps.txt
  DQ_LIB_PATH=scripts/variants/lib_$v.so timeout 300 python scripts/tpp_ab.py --paths 3 --elems 8 --caps 32,48 2>&1 | grep "path 3" | tee -a gpurun_out/${tag}_pps.txt
done
echo "== default (7)" | tee -a gpurun_out/${tag}_pps.txt
timeout 300 python scripts/tpp_ab.py --paths 3 --elems 8 --caps 32,48 2>&1 | grep "path 3" | tee -a gpurun_out/${tag}_pps.txt
for wl in qp_dense_n8; do
  timeout 600 python bench.py --workload $wl --steps 300 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('$wl', 'value', l['value'], 'ms', l['ms_per_step'], l['detail'], l['roofline']['kernel_ms'])"
done

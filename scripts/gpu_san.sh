#!/bin/bash
tag=${1:-san}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "$tool: $(grep -c '^ok' gpurun_out/${tag}_$tool.log) paths ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_$tool.log | tail -1)" | tee -a gpurun_out/${tag}_summary.txt
done

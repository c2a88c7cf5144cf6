#!/bin/bash
# compute-sanitizer over every kernel path (scripts/sanitize.py): memcheck, racecheck, synccheck, initcheck
tag=${1:-san}
mkdir -p gpurun_out
rm -f gpurun_out/${tag}_summary.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "$tool: $(grep -c '^ok' gpurun_out/${tag}_$tool.log) paths ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_$tool.log | tail -1)" | tee -a gpurun_out/${tag}_summary.txt
done
grep -E "Warning: Race reported" gpurun_out/${tag}_racecheck.log | sed -E 's/\+0x[0-9a-f]+//g; s/<[^>]*>//g' | sort | uniq -c | sort -rn | head -20 >> gpurun_out/${tag}_summary.txt

#!/bin/bash
# Full GPU visit without profiling: all GPU tests, smoke, sanitizers on the kernel-path exercise, bench, A/B timing.
tag=${1:-full}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/${tag}_san_$tool.log 2>&1
  echo "$tool: $(grep -c '^ok' gpurun_out/${tag}_san_$tool.log) paths ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_san_$tool.log | tail -1)"
done
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/${tag}_bench.json | cut -c1-1500
timeout 300 python scripts/fwd_ab.py qp_diag qp_dense 2>&1 | tail -4 | tee gpurun_out/${tag}_ab.txt

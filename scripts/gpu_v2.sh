#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
for v in "" scripts/variants/lib_wps28.so; do
  DQ_LIB_PATH=$v python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lib [$v]  ms/step %.4f  %.4g solves/s single-stream %.4f kernel_ms %s' % (d['ms_per_step'], d['value'], d['config']['single_stream_ms_per_step'], d['roofline']['kernel_ms']))" | tee -a gpurun_out/${tag}_wps.txt
done

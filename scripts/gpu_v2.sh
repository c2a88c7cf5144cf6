#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "host_entry or autograd_surface" 2>&1 | grep -E "^E  .*Assert|passed|failed" | cut -c1-400 | tee gpurun_out/${tag}_host.txt
for taper in 0 1; do for chunks in 4 6 8; do
  DQ_HOST_TAPER=$taper DQ_HOST_CHUNKS=$chunks python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('taper $taper chunks $chunks  e2e ms/step %.3f  %.3g solves/s' % (d['e2e']['ms_per_step'], d['e2e']['value']))" | tee -a gpurun_out/${tag}_e2e_sweep.txt
done; done
DQ_HOST_TRACE=1 python scripts/micro/e2e_trace.py 2>&1 | tail -45 | tee gpurun_out/${tag}_trace.txt

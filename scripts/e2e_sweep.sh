#!/bin/bash
# Sweep the host-pipeline knobs of dq_*_solve_host and print the e2e figure of bench.py for each.
for slots in 3 6 8; do for chunks in 4 8 12 16 24; do
  DQ_HOST_SLOTS=$slots DQ_HOST_CHUNKS=$chunks python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slots $slots chunks $chunks  e2e ms/step %.3f  %.3g solves/s' % (d['e2e']['ms_per_step'], d['e2e']['value']))"
done; done

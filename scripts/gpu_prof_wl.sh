#!/bin/bash
# Profile one workload: bench line + ncu full capture of fwd and bwd.  usage: gpu_prof_wl.sh <tag> <workload> [batch]
tag=$1; w=$2; b=${3:-0}
mkdir -p gpurun_out
python bench.py --workload $w --batch $b --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | tee gpurun_out/${tag}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['name'], d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'admm_fwd|_bwd' -s 6 -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --workload $w --batch $b --steps 5 --warmup 3 --streams 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1

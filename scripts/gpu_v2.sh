#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 | tee gpurun_out/${tag}_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for flag in "" "--no-numa"; do
timeout 300 $TR --master-port 29531 bench.py --gpus 2 --steps 300 --warmup 10 $flag 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 $flag', d['value'], d['e2e'])" | tee -a gpurun_out/${tag}_numa.txt
done

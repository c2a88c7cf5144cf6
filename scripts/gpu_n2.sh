#!/bin/bash
# Two-GPU visit: the driver's torchrun launch of both bench arms at N=2 (the default run includes the cfg5 section).
tag=${1:-n2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 10 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
timeout 300 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_n2.json 2>> gpurun_out/${tag}_bench_n2.err
tail -c 3000 gpurun_out/${tag}_bench_n2.json; tail -c 600 gpurun_out/${tag}_bench_reference_n2.json; tail -5 gpurun_out/${tag}_bench_n2.err

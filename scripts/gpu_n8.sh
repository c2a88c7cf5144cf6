#!/bin/bash
# Eight-GPU visit: the driver's torchrun launch of the bench at N=8 and N=4 (the default run includes the cfg5 section), reference arm at N=8.
tag=${1:-n8}
mkdir -p gpurun_out
for n in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 600 $TR --master-port 2951$n bench.py --gpus $n --steps 500 --warmup 10 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -c 1600 gpurun_out/${tag}_bench_n$n.json; echo; tail -2 gpurun_out/${tag}_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_n8.json 2>> gpurun_out/${tag}_bench_n8.err
tail -c 300 gpurun_out/${tag}_bench_reference_n8.json

#!/bin/bash
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${tag}_bench.json
cat gpurun_out/${tag}_pytest_gpu.txt; cat gpurun_out/${tag}_bench.json

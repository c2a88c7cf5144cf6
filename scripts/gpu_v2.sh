#!/bin/bash
tag=${1:-v2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "warm_start" 2>&1 | grep -E "^E  .*|passed|failed" | cut -c1-300 | tee gpurun_out/${tag}_ws.txt
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/${tag}_pytest_gpu.txt

#!/bin/bash
tag=${1:-lg}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -q -x -s -k "large_n or edge_sizes" 2>&1 | tail -30 | tee gpurun_out/${tag}_pytest.txt

#!/bin/bash
tag=${1:-q8}
mkdir -p gpurun_out
timeout 600 python scripts/tpp_ab.py --gen qcqp_diag --paths 1,2,3 --elems 8,4 --caps 32,48 --batches 65536 2>&1 | tee gpurun_out/${tag}_qcqp_diag.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "bit_identical or large_n or handoff" 2>&1 | tail -3

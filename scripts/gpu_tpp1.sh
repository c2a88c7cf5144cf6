#!/bin/bash
# GPU visit: thread-per-problem forward -- bit-identity tests, A/B timing, ncu capture.
tag=${1:-t1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "bit_identical or handoff or headline" 2>&1 | tail -15 > gpurun_out/${tag}_pytest.txt
cat gpurun_out/${tag}_pytest.txt
timeout 600 python scripts/tpp_ab.py > gpurun_out/${tag}_ab.txt 2>&1
cat gpurun_out/${tag}_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpp8 -s 3 -c 1 -o gpurun_out/${tag}_prof -f \
    python scripts/tpp_ab.py --caps 48 --paths 3 > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log

#!/bin/bash
tag=${1:-fin}
mkdir -p gpurun_out
DQ_PARITY_LOG=gpurun_out/${tag}_parity.txt timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in qp_dense_n8 qcqp_n8 qcqp_n16; do
  steps=100; case $w in *n8) steps=300;; esac
  timeout 900 python bench.py --workload $w --steps $steps --warmup 5 2>gpurun_out/${tag}_bench_$w.err | tail -1 > gpurun_out/${tag}_bench_$w.json
  python -c "
import json
l=json.loads(open('gpurun_out/${tag}_bench_$w.json').read()); print('$w', 'B', l['config']['B_per_gpu'], 'ms', round(l['ms_per_step'],4), 'value %.3e' % l['value'], {k: round(v,4) for k,v in l['roofline']['kernel_ms'].items()}, 'e2e %.3e' % (l['e2e']['value'] if l.get('e2e') else 0), 'cpu %.3e' % (l['cpu_baseline']['value'] if l.get('cpu_baseline') else 0))"
done

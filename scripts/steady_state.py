#!/usr/bin/env python
"""Steady state of the headline forward: N launches back to back on 4 streams between cudaProfilerStart/Stop, for
    ncu --replay-mode app-range --profile-from-start off --metrics ... python scripts/steady_state.py [workload-gen] [B] [launches]
(kernel replay would serialise the launches; a range keeps them concurrent as in bench.py's timed region)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffqcqp_b200 import _lib, workloads as wl

gen = sys.argv[1] if len(sys.argv) > 1 else "qp_diag"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
N = 8
L = _lib.load()
dev = torch.device("cuda", 0)
sets = [[x.to(dev) for x in getattr(wl, gen)(B, N, seed=r)] for r in range(4)]
xs = [torch.empty(B, N, 1, dtype=torch.float64, device=dev) for _ in range(4)]
streams = [torch.cuda.Stream(dev) for _ in range(4)]


def run(k):
    for i in range(k):
        d = sets[i % 4]
        rc = L.dq_qp_forward(d[0].data_ptr(), d[1].data_ptr(), None, xs[i % 4].data_ptr(), None, B, N, 1e-7, 1e-7, 1000, 1,
                             streams[i % 4].cuda_stream)
        assert rc == 0


run(8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for st in streams:
    st.wait_event(e0)
run(n)
for st in streams:
    ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream(dev).wait_event(ev)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{n} launches of {gen} B={B} on 4 streams: {e0.elapsed_time(e1) * 1e3 / n:.1f} us per launch")

#!/usr/bin/env python
"""A/B timing of the forward kernels on the headline workload (isolated launches, CUDA events).
usage: python scripts/fwd_ab.py [workload ...]   (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffqcqp_b200 import _lib, workloads as wl

L = _lib.load()
dev = torch.device("cuda", 0)
names = sys.argv[1:] or ["qp_diag", "qcqp_diag", "qp_dense", "qcqp_dense"]
B, N = 65536, 8
for name in names:
    sets = []
    for r in range(4):
        t = getattr(wl, name)(B, N, seed=r)
        d = [x.to(dev) for x in t]
        sets.append(d)
    x = torch.empty(B, N, 1, dtype=torch.float64, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream

    def fwd(d, iters=None):
        if name.startswith("qp"):
            rc = L.dq_qp_forward(d[0].data_ptr(), d[1].data_ptr(), None, x.data_ptr(), iters, B, N, 1e-7, 1e-7, 1000, 1, sp)
        else:
            rc = L.dq_qcqp_forward(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), None, x.data_ptr(),
                                   iters, B, N, 1e-7, 1e-7, 1000, 1, sp)
        assert rc == 0, rc

    res = {}
    for path in (1, 0, 1, 0):
        L.dq_set_forward_path(path)
        for k in range(5):
            fwd(sets[k % 4])
        torch.cuda.synchronize()
        ts = []
        for k in range(40):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fwd(sets[k % 4]); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res.setdefault(path, []).append((ts[len(ts) // 2], ts[0], sum(ts) / len(ts)))
        # back-to-back on 4 streams: throughput without the per-launch tail
        streams = [torch.cuda.Stream(dev) for _ in range(4)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for k in range(200):
            sp = streams[k % 4].cuda_stream
            fwd(sets[k % 4])
        sp = torch.cuda.current_stream(dev).cuda_stream
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream(dev).wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        res[path].append(("4-stream us/launch", e0.elapsed_time(e1) * 1e3 / 200))
    L.dq_set_forward_path(0)
    print(name, "generic (median, min, mean us):", res[1], flush=True)
    print(name, "persistent              :", res[0], flush=True)

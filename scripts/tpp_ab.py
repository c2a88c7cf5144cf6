#!/usr/bin/env python
"""A/B timing of the N = 8 forward kernels on the headline workload (run on the GPU box).
   paths: 1 generic tile kernel, 2 persistent tile kernel, 3 thread-per-problem kernel (with its park threshold swept).
usage: python scripts/tpp_ab.py [--caps 0,32,48,64] [--batches 65536] [--gen qp_diag]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffqcqp_b200 import _lib, workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument("--caps", default="0,24,32,48,64,96")
ap.add_argument("--batches", default="65536")
ap.add_argument("--gen", default="qp_diag")
ap.add_argument("--paths", default="2,3")
ap.add_argument("--elems", default="8")
a = ap.parse_args()
L = _lib.load()
dev = torch.device("cuda", 0)
N = 8
for B in [int(b) for b in a.batches.split(",")]:
    sets = [[x.to(dev) for x in getattr(wl, a.gen)(B, N, seed=r)] for r in range(4)]
    x = torch.empty(B, N, 1, dtype=torch.float64, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream(dev).cuda_stream

    def fwd(d, iters=None):
        if a.gen.startswith("qp"):
            rc = L.dq_qp_forward(d[0].data_ptr(), d[1].data_ptr(), None, x.data_ptr(), iters, B, N, 1e-7, 1e-7, 1000, 1, sp)
        else:
            rc = L.dq_qcqp_forward(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), None, x.data_ptr(),
                                   iters, B, N, 1e-7, 1e-7, 1000, 1, sp)
        assert rc == 0, rc

    L.dq_set_forward_path(1)
    fwd(sets[0], it.data_ptr()); torch.cuda.synchronize()
    xref, itref = x.clone(), it.clone()
    itf = itref.double()
    print(f"B={B} {a.gen}: iterations mean {itf.mean():.2f} p50 {itf.median():.0f} p99 {itf.quantile(0.99):.0f} max {itf.max():.0f}; "
          f">32: {(itref > 32).float().mean():.4f} >48: {(itref > 48).float().mean():.4f} >64: {(itref > 64).float().mean():.4f}", flush=True)
    configs = []
    for path in [int(p) for p in a.paths.split(",")]:
        if path == 3:
            configs += [(3, int(c), int(e)) for e in a.elems.split(",") for c in a.caps.split(",")]
        else:
            configs.append((path, None, 8))
    for path, cap, elems in configs:
        L.dq_set_forward_path(path)
        L.dq_set_forward_tuning(2, elems)
        if cap is not None:
            L.dq_set_forward_tuning(0, cap)
        fwd(sets[0], it.data_ptr()); torch.cuda.synchronize()
        same = bool(torch.equal(x.view(torch.int64), xref.view(torch.int64))) and bool(torch.equal(it, itref))
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
        streams = [torch.cuda.Stream(dev) for _ in range(4)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for k in range(400):
            sp = streams[k % 4].cuda_stream
            fwd(sets[k % 4])
        sp = torch.cuda.current_stream(dev).cuda_stream
        for st in streams:
            ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream(dev).wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        print(f"  path {path} E {elems} cap {cap}: bit-identical to generic {same}; isolated us median {ts[len(ts)//2]:.1f} min {ts[0]:.1f}; "
              f"4-stream us/launch {e0.elapsed_time(e1) * 1e3 / 400:.1f}", flush=True)
    L.dq_set_forward_path(0)
    L.dq_set_forward_tuning(0, 48)
    L.dq_set_forward_tuning(2, 0)

#!/usr/bin/env python
"""How an isolated forward launch's time splits into bulk and straggler tail: the headline batch with max_iter capped."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffqcqp_b200 import _lib, workloads as wl
L = _lib.load()
dev = torch.device("cuda", 0)
B, N = 65536, 8
sets = [[x.to(dev) for x in wl.qp_diag(B, N, seed=r)] for r in range(4)]
x = torch.empty(B, N, 1, dtype=torch.float64, device=dev)
it = torch.empty(B, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
for path in (0, 1):
    L.dq_set_forward_path(path)
    for cap in (1, 8, 16, 32, 64, 128, 256, 1000):
        ts = []
        for k in range(24):
            d = sets[k % 4]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.dq_qp_forward(d[0].data_ptr(), d[1].data_ptr(), None, x.data_ptr(), it.data_ptr(), B, N, 1e-7, 1e-7, cap, 1, sp)
            e1.record(); torch.cuda.synchronize(); assert rc == 0
            if k >= 4: ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"path {path} max_iter {cap:5d}: {sum(ts)/len(ts):7.1f} us   (unfinished at the cap: {int((it >= cap).sum())})", flush=True)
L.dq_set_forward_path(0)

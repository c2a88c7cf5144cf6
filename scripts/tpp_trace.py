#!/usr/bin/env python
"""Phase timeline of one isolated launch of the thread-per-problem forward (needs a -DDQ_TPP_TRACE build: DQ_LIB_PATH).
usage: DQ_LIB_PATH=scripts/variants/lib_trace.so python scripts/tpp_trace.py [cap_it] [warps_per_cta]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffqcqp_b200 import _lib, workloads as wl

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 48
W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ADAPT = int(sys.argv[3]) if len(sys.argv) > 3 else 1
MAXIT = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
ELEMS = int(sys.argv[5]) if len(sys.argv) > 5 else 8
L = _lib.load()
dev = torch.device("cuda", 0)
B, N = 65536, 8
P, q, _ = [t.to(dev) for t in wl.qp_diag(B, N, seed=0)]
x = torch.empty(B, N, 1, dtype=torch.float64, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
L.dq_set_forward_path(3)
L.dq_set_forward_tuning(0, cap)
L.dq_set_forward_tuning(2, ELEMS)
buf = torch.zeros(8192 * 8 * 32, dtype=torch.int64, device=dev)
L.dq_debug_set_trace.argtypes = [ctypes.c_void_p]
for k in range(3):
    assert L.dq_qp_forward(P.data_ptr(), q.data_ptr(), None, x.data_ptr(), None, B, N, 1e-7, 1e-7, MAXIT, ADAPT, sp) == 0
torch.cuda.synchronize()
assert L.dq_debug_set_trace(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
assert L.dq_qp_forward(P.data_ptr(), q.data_ptr(), None, x.data_ptr(), None, B, N, 1e-7, 1e-7, MAXIT, ADAPT, sp) == 0
e1.record()
torch.cuda.synchronize()
print(f"cap_it {cap}: launch {e0.elapsed_time(e1) * 1e3:.1f} us (event to event)")
t = buf.cpu().numpy().reshape(-1, 32)
t = t[t[:, 0] > 0]
nw = t.shape[0]
t0 = t[:, 0].min()
rel = (t[:, :8] - t0) / 1e3  # us
names = ["start", "P read done", "after sync", "setup done", "sorted", "thread loop done", "after sync", "end"]
W = W + int(os.environ.get("TPP_TILE_WARPS", "0"))  # + the tile warps of each CTA
print(f"{nw} warps ({nw // W} CTAs)")
for k, nm in enumerate(names):
    c = rel[:, k]
    print(f"  {nm:18s} min {c.min():7.1f}  p50 {np.median(c):7.1f}  p90 {np.percentile(c, 90):7.1f}  max {c.max():7.1f} us")
d = np.diff(rel, axis=1)
dn = ["P read", "sync", "setup", "sort+sync", "thread loop", "wait for CTA", "tile phase"]
for k, nm in enumerate(dn):
    c = d[:, k]
    print(f"  d {nm:16s} p50 {np.median(c):7.2f}  mean {c.mean():7.2f}  p90 {np.percentile(c, 90):7.2f}  max {c.max():7.2f} us")
trips = t[:, 8]
print(f"  trips per warp: mean {trips.mean():.1f} p50 {np.median(trips):.0f} max {trips.max()}  total {trips.sum()}  "
      f"us/trip (p50 loop / p50 trips) {np.median(d[:, 4]) / max(1, np.median(trips)):.3f}")
ns = t[::W, 9]
print(f"  parked per CTA: mean {ns.mean():.1f} max {ns.max()}")
acc = t[:, 10:15].astype(np.float64)
nm = ["body", "decisions", "update", "fin/park", "refill"]
tot = acc.sum(1)
print("  cycles per trip (lane-0 clock64, mean over warps): " + ", ".join(f"{n} {(acc[:, k] / np.maximum(trips, 1)).mean():.0f}" for k, n in enumerate(nm))
      + f"; sum {(tot / np.maximum(trips, 1)).mean():.0f}")
uu = t[:, 16:19].astype(np.float64)
print("  update section split (cycles per trip): " + ", ".join(f"{n} {(uu[:, k] / np.maximum(trips, 1)).mean():.0f}"
      for k, n in enumerate(["scalar part + post + sync", "shared items", "sync + pickup"])))

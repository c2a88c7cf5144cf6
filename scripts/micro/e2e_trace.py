import sys, time, torch
sys.path.insert(0, ".")
from diffqcqp_b200 import _lib, workloads as wl
L = _lib.load()
B, N = 65536, 8
P, q, g = wl.qp_diag(B, N, seed=0)
h = [t.pin_memory() for t in (P, q, g)]
hx = torch.empty(B, N, 1, dtype=torch.float64).pin_memory(); hgP = torch.empty(B, N, N, dtype=torch.float64).pin_memory(); hgq = torch.empty(B, N, 1, dtype=torch.float64).pin_memory()
for i in range(3):
    if i == 2: print("---- traced call", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    rc = L.dq_qp_solve_host(h[0].data_ptr(), h[1].data_ptr(), hx.data_ptr(), h[2].data_ptr(), hgP.data_ptr(), hgq.data_ptr(), B, N, 1e-7, 1e-7, 1000, 0)
    print("call %d: %.3f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)

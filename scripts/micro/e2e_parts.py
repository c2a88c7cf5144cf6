"""Where the host-buffer path's time goes: forward only / no grad_P read-back / full."""
import sys, time, torch
sys.path.insert(0, ".")
from diffqcqp_b200 import _lib, workloads as wl
L = _lib.load()
B, N = 65536, 8
P, q, g = wl.qp_diag(B, N, seed=0)
h = [t.pin_memory() for t in (P, q, g)]
hx = torch.empty(B, N, 1, dtype=torch.float64).pin_memory(); hgP = torch.empty(B, N, N, dtype=torch.float64).pin_memory(); hgq = torch.empty(B, N, 1, dtype=torch.float64).pin_memory()
def run(gx, gP, gq, reps=100):
    def one():
        rc = L.dq_qp_solve_host(h[0].data_ptr(), h[1].data_ptr(), hx.data_ptr(), gx, gP, gq, B, N, 1e-7, 1e-7, 1000, 0); assert rc == 0
    for _ in range(5): one()
    t0 = time.perf_counter()
    for _ in range(reps): one()
    return (time.perf_counter() - t0) / reps * 1e3
print("forward only (P,q in; x out)        : %.3f ms" % run(None, None, None))
print("fwd+bwd, grad_q only (no grad_P out): %.3f ms" % run(h[2].data_ptr(), None, hgq.data_ptr()))
print("fwd+bwd, full                       : %.3f ms" % run(h[2].data_ptr(), hgP.data_ptr(), hgq.data_ptr()))

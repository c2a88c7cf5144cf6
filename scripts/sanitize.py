"""Small end-to-end exercise of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool racecheck python scripts/sanitize.py
"""
import sys
import torch
sys.path.insert(0, ".")
from diffqcqp_b200 import qcqp as dq, workloads as wl

def run(tag, f):
    f(); torch.cuda.synchronize(); print("ok", tag, flush=True)

for N, B in ((8, 37), (5, 9), (16, 10), (24, 5), (32, 3)):
    P, q, g = wl.qp_dense(B, N, seed=N)
    Pd, qd, gd = P.cuda(), q.cuda(), g.cuda()
    run(f"qp dense N={N}", lambda: dq.qp_backward(Pd, qd, dq.qp_forward(Pd, qd, 1e-7, 200), gd))
    lo, hi, v = -torch.rand_like(qd), torch.rand_like(qd), torch.randn_like(qd)
    run(f"box N={N}", lambda: dq.boxqp_backward(Pd, qd, lo, hi, dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200), gd))
    run(f"signed box N={N}", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200, v=v))
P, q, g = wl.qp_diag(70, 8, seed=1)
run("qp diag N=8", lambda: dq.qp_backward(P.cuda(), q.cuda(), dq.qp_forward(P.cuda(), q.cuda(), 1e-7, 500), g.cuda()))
for N, B, diag in ((8, 21, False), (8, 21, True), (6, 7, False), (16, 9, False), (24, 4, False), (32, 3, False), (32, 3, True)):
    P, q, l_n, mu, g = wl.qcqp_dense(B, N, seed=N, diag=diag)
    a = [t.cuda() for t in (P, q, l_n, mu)]
    run(f"qcqp N={N} diag={diag}", lambda: dq.qcqp_backward(*a, dq.qcqp_forward(*a, 1e-7, 200), g.cuda()))

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
# the persistent-CTA forward (N == 8): shared queue with refills (several problems per tile slot), every prox, a
# batch that mixes a dense problem in (generic group routine inside the persistent kernel)
from diffqcqp_b200 import _lib
L = _lib.load()
P, q, g = wl.qp_diag(300, 8, seed=2)
Pd, qd = P.cuda(), q.cuda()
lo, hi, v = -torch.rand_like(qd), torch.rand_like(qd), torch.randn_like(qd)
run("persistent qp diag N=8 B=300", lambda: dq.qp_forward(Pd, qd, 1e-7, 500))
run("persistent box diag N=8", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200))
run("persistent signed box diag N=8", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200, v=v))
Pm = Pd.clone(); Pm[17] = wl.qp_dense(1, 8, seed=3)[0][0].cuda()
run("persistent qp mixed N=8", lambda: dq.qp_forward(Pm, qd, 1e-7, 200))
Pq, qq, l_n, mu, _ = wl.qcqp_diag(150, 8, seed=4)
L.dq_set_forward_path(2)
run("persistent qcqp diag N=8", lambda: dq.qcqp_forward(Pq.cuda(), qq.cuda(), l_n.cuda(), mu.cuda(), 1e-7, 200))
L.dq_set_forward_path(0)
for N, B, diag in ((8, 21, False), (8, 21, True), (6, 7, False), (16, 9, False), (24, 4, False), (32, 3, False), (32, 3, True)):
    P, q, l_n, mu, g = wl.qcqp_dense(B, N, seed=N, diag=diag)
    a = [t.cuda() for t in (P, q, l_n, mu)]
    run(f"qcqp N={N} diag={diag}", lambda: dq.qcqp_backward(*a, dq.qcqp_forward(*a, 1e-7, 200), g.cuda()))

# the warm-start extension (generic kernel, diagonal and dense)
P, q, g = wl.qp_dense(33, 8, seed=5)
x0 = dq.qp_forward(P.cuda(), q.cuda(), 1e-7, 200)
run("warm start qp dense N=8", lambda: dq.qp_forward(P.cuda(), q.cuda(), 1e-7, 200, warm_start=x0))
P, q, g = wl.qp_diag(33, 8, seed=6)
x0 = dq.qp_forward(P.cuda(), q.cuda(), 1e-7, 200)
run("warm start qp diag N=8", lambda: dq.qp_forward(P.cuda(), q.cuda(), 1e-7, 200, warm_start=x0))

# forward -> backward hand-off (dq_qp_forward_ex / dq_qp_backward_ex) on both forward kernels, diagonal and mixed batches
for path in (0, 1):
    L.dq_set_forward_path(path)
    for PP, tag in ((Pd, "diag"), (Pm, "mixed")):
        st = torch.empty(300, 8, 1, dtype=torch.float64, device="cuda")
        run(f"hand-off {tag} path={path}", lambda: dq.qp_backward(PP, qd, dq.qp_forward(PP, qd, 1e-7, 300, state=st), torch.ones_like(qd), state=st))
L.dq_set_forward_path(0)

# the thread-per-problem forward (N == 8, forced: path 3): refill from the queue, shared rho updates, parked problems on
# tiles (park threshold 5 so that the dump / tile phase is exercised, 0 = never park), every prox, E = 8 and 4, a dense chunk
P, q, g = wl.qp_diag(700, 8, seed=7)
Pd, qd = P.cuda(), q.cuda()
lo, hi, v = -torch.rand_like(qd), torch.rand_like(qd), torch.randn_like(qd)
Pq, qq, l_n, mu, _ = wl.qcqp_diag(700, 8, seed=8)
Pm = Pd.clone(); Pm[333] = wl.qp_dense(1, 8, seed=3)[0][0].cuda()
L.dq_set_forward_path(3)
for elems in (8, 4):
    L.dq_set_forward_tuning(2, elems)
    for cap in (48, 5, 0):
        L.dq_set_forward_tuning(0, cap)
        run(f"tpp qp diag E={elems} cap={cap}", lambda: dq.qp_forward(Pd, qd, 1e-7, 500))
    L.dq_set_forward_tuning(0, 5)
    run(f"tpp box E={elems}", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200))
    run(f"tpp signed box E={elems}", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200, v=v))
    run(f"tpp qcqp diag E={elems}", lambda: dq.qcqp_forward(Pq.cuda(), qq.cuda(), l_n.cuda(), mu.cuda(), 1e-7, 200))
    run(f"tpp qp mixed E={elems}", lambda: dq.qp_forward(Pm, qd, 1e-7, 200))
    st = torch.empty(700, 8, 1, dtype=torch.float64, device="cuda")
    run(f"tpp hand-off E={elems}", lambda: dq.qp_backward(Pd, qd, dq.qp_forward(Pd, qd, 1e-7, 300, state=st), torch.ones_like(qd), state=st))
L.dq_set_forward_tuning(0, 48); L.dq_set_forward_tuning(2, 0); L.dq_set_forward_path(0)
# legacy Box entry point with gamma / dgamma outputs
import numpy as np
from diffqcqp_b200 import legacy
r = np.random.default_rng(0)
S = r.random((8, 8)); Pn = S @ S.T / 8 + 0.1 * np.eye(8)
xb = legacy.solveBoxQP(Pn, r.random(8) - 0.5, -0.2 * np.ones(8), 0.2 * np.ones(8), np.zeros(8), 1e-8)
run("legacy box derivatives", lambda: legacy.solveDerivativesBoxQP(Pn, r.random(8) - 0.5, -0.2 * np.ones(8), 0.2 * np.ones(8), xb, r.random(8)))
# 32 < N <= 128: the warp-per-problem path (csrc/large_n.cu)
for N, B in ((40, 9), (64, 5)):
    P, q, g = wl.qp_dense(B, N, seed=N)
    Pd, qd, gd = P.cuda(), q.cuda(), g.cuda()
    run(f"large-N qp N={N}", lambda: dq.qp_backward(Pd, qd, dq.qp_forward(Pd, qd, 1e-7, 200), gd))
    lo, hi, v = -torch.rand_like(qd), torch.rand_like(qd), torch.randn_like(qd)
    run(f"large-N signed box N={N}", lambda: dq.boxqp_forward(Pd, qd, lo, hi, 1e-7, 200, v=v))
    P, q, l_n, mu, g = wl.qcqp_dense(B, N, seed=N)
    a = [t.cuda() for t in (P, q, l_n, mu)]
    run(f"large-N qcqp N={N}", lambda: dq.qcqp_backward(*a, dq.qcqp_forward(*a, 1e-7, 200), g.cuda()))

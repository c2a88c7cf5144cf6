import sys, torch, numpy as np
sys.path.insert(0, ".")
from diffqcqp_b200 import qcqp as dq, workloads as wl
from oracle import pyoracle as orc
for (B, N, diag) in ((16, 8, True), (16, 8, False), (64, 16, False), (33, 24, False)):
    P, q, ln, mu, g = wl.qcqp_dense(B, N, seed=3, diag=diag)
    x, it = dq.qcqp_forward(P.cuda(), q.cuda(), ln.cuda(), mu.cuda(), 1e-7, 1000, return_iters=True)
    torch.cuda.synchronize()
    xo, ito = orc.qcqp_forward(P.numpy(), q.numpy(), ln.numpy(), mu.numpy(), None, 1e-7, 1000, return_iters=True)
    print(B, N, diag, "max|dx|", np.abs(x.cpu().numpy() - xo).max(), "it mismatch", int((it.cpu().numpy() != ito).sum()), flush=True)

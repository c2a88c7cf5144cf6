import sys, torch, numpy as np
sys.path.insert(0, ".")
from diffqcqp_b200 import qcqp as dq, workloads as wl
from oracle import pyoracle as orc
def rel_rows(a, b):
    B = a.shape[0]
    return np.abs(a - b).reshape(B, -1).max(1) / (np.abs(b).reshape(B, -1).max(1) + 1e-300)
for (B, N, seed, diag) in ((2048, 8, 30, False), (1024, 16, 31, False), (512, 24, 32, False), (300, 32, 33, False), (512, 16, 34, True), (100, 6, 35, False)):
    P, q, ln, mu, g = wl.qcqp_dense(B, N, seed=seed, diag=diag)
    a = [t.numpy() for t in (P, q, ln, mu)]
    xo = orc.qcqp_forward(*a, None, 1e-7, 1000)
    gg = dq.qcqp_backward(P.cuda(), q.cuda(), ln.cuda(), mu.cuda(), torch.from_numpy(xo).cuda(), g.cuda())
    gg = [t.cpu().numpy() for t in gg]
    cands = []
    for f in (0, 1, 2, 3, 4, 5):
        orc.set_ir_force(f)
        cands.append(orc.qcqp_backward(*a, xo, g.numpy()))
    orc.set_ir_force(0)
    print(f"B={B} N={N} diag={diag}")
    for k, name in enumerate(("grad_P", "grad_q", "grad_l_n", "grad_mu")):
        r0 = rel_rows(gg[k], cands[0][k])
        rbest = np.min(np.stack([rel_rows(gg[k], c[k]) for c in cands[1:]]), 0)
        which = np.argmin(np.stack([rel_rows(gg[k], c[k]) for c in cands[1:]]), 0) + 1
        print(f"  {name:9s} vs default: med {np.median(r0):.1e} p90 {np.percentile(r0,90):.1e} p99 {np.percentile(r0,99):.1e} max {r0.max():.1e} | vs best forced: med {np.median(rbest):.1e} p90 {np.percentile(rbest,90):.1e} p99 {np.percentile(rbest,99):.1e} max {rbest.max():.1e}  which {np.bincount(which, minlength=6)[1:]}")

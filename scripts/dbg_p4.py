import sys, torch, numpy as np
sys.path.insert(0, ".")
from diffqcqp_b200 import qcqp as dq
from oracle import pyoracle as orc
g = torch.Generator().manual_seed(123)
B, N, eps = 4096, 8, 1e-10
p = torch.exp(20 * torch.rand(B, N, generator=g, dtype=torch.float64) - 10)
q = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
l_n = torch.rand(B, N // 2, 1, generator=g, dtype=torch.float64); mu = torch.rand(B, N // 2, 1, generator=g, dtype=torch.float64)
P4 = torch.diag_embed(p)
xo, ito = orc.qp_forward(P4.numpy(), q.numpy(), None, eps, 100000, return_iters=True)
x, it = dq.qp_forward(P4.cuda(), q.cuda(), eps, 100000, return_iters=True)
x = x.cpu().numpy(); it = it.cpu().numpy()
bad = np.nonzero(~np.isfinite(x).all(axis=(1, 2)))[0]
print("non-finite problems:", len(bad), bad[:10], "oracle finite:", np.isfinite(xo).all())
for i in bad[:3]:
    print("prob", i, "p^4", (p[i] ** 4).numpy(), "q", q[i, :, 0].numpy(), "\n  x_gpu", x[i, :, 0], "\n  x_orc", xo[i, :, 0], "it gpu", it[i], "it orc", ito[i])
d = np.abs(x - xo).reshape(B, -1).max(1); sc = np.maximum(1, np.abs(xo).reshape(B, -1).max(1))
ok = np.isfinite(d)
print("it mismatches", (it != ito).sum(), "max rel err (finite)", (d[ok] / sc[ok]).max(), "n rel>1e-9", ((d[ok] / sc[ok]) > 1e-9).sum(), "iters mean", ito.mean(), "max", ito.max())
w = np.argsort(-(d / sc)[ok])[:3]
for i in np.nonzero(ok)[0][w]:
    print("worst", i, "rel", d[i] / sc[i], "it", it[i], ito[i], "x_gpu", x[i, :, 0], "x_orc", xo[i, :, 0])

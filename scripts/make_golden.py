"""Generate tests/golden/golden_v1.npz: seeded inputs + the CPU oracle's outputs.

    python scripts/make_golden.py

The reference ships no expected values and cannot be built with its real dependency (Eigen) in this
container.  Two sets of outputs are stored for every case:
  * ``<tag>_x, _iters, _gP, _gq[, _gl, _gm]``  from the oracle restatement (oracle/dq_oracle.c);
  * ``<tag>_xref, _gPref, _gqref[, _glref, _gmref]``  from the reference's OWN qcqplib/Solver.cpp, compiled
    unmodified against the stand-in linear-algebra header (oracle/_ref, `make -C oracle ref`) and run here.
The second set is what pins the oracle: tests/test_oracle.py requires the oracle to reproduce it (bit for
bit for x, grad_P, grad_q of the QP and x of the QCQP; to a few ulp for the QCQP gradients) on machines where
/root/reference does not exist.  ``ref_checked`` records that the reference build was available.
Small on purpose (a few hundred KB): the GPU parity tests use it as a travel-safe fixture.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffqcqp_b200 import workloads as wl  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

try:
    from oracle import pyref
    REF = pyref if pyref.available() else None
except Exception:
    REF = None

out = {}
ref_max = 0.0


def qp(tag, P, q, g, eps):
    global ref_max
    P, q, g = P.numpy(), q.numpy(), g.numpy()
    x, it = orc.qp_forward(P, q, None, eps, 1000, return_iters=True)
    gP, gq = orc.qp_backward(P, q, x, g)
    if REF is not None:
        xr = REF.qp_forward(P, q, None, eps, 1000)
        gPr, gqr = REF.qp_backward(P, q, x, g)
        ref_max = max(ref_max, float(np.abs(xr - x).max() / max(1, np.abs(x).max())))
        out.update({f"{tag}_xref": xr, f"{tag}_gPref": gPr, f"{tag}_gqref": gqr})
    for k, v in dict(P=P, q=q, g=g, x=x, iters=it, gP=gP, gq=gq, eps=np.float64(eps)).items():
        out[f"{tag}_{k}"] = v


def qcqp(tag, P, q, l_n, mu, g, eps):
    global ref_max
    P, q, l_n, mu, g = (a.numpy() for a in (P, q, l_n, mu, g))
    x, it = orc.qcqp_forward(P, q, l_n, mu, None, eps, 1000, return_iters=True)
    gP, gq, gl, gm = orc.qcqp_backward(P, q, l_n, mu, x, g)
    if REF is not None:
        xr = REF.qcqp_forward(P, q, l_n, mu, None, eps, 1000)
        gr = REF.qcqp_backward(P, q, l_n, mu, x, g)
        ref_max = max(ref_max, float(np.abs(xr - x).max() / max(1, np.abs(x).max())))
        out.update({f"{tag}_xref": xr, f"{tag}_gPref": gr[0], f"{tag}_gqref": gr[1], f"{tag}_glref": gr[2],
                    f"{tag}_gmref": gr[3]})
    for k, v in dict(P=P, q=q, l_n=l_n, mu=mu, g=g, x=x, iters=it, gP=gP, gq=gq, gl=gl, gm=gm,
                     eps=np.float64(eps)).items():
        out[f"{tag}_{k}"] = v


qp("qp_diag8", *wl.qp_diag(257, 8, seed=101), 1e-7)
qp("qp_dense8", *wl.qp_dense(130, 8, seed=102), 1e-7)
qp("qp_dense5", *wl.qp_dense(67, 5, seed=103), 1e-10)
qp("qp_dense32", *wl.qp_dense(19, 32, seed=104), 1e-7)
qcqp("qcqp_dense8", *wl.qcqp_dense(131, 8, seed=105), 1e-7)
qcqp("qcqp_dense16", *wl.qcqp_dense(66, 16, seed=106), 1e-7)
qcqp("qcqp_dense24", *wl.qcqp_dense(21, 24, seed=107), 1e-10)
qcqp("qcqp_diag32", *wl.qcqp_dense(17, 32, seed=108, diag=True), 1e-7)
# Solver.cpp:708-712 (the reference's own smoke input): P = diag(5e-4, 3, 0, 0), q = (-8000, 0, 0, 0), eps = 1e-10
import torch  # noqa: E402
Pf = torch.diag(torch.tensor([5e-4, 3.0, 0.0, 0.0], dtype=torch.float64))[None]
qf = torch.tensor([-8000.0, 0, 0, 0], dtype=torch.float64)[None, :, None]
qp("qp_solver_cpp_708", Pf, qf, torch.ones(1, 4, 1, dtype=torch.float64), 1e-10)
out["ref_checked"] = np.bool_(REF is not None)
out["ref_max_dx"] = np.float64(ref_max)
path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes; ref_checked:", REF is not None, "ref_max_dx:", ref_max)

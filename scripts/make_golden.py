"""Generate tests/golden/golden_v1.npz: seeded inputs + the CPU oracle's outputs.

    python scripts/make_golden.py

The reference ships no expected values and cannot be imported/built with its real dependency (Eigen) in
this container, so the golden vectors come from the oracle restatement (oracle/dq_oracle.c).  When the
reference-source build (oracle/_ref: the reference's own Solver.cpp compiled against the stand-in
linear-algebra header) is available, every case is cross-checked against it before being written and
the agreement is recorded in the file (``ref_checked`` / ``ref_max_dx``).
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
        ref_max = max(ref_max, float(np.abs(xr - x).max() / max(1, np.abs(x).max())))
    for k, v in dict(P=P, q=q, g=g, x=x, iters=it, gP=gP, gq=gq, eps=np.float64(eps)).items():
        out[f"{tag}_{k}"] = v


def qcqp(tag, P, q, l_n, mu, g, eps):
    global ref_max
    P, q, l_n, mu, g = (a.numpy() for a in (P, q, l_n, mu, g))
    x, it = orc.qcqp_forward(P, q, l_n, mu, None, eps, 1000, return_iters=True)
    gP, gq, gl, gm = orc.qcqp_backward(P, q, l_n, mu, x, g)
    if REF is not None:
        xr = REF.qcqp_forward(P, q, l_n, mu, None, eps, 1000)
        ref_max = max(ref_max, float(np.abs(xr - x).max() / max(1, np.abs(x).max())))
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
out["ref_checked"] = np.bool_(REF is not None)
out["ref_max_dx"] = np.float64(ref_max)
path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes; ref_checked:", REF is not None, "ref_max_dx:", ref_max)

"""The reference's lower-level surfaces for the hot path, on top of the batched sm_100a kernels.

* the four hot functions of the pybind11 module ``diffqcqp`` (pybindings.cpp:76-82) -- ``solveQP``, ``solveQCQP``,
  ``solveDerivativesQP``, ``solveDerivativesQCQP`` -- with the binding's argument names, defaults and return shapes,
  taking numpy arrays for ONE problem (a batch of one through the same kernels; there is no CPU path here either);
* the unbatched autograd functions of ``qcqp_no_batch.py`` (:23-108): ``QPFn2`` / ``QCQPFn2`` on ``P (N,N)``,
  ``q (N,1)``, ``l_n, mu (N/2,1)``.

* the Box / SignedBox functions of the same module -- ``solveBoxQP``, ``solveSignedBoxQP``, ``solveDerivativesBoxQP``
  (pybindings.cpp:32-52,77-81) -- likewise.

The root-level ``diffqcqp.py`` and ``qcqp_no_batch.py`` re-export these so the reference's import lines keep working.
One kernel launch per call: this surface exists for compatibility, the batched one in ``qcqp.py`` for speed.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function

from . import _lib
from . import qcqp as _b

__all__ = ["solveQP", "solveQCQP", "solveDerivativesQP", "solveDerivativesQCQP", "solveBoxQP", "solveSignedBoxQP",
           "solveDerivativesBoxQP", "QPFn2", "QCQPFn2"]


def _dev():
    if not torch.cuda.is_available():
        raise _lib.DiffQCQPError("diffqcqp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _t(a, shape):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape)).to(_dev())


def solveQP(P, q, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    """pybindings.cpp:17-22,76 -> (N,) array."""
    n = np.asarray(q).size
    x = _b.qp_forward(_t(P, (1, n, n)), _t(q, (1, n, 1)), epsilon, max_iter, mu_prox, adaptative_rho)
    return x.reshape(n).cpu().numpy()


def solveQCQP(P, q, l_n, mu, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    """pybindings.cpp:54-60,79 -> (N,) array."""
    n = np.asarray(q).size
    x = _b.qcqp_forward(_t(P, (1, n, n)), _t(q, (1, n, 1)), _t(l_n, (1, n // 2, 1)), _t(mu, (1, n // 2, 1)), epsilon,
                        max_iter, mu_prox, adaptative_rho)
    return x.reshape(n).cpu().numpy()


def solveDerivativesQP(P, q, l, grad_l, epsilon=1e-10):
    """pybindings.cpp:24-30,80 -> bl (N,).  `epsilon` is accepted for signature parity: like qcqp.py:47 the kernels use
    the binding's default 1e-10 (other values raise rather than being silently ignored)."""
    if epsilon != 1e-10:
        raise ValueError("the backward kernels implement the binding's default epsilon=1e-10 only")
    n = np.asarray(q).size
    _, gq = _b.qp_backward(_t(P, (1, n, n)), _t(q, (1, n, 1)), _t(l, (1, n, 1)), _t(grad_l, (1, n, 1)), need_P=False)
    return (-gq).reshape(n).cpu().numpy()  # grad_q = -dl  (qcqp.py:51)


def solveDerivativesQCQP(P, q, l_n, mu, l, grad_l, epsilon=1e-10):
    """pybindings.cpp:62-71,82 -> (E1 (nc,nc), E2 (nc,nc), blgamma (nc+N,))."""
    if epsilon != 1e-10:
        raise ValueError("the backward kernels implement the binding's default epsilon=1e-10 only")
    n = np.asarray(q).size
    nc = n // 2
    dev = _dev()
    Pd, qd = _t(P, (1, n, n)), _t(q, (1, n, 1))
    ld, md = _t(l_n, (1, nc, 1)), _t(mu, (1, nc, 1))
    xd, gd = _t(l, (1, n, 1)), _t(grad_l, (1, n, 1))
    gq = torch.empty((1, n, 1), dtype=torch.float64, device=dev)
    gam = torch.empty((1, nc, 1), dtype=torch.float64, device=dev)
    dgam = torch.empty((1, nc, 1), dtype=torch.float64, device=dev)
    L = _lib.load()
    with torch.cuda.device(dev):
        rc = L.dq_qcqp_backward_ex(Pd.data_ptr(), qd.data_ptr(), ld.data_ptr(), md.data_ptr(), xd.data_ptr(), gd.data_ptr(),
                                   None, gq.data_ptr(), None, None, gam.data_ptr(), dgam.data_ptr(), 1, n,
                                   torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "dq_qcqp_backward_ex")
    gamma = gam.reshape(nc).cpu().numpy()
    ln, m = np.asarray(l_n, dtype=np.float64).reshape(nc), np.asarray(mu, dtype=np.float64).reshape(nc)
    E1 = np.diag(2 * gamma * ln * ln * m)  # Solver.cpp:686
    E2 = np.diag(2 * gamma * ln * m * m)   # Solver.cpp:687
    blgamma = np.concatenate([dgam.reshape(nc).cpu().numpy(), (-gq).reshape(n).cpu().numpy()])
    return E1, E2, blgamma


def solveBoxQP(P, q, l_min, l_max, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    """pybindings.cpp:32-37,77 -> (N,) array."""
    n = np.asarray(q).size
    x = _b.boxqp_forward(_t(P, (1, n, n)), _t(q, (1, n, 1)), _t(l_min, (1, n, 1)), _t(l_max, (1, n, 1)), epsilon, max_iter,
                         mu_prox, adaptative_rho)
    return x.reshape(n).cpu().numpy()


def solveSignedBoxQP(P, q, l_min, l_max, v, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    """pybindings.cpp:47-52,78 -> (N,) array."""
    n = np.asarray(q).size
    x = _b.boxqp_forward(_t(P, (1, n, n)), _t(q, (1, n, 1)), _t(l_min, (1, n, 1)), _t(l_max, (1, n, 1)), epsilon, max_iter,
                         mu_prox, adaptative_rho, v=_t(v, (1, n, 1)))
    return x.reshape(n).cpu().numpy()


def solveDerivativesBoxQP(P, q, l_min, l_max, l, grad_l, epsilon=1e-10):
    """pybindings.cpp:39-45,81 -> (blgamma (3N,), gamma (2N,)): blgamma = [dgamma_lower ; dgamma_upper ; dl]."""
    if epsilon != 1e-10:
        raise ValueError("the backward kernels implement the binding's default epsilon=1e-10 only")
    n = np.asarray(q).size
    dev = _dev()
    Pd, qd = _t(P, (1, n, n)), _t(q, (1, n, 1))
    lo, hi = _t(l_min, (1, n, 1)), _t(l_max, (1, n, 1))
    xd, gd = _t(l, (1, n, 1)), _t(grad_l, (1, n, 1))
    gq = torch.empty((1, n, 1), dtype=torch.float64, device=dev)
    gam = torch.empty((1, 2 * n), dtype=torch.float64, device=dev)
    dgam = torch.empty((1, 2 * n), dtype=torch.float64, device=dev)
    L = _lib.load()
    with torch.cuda.device(dev):
        rc = L.dq_boxqp_backward_ex(Pd.data_ptr(), qd.data_ptr(), lo.data_ptr(), hi.data_ptr(), xd.data_ptr(), gd.data_ptr(),
                                    None, gq.data_ptr(), None, None, gam.data_ptr(), dgam.data_ptr(), 1, n,
                                    torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "dq_boxqp_backward_ex")
    blgamma = np.concatenate([dgam.reshape(2 * n).cpu().numpy(), (-gq).reshape(n).cpu().numpy()])  # grad_q = -dl
    return blgamma, gam.reshape(2 * n).cpu().numpy()


class QPFn2(Function):
    """Unbatched QP layer (qcqp_no_batch.py:23-55): P (N,N), q (N,1) -> l (N,) ; grads (N,N), (N,1)."""

    @staticmethod
    def forward(ctx, P, q, warm_start, eps, max_iter, mu_prox=1e-7):
        n = q.numel()
        dev = _b._compute_device(P, q)
        Pd, qd = _b._as_dev(P.reshape(1, n, n), dev, "P"), _b._as_dev(q.reshape(1, n, 1), dev, "q")
        x = _b.qp_forward(Pd, qd, eps, max_iter, mu_prox, True)
        ctx.save_for_backward(Pd, qd, x)
        ctx.out_device = q.device
        return x.reshape(n).to(q.device)

    @staticmethod
    def backward(ctx, grad_l):
        Pd, qd, x = ctx.saved_tensors
        n = x.numel()
        g = _b._as_dev(grad_l.reshape(1, n, 1), Pd.device, "grad_l")
        gP, gq = _b.qp_backward(Pd, qd, x, g, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        o = ctx.out_device
        return (None if gP is None else gP.reshape(n, n).to(o)), (None if gq is None else gq.reshape(n, 1).to(o)), None, None, None, None


class QCQPFn2(Function):
    """Unbatched QCQP layer (qcqp_no_batch.py:58-108): P (N,N), q (N,1), l_n, mu (N/2,1) -> l (N,)."""

    @staticmethod
    def forward(ctx, P, q, l_n, mu, warm_start, eps, max_iter, mu_prox=1e-7):
        n = q.numel()
        nc = n // 2
        dev = _b._compute_device(P, q, l_n, mu)
        Pd, qd = _b._as_dev(P.reshape(1, n, n), dev, "P"), _b._as_dev(q.reshape(1, n, 1), dev, "q")
        ld, md = _b._as_dev(l_n.reshape(1, nc, 1), dev, "l_n"), _b._as_dev(mu.reshape(1, nc, 1), dev, "mu")
        x = _b.qcqp_forward(Pd, qd, ld, md, eps, max_iter, mu_prox, True)
        ctx.save_for_backward(Pd, qd, ld, md, x)
        ctx.out_device = q.device
        return x.reshape(n).to(q.device)

    @staticmethod
    def backward(ctx, grad_l):
        Pd, qd, ld, md, x = ctx.saved_tensors
        n = x.numel()
        nc = n // 2
        g = _b._as_dev(grad_l.reshape(1, n, 1), Pd.device, "grad_l")
        gP, gq, gl, gm = _b.qcqp_backward(Pd, qd, ld, md, x, g, tuple(ctx.needs_input_grad[:4]))
        o = ctx.out_device
        r = lambda t, shp: None if t is None else t.reshape(shp).to(o)
        return r(gP, (n, n)), r(gq, (n, 1)), r(gl, (nc, 1)), r(gm, (nc, 1)), None, None, None, None

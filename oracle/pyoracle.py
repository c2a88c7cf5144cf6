"""ctypes front-end of the CPU oracle (oracle/dq_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product package (diffqcqp_b200/) never does.

Parity status: pinned bit for bit against the reference's own Solver.cpp built with a stand-in for Eigen
(oracle/_ref); Eigen's internal summation order is the stated gap -- see oracle/dq_oracle.h.

The functions mirror the reference surface:
  * per-problem: solveQP / solveQCQP / solveDerivativesQP / solveDerivativesQCQP with the argument
    order and defaults of pybindings.cpp:76-82;
  * batched: qp_forward / qp_backward / qcqp_forward / qcqp_backward do what the Python loops of
    qcqp.py:22-52,141-181 do, in one C call, on numpy arrays shaped like the reference's tensors
    ((B,N,N), (B,N,1), (B,nc,1)).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdq_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile oracle/dq_oracle.c (gcc is the only requirement)."""
    src = os.path.join(_HERE, "dq_oracle.c")
    hdr = os.path.join(_HERE, "dq_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libdq_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.dq_oracle_solveQP.restype = ctypes.c_int
        L.dq_oracle_solveQP.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.dq_oracle_solveQCQP.restype = ctypes.c_int
        L.dq_oracle_solveQCQP.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                          ctypes.c_int]
        L.dq_oracle_solveDerivativesQP.restype = None
        L.dq_oracle_solveDerivativesQP.argtypes = [_dp, _dp, _dp, _dp, _dp, ctypes.c_int,
                                                   ctypes.c_double]
        L.dq_oracle_solveDerivativesQCQP.restype = None
        L.dq_oracle_solveDerivativesQCQP.argtypes = [_dp] * 9 + [ctypes.c_int, ctypes.c_double]
        L.dq_oracle_power_iteration.restype = ctypes.c_double
        L.dq_oracle_power_iteration.argtypes = [_dp, ctypes.c_int, ctypes.c_int]
        L.dq_oracle_iterative_refinement.restype = ctypes.c_int
        L.dq_oracle_iterative_refinement.argtypes = [_dp, _dp, _dp, ctypes.c_int]
        L.dq_oracle_qp_forward_batch.restype = None
        L.dq_oracle_qp_forward_batch.argtypes = [_dp, _dp, _dp, _dp, _ip, ctypes.c_int64,
                                                 ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                 ctypes.c_int, ctypes.c_int]
        L.dq_oracle_qp_backward_batch.restype = None
        L.dq_oracle_qp_backward_batch.argtypes = [_dp] * 6 + [ctypes.c_int64, ctypes.c_int,
                                                              ctypes.c_int]
        L.dq_oracle_qcqp_forward_batch.restype = None
        L.dq_oracle_qcqp_forward_batch.argtypes = [_dp] * 6 + [_ip, ctypes.c_int64, ctypes.c_int,
                                                               ctypes.c_double, ctypes.c_double,
                                                               ctypes.c_int, ctypes.c_int]
        L.dq_oracle_qcqp_backward_batch.restype = None
        L.dq_oracle_qcqp_backward_batch.argtypes = [_dp] * 10 + [ctypes.c_int64, ctypes.c_int,
                                                                 ctypes.c_int]
        L.dq_oracle_max_threads.restype = ctypes.c_int
        L.dq_oracle_boxqp_backward_batch.restype = None
        L.dq_oracle_boxqp_backward_batch.argtypes = [_dp] * 10 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
        L.dq_oracle_solveDerivativesBoxQP.restype = None
        L.dq_oracle_solveDerivativesBoxQP.argtypes = [_dp] * 8 + [ctypes.c_int, ctypes.c_double]
        L.dq_oracle_boxqp_forward_batch.restype = None
        L.dq_oracle_boxqp_forward_batch.argtypes = [_dp] * 6 + [_ip, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                                              ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.dq_oracle_set_ir_force.restype = None
        L.dq_oracle_set_ir_force.argtypes = [ctypes.c_int]
        L.dq_oracle_set_rho_nudge.restype = None
        L.dq_oracle_set_rho_nudge.argtypes = [ctypes.c_int]
        L.dq_oracle_set_batch_flags.restype = None
        L.dq_oracle_set_batch_flags.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def max_threads() -> int:
    return int(lib().dq_oracle_max_threads())


def set_batch_flags(flags: int) -> None:
    """Test hook: 1 (default) = what qcqp.py passes; 3 = also start the ADMM iteration at warm_start (the CUDA
    library's DQ_FLAG_WARM_START extension; NOT reference behaviour)."""
    lib().dq_oracle_set_batch_flags(int(flags))


def set_rho_nudge(ulps: int) -> None:
    """Test hook: start every ADMM solve with rho moved by `ulps` ulps (0 restores the reference value)."""
    lib().dq_oracle_set_rho_nudge(int(ulps))


def set_ir_force(n: int) -> None:
    """Test hook: n > 0 makes iterative_refinement run exactly n steps; 0 restores the reference's stop rule."""
    lib().dq_oracle_set_ir_force(int(n))


# ------------------------------------------------------------------ per-problem (pybindings.cpp)
def solveQP(P, q, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True,
            return_iters=False):
    P, q, ws = _c(P), _c(q).reshape(-1), _c(warm_start).reshape(-1)
    n = q.shape[0]
    x = np.empty(n)
    it = lib().dq_oracle_solveQP(_p(P), _p(q), _p(ws), _p(x), n, epsilon, mu_prox, int(max_iter),
                                 int(adaptative_rho))
    return (x, it) if return_iters else x


def solveQCQP(P, q, l_n, mu, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000,
              adaptative_rho=True, return_iters=False):
    P, q, ws = _c(P), _c(q).reshape(-1), _c(warm_start).reshape(-1)
    l_n, mu = _c(l_n).reshape(-1), _c(mu).reshape(-1)
    n = q.shape[0]
    x = np.empty(n)
    it = lib().dq_oracle_solveQCQP(_p(P), _p(q), _p(l_n), _p(mu), _p(ws), _p(x), n, epsilon,
                                   mu_prox, int(max_iter), int(adaptative_rho))
    return (x, it) if return_iters else x


def solveDerivativesQP(P, q, l, grad_l, epsilon=1e-10):
    P, q, l, g = _c(P), _c(q).reshape(-1), _c(l).reshape(-1), _c(grad_l).reshape(-1)
    n = q.shape[0]
    bl = np.empty(n)
    lib().dq_oracle_solveDerivativesQP(_p(P), _p(q), _p(l), _p(g), _p(bl), n, epsilon)
    return bl


def solveDerivativesQCQP(P, q, l_n, mu, l, grad_l, epsilon=1e-10):
    P, q, l, g = _c(P), _c(q).reshape(-1), _c(l).reshape(-1), _c(grad_l).reshape(-1)
    l_n, mu = _c(l_n).reshape(-1), _c(mu).reshape(-1)
    n = q.shape[0]
    nc = n // 2
    E1, E2, blg = np.empty((nc, nc)), np.empty((nc, nc)), np.empty(nc + n)
    lib().dq_oracle_solveDerivativesQCQP(_p(P), _p(q), _p(l_n), _p(mu), _p(l), _p(g), _p(E1),
                                         _p(E2), _p(blg), n, epsilon)
    return E1, E2, blg


def power_iteration(A, max_iter):
    A = _c(A)
    return float(lib().dq_oracle_power_iteration(_p(A), A.shape[0], int(max_iter)))


def iterative_refinement(A, b):
    A, b = _c(A), _c(b).reshape(-1)
    x = np.empty(b.shape[0])
    it = lib().dq_oracle_iterative_refinement(_p(A), _p(b), _p(x), b.shape[0])
    return x, it


# ------------------------------------------------------------------ batched (qcqp.py loops)
def qp_forward(P, q, warm_start, eps, max_iter, mu_prox=1e-7, threads=0, return_iters=False):
    P, q = _c(P), _c(q)
    B, N = P.shape[0], P.shape[1]
    ws = None if warm_start is None else _c(warm_start)
    x = np.empty((B, N, 1))
    iters = np.zeros(B, dtype=np.int32)
    lib().dq_oracle_qp_forward_batch(_p(P), _p(q), _p(ws), _p(x), iters.ctypes.data_as(_ip), B, N,
                                     eps, mu_prox, int(max_iter), threads)
    return (x, iters) if return_iters else x


def boxqp_forward(P, q, l_min, l_max, eps, max_iter, mu_prox=1e-7, v=None, threads=0, return_iters=False):
    """BoxQPFn2.forward (qcqp.py:56-66) / SignedBoxQPFn2.forward (:99-108, when v is given) in one C call."""
    P, q, lo, hi = _c(P), _c(q), _c(l_min), _c(l_max)
    vv = None if v is None else _c(v)
    B, N = P.shape[0], P.shape[1]
    x = np.empty((B, N, 1))
    iters = np.zeros(B, dtype=np.int32)
    lib().dq_oracle_boxqp_forward_batch(_p(P), _p(q), _p(lo), _p(hi), _p(vv), _p(x), iters.ctypes.data_as(_ip), B, N,
                                        eps, mu_prox, int(max_iter), threads)
    return (x, iters) if return_iters else x


def boxqp_backward(P, q, l_min, l_max, x, grad_x, threads=0):
    """BoxQPFn2.backward as qcqp.py:68-94 intends it -> (grad_P, grad_q, grad_l_min, grad_l_max)."""
    P, q, lo, hi, x, g = _c(P), _c(q), _c(l_min), _c(l_max), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    gP, gq, glo, ghi = np.empty((B, N, N)), np.empty((B, N, 1)), np.empty((B, N, 1)), np.empty((B, N, 1))
    lib().dq_oracle_boxqp_backward_batch(_p(P), _p(q), _p(lo), _p(hi), _p(x), _p(g), _p(gP), _p(gq), _p(glo), _p(ghi),
                                         B, N, threads)
    return gP, gq, glo, ghi


def solveDerivativesBoxQP(P, q, l_min, l_max, l, grad_l, epsilon=1e-10):
    """pybindings.cpp:39-45 -> (blgamma (3N,), gamma (2N,))."""
    P, q, l, g = _c(P), _c(q).reshape(-1), _c(l).reshape(-1), _c(grad_l).reshape(-1)
    lo, hi = _c(l_min).reshape(-1), _c(l_max).reshape(-1)
    n = q.shape[0]
    blg, gam = np.empty(3 * n), np.empty(2 * n)
    lib().dq_oracle_solveDerivativesBoxQP(_p(P), _p(q), _p(lo), _p(hi), _p(l), _p(g), _p(blg), _p(gam), n, epsilon)
    return blg, gam


def qp_backward(P, q, x, grad_x, threads=0, need_P=True, need_q=True):
    P, q, x, g = _c(P), _c(q), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    gP = np.empty((B, N, N)) if need_P else None
    gq = np.empty((B, N, 1)) if need_q else None
    lib().dq_oracle_qp_backward_batch(_p(P), _p(q), _p(x), _p(g), _p(gP), _p(gq), B, N, threads)
    return gP, gq


def qcqp_forward(P, q, l_n, mu, warm_start, eps, max_iter, mu_prox=1e-7, threads=0,
                 return_iters=False):
    P, q, l_n, mu = _c(P), _c(q), _c(l_n), _c(mu)
    B, N = P.shape[0], P.shape[1]
    ws = None if warm_start is None else _c(warm_start)
    x = np.empty((B, N, 1))
    iters = np.zeros(B, dtype=np.int32)
    lib().dq_oracle_qcqp_forward_batch(_p(P), _p(q), _p(l_n), _p(mu), _p(ws), _p(x),
                                       iters.ctypes.data_as(_ip), B, N, eps, mu_prox,
                                       int(max_iter), threads)
    return (x, iters) if return_iters else x


def qcqp_backward(P, q, l_n, mu, x, grad_x, threads=0):
    P, q, l_n, mu, x, g = _c(P), _c(q), _c(l_n), _c(mu), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    nc = N // 2
    gP, gq = np.empty((B, N, N)), np.empty((B, N, 1))
    gl, gm = np.empty((B, nc, 1)), np.empty((B, nc, 1))
    lib().dq_oracle_qcqp_backward_batch(_p(P), _p(q), _p(l_n), _p(mu), _p(x), _p(g), _p(gP),
                                        _p(gq), _p(gl), _p(gm), B, N, threads)
    return gP, gq, gl, gm

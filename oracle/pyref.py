"""ctypes front-end of oracle/_ref/libdq_ref.so: the reference's OWN qcqplib/Solver.cpp, compiled
unmodified from where it lies against oracle/eigen_standin (Eigen itself is not installed).
TEST INFRASTRUCTURE ONLY -- same rules as oracle/pyoracle.py.

Built by ``make -C oracle ref`` (needs /root/reference, i.e. the authoring container); the GPU box only
uses the prebuilt .so, which travels with the repo snapshot.  ``available()`` says whether it is there.
The batched functions have the signatures of oracle/pyoracle.py so either can serve as the checker or
as bench.py's CPU baseline (kind "reference").
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdq_ref.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        i, i64, d = ctypes.c_int, ctypes.c_int64, ctypes.c_double
        L.dq_ref_max_threads.restype = i
        L.dq_ref_solveQP.argtypes = [_dp] * 4 + [i, d, d, i, i]
        L.dq_ref_solveDerivativesQP.argtypes = [_dp] * 5 + [i, d]
        L.dq_ref_solveQCQP.argtypes = [_dp] * 6 + [i, d, d, i, i]
        L.dq_ref_solveDerivativesQCQP.argtypes = [_dp] * 9 + [i, d]
        L.dq_ref_qp_forward_batch.argtypes = [_dp] * 4 + [i64, i, d, d, i, i]
        L.dq_ref_qp_backward_batch.argtypes = [_dp] * 6 + [i64, i, i]
        L.dq_ref_qcqp_forward_batch.argtypes = [_dp] * 6 + [i64, i, d, d, i, i]
        L.dq_ref_qcqp_backward_batch.argtypes = [_dp] * 10 + [i64, i, i]
        L.dq_ref_boxqp_backward_batch.argtypes = [_dp] * 10 + [i64, i, i]
        L.dq_ref_boxqp_backward_batch.restype = None
        L.dq_ref_boxqp_forward_batch.argtypes = [_dp] * 6 + [i64, i, d, d, i, i]
        L.dq_ref_boxqp_forward_batch.restype = None
        for f in ("dq_ref_solveQP", "dq_ref_solveDerivativesQP", "dq_ref_solveQCQP", "dq_ref_solveDerivativesQCQP",
                  "dq_ref_qp_forward_batch", "dq_ref_qp_backward_batch", "dq_ref_qcqp_forward_batch",
                  "dq_ref_qcqp_backward_batch"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def max_threads() -> int:
    return int(lib().dq_ref_max_threads())


# ------------------------------------------------------------------ per-problem (pybindings.cpp:76-82)
def solveQP(P, q, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    P, q, ws = _c(P), _c(q).reshape(-1), _c(warm_start).reshape(-1)
    x = np.empty(q.shape[0])
    lib().dq_ref_solveQP(_p(P), _p(q), _p(ws), _p(x), q.shape[0], epsilon, mu_prox, int(max_iter), int(adaptative_rho))
    return x


def solveQCQP(P, q, l_n, mu, warm_start, epsilon=1e-10, mu_prox=1e-7, max_iter=1000, adaptative_rho=True):
    P, q, ws = _c(P), _c(q).reshape(-1), _c(warm_start).reshape(-1)
    l_n, mu = _c(l_n).reshape(-1), _c(mu).reshape(-1)
    x = np.empty(q.shape[0])
    lib().dq_ref_solveQCQP(_p(P), _p(q), _p(l_n), _p(mu), _p(ws), _p(x), q.shape[0], epsilon, mu_prox, int(max_iter),
                           int(adaptative_rho))
    return x


def solveDerivativesQP(P, q, l, grad_l, epsilon=1e-10):
    P, q, l, g = _c(P), _c(q).reshape(-1), _c(l).reshape(-1), _c(grad_l).reshape(-1)
    bl = np.empty(q.shape[0])
    lib().dq_ref_solveDerivativesQP(_p(P), _p(q), _p(l), _p(g), _p(bl), q.shape[0], epsilon)
    return bl


def solveDerivativesQCQP(P, q, l_n, mu, l, grad_l, epsilon=1e-10):
    P, q, l, g = _c(P), _c(q).reshape(-1), _c(l).reshape(-1), _c(grad_l).reshape(-1)
    l_n, mu = _c(l_n).reshape(-1), _c(mu).reshape(-1)
    n = q.shape[0]
    nc = n // 2
    E1, E2, blg = np.empty((nc, nc)), np.empty((nc, nc)), np.empty(nc + n)
    lib().dq_ref_solveDerivativesQCQP(_p(P), _p(q), _p(l_n), _p(mu), _p(l), _p(g), _p(E1), _p(E2), _p(blg), n, epsilon)
    return E1, E2, blg


# ------------------------------------------------------------------ batched (qcqp.py loops)
def qp_forward(P, q, warm_start, eps, max_iter, mu_prox=1e-7, threads=0):
    P, q = _c(P), _c(q)
    B, N = P.shape[0], P.shape[1]
    ws = None if warm_start is None else _c(warm_start)
    x = np.empty((B, N, 1))
    lib().dq_ref_qp_forward_batch(_p(P), _p(q), _p(ws), _p(x), B, N, eps, mu_prox, int(max_iter), threads)
    return x


def boxqp_forward(P, q, l_min, l_max, eps, max_iter, mu_prox=1e-7, v=None, threads=0):
    P, q, lo, hi = _c(P), _c(q), _c(l_min), _c(l_max)
    vv = None if v is None else _c(v)
    B, N = P.shape[0], P.shape[1]
    x = np.empty((B, N, 1))
    lib().dq_ref_boxqp_forward_batch(_p(P), _p(q), _p(lo), _p(hi), _p(vv), _p(x), B, N, eps, mu_prox, int(max_iter), threads)
    return x


def boxqp_backward(P, q, l_min, l_max, x, grad_x, threads=0):
    P, q, lo, hi, x, g = _c(P), _c(q), _c(l_min), _c(l_max), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    gP, gq, glo, ghi = np.empty((B, N, N)), np.empty((B, N, 1)), np.empty((B, N, 1)), np.empty((B, N, 1))
    lib().dq_ref_boxqp_backward_batch(_p(P), _p(q), _p(lo), _p(hi), _p(x), _p(g), _p(gP), _p(gq), _p(glo), _p(ghi), B, N,
                                      threads)
    return gP, gq, glo, ghi


def qp_backward(P, q, x, grad_x, threads=0):
    P, q, x, g = _c(P), _c(q), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    gP, gq = np.empty((B, N, N)), np.empty((B, N, 1))
    lib().dq_ref_qp_backward_batch(_p(P), _p(q), _p(x), _p(g), _p(gP), _p(gq), B, N, threads)
    return gP, gq


def qcqp_forward(P, q, l_n, mu, warm_start, eps, max_iter, mu_prox=1e-7, threads=0):
    P, q, l_n, mu = _c(P), _c(q), _c(l_n), _c(mu)
    B, N = P.shape[0], P.shape[1]
    ws = None if warm_start is None else _c(warm_start)
    x = np.empty((B, N, 1))
    lib().dq_ref_qcqp_forward_batch(_p(P), _p(q), _p(l_n), _p(mu), _p(ws), _p(x), B, N, eps, mu_prox, int(max_iter),
                                    threads)
    return x


def qcqp_backward(P, q, l_n, mu, x, grad_x, threads=0):
    P, q, l_n, mu, x, g = _c(P), _c(q), _c(l_n), _c(mu), _c(x), _c(grad_x)
    B, N = P.shape[0], P.shape[1]
    nc = N // 2
    gP, gq = np.empty((B, N, N)), np.empty((B, N, 1))
    gl, gm = np.empty((B, nc, 1)), np.empty((B, nc, 1))
    lib().dq_ref_qcqp_backward_batch(_p(P), _p(q), _p(l_n), _p(mu), _p(x), _p(g), _p(gP), _p(gq), _p(gl), _p(gm), B, N,
                                     threads)
    return gP, gq, gl, gm

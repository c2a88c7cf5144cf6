"""Drop-in for the hot path of the reference's ``qcqp.py``: ``QPFn2`` and ``QCQPFn2``.

Same class names, argument order, tensor shapes and gradient-tuple arity as the reference
(qcqp.py:22-52 and :141-181).  What differs is what runs underneath: instead of a Python loop over the
batch calling a per-problem C++ solver (qcqp.py:29-31, :45-47, :149-151, :167-168), each of
forward / backward is ONE launch of an sm_100a CUDA kernel over the whole batch through the C ABI in
``include/diffqcqp_b200.h``.

Devices.  The reference is CPU-only (``.numpy()`` at qcqp.py:30).  Here:
  * CUDA tensors in -> CUDA tensors out, asynchronous on the current stream, no host sync;
  * CPU tensors in (what a user of the reference has) -> inputs are copied to the current CUDA
    device, solved there, and the result is returned as a CPU tensor; device copies are kept on the
    autograd context so backward only moves ``grad_l`` in and the gradients out.  Batches of
    >= HOST_PIPE_MIN_BATCH problems move through a chunked copy/compute pipeline (copies on side
    streams overlap the kernels; outputs are pinned CPU tensors).
There is no CPU compute path: without the CUDA extension or a CUDA device this raises.

Notes carried over from the reference's behaviour:
  * ``warm_start`` is accepted and ignored, as in the reference (Solver.cpp:70 -> :80 overwrites it) -- unless the
    caller opts into the extension with ``use_warm_start(True)``: the ADMM iteration then starts
    at ``warm_start`` instead of zero (fewer iterations when it is close to the solution, e.g. the previous time step
    of a simulator; results then differ from the reference's within its own stopping tolerance).
  * backward uses the binding's default ``epsilon=1e-10`` regardless of the forward ``eps``
    (qcqp.py:47,168; pybindings.cpp:80,82).
  * ``adaptative_rho`` is fixed to True (qcqp.py:27,148).
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

__all__ = ["QPFn2", "QCQPFn2", "BoxQPFn2", "SignedBoxQPFn2", "qp_forward", "qp_backward", "qcqp_forward",
           "qcqp_backward", "boxqp_forward", "boxqp_backward", "use_warm_start"]

FLAG_ADAPTIVE_RHO, FLAG_WARM_START = 1, 2  # include/diffqcqp_b200.h: DQ_FLAG_*
_warm_start_enabled = False


class use_warm_start:
    """Opt into the warm-start extension for the layers (``QPFn2`` ... read ``warm_start`` instead of ignoring it).
    ``use_warm_start(True)`` switches it on process-wide; as a context manager it restores the previous setting."""

    def __init__(self, enabled: bool = True):
        global _warm_start_enabled
        self.prev = _warm_start_enabled
        _warm_start_enabled = bool(enabled)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        global _warm_start_enabled
        _warm_start_enabled = self.prev
        return False


def _flags(adaptative_rho, warm_start):
    return (FLAG_ADAPTIVE_RHO if adaptative_rho else 0) | (FLAG_WARM_START if warm_start is not None else 0)


def _layer_warm(warm_start, dev, like):
    """The layers' warm_start argument: None (ignored, the reference's behaviour) unless use_warm_start is on."""
    if not _warm_start_enabled or warm_start is None:
        return None
    if tuple(warm_start.shape) != tuple(like.shape):
        raise ValueError(f"warm_start must have shape {tuple(like.shape)}, got {tuple(warm_start.shape)}")
    return _as_dev(warm_start, dev, "warm_start")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _as_dev(t: torch.Tensor, device, name: str) -> torch.Tensor:
    """fp64, contiguous, on `device`, 32-byte aligned (the 256-bit load / store paths want that)."""
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.dtype.is_floating_point:
        raise ValueError(f"{name} must be a floating-point tensor, got {t.dtype}")
    t = t.detach()
    if t.device != device or t.dtype != torch.float64:
        t = t.to(device=device, dtype=torch.float64, non_blocking=True)
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 32:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def _compute_device(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise _lib.DiffQCQPError(
            "diffqcqp_b200 needs a CUDA device (sm_100a); there is no CPU fallback in the product path")
    return torch.device("cuda", torch.cuda.current_device())


MAX_N = 128  # DQ_MAX_N of include/diffqcqp_b200.h (dq_max_n()); N <= 32 runs the warp-tile kernels, above that a slow warp-per-problem path


def _check_shapes(P, q, l_n=None, mu=None):
    if P.dim() != 3 or P.size(1) != P.size(2):
        raise ValueError(f"P must have shape (B,N,N), got {tuple(P.shape)}")
    B, N = P.size(0), P.size(1)
    if N > MAX_N:
        raise ValueError(f"N={N} exceeds the {MAX_N} unknowns per problem these kernels support (QPs up to N={MAX_N}, QCQPs up to "
                         f"{MAX_N // 2} contacts); the reference's solveQP / solveQCQP have no such limit")
    if q.dim() != 3 or tuple(q.shape) != (B, N, 1):
        raise ValueError(f"q must have shape (B,N,1)=({B},{N},1), got {tuple(q.shape)}")
    if l_n is not None:
        if N % 2:
            raise ValueError(f"the QCQP needs an even N (two tangential components per contact), got N={N}")
        for nm, t in (("l_n", l_n), ("mu", mu)):
            if tuple(t.shape) != (B, N // 2, 1):
                raise ValueError(f"{nm} must have shape (B,N/2,1)=({B},{N // 2},1), got {tuple(t.shape)}")
    return B, N


# --------------------------------------------------------------------------- raw batched ops
def qp_forward(P, q, eps, max_iter, mu_prox=1e-7, adaptative_rho=True, return_iters=False, warm_start=None, state=None,
               out=None):
    """Batched solveQP on CUDA tensors: P (B,N,N), q (B,N,1) -> x (B,N,1) [, iters (B,) int32].
    warm_start (B,N,1), when given, is where the iteration starts (extension; None = the reference's behaviour).
    out: optional contiguous (B,N,1) CUDA tensor to write x into."""
    dev = P.device
    B, N = P.size(0), P.size(1)
    x = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if out is None else out
    iters = torch.empty((B,), dtype=torch.int32, device=dev) if return_iters else None
    L = _lib.load()
    with torch.cuda.device(dev):
        # state: optional (B,N,1) buffer the forward fills for the backward (diag(P) of diagonal problems, NaN otherwise)
        rc = L.dq_qp_forward_ex(_ptr(P), _ptr(q), _ptr(warm_start), _ptr(x), _ptr(iters), _ptr(state), B, N, float(eps),
                                float(mu_prox), int(max_iter), _flags(adaptative_rho, warm_start), _stream_ptr(dev))
    _lib.check(rc, "dq_qp_forward")
    return (x, iters) if return_iters else x


def qp_backward(P, q, x, grad_x, need_P=True, need_q=True, state=None, out=None):
    """Batched solveDerivativesQP + the products of qcqp.py:48-51 -> (grad_P, grad_q); out = optional (grad_P, grad_q) to fill."""
    dev = P.device
    B, N = P.size(0), P.size(1)
    if out is not None:
        gP, gq = out
        need_P, need_q = gP is not None, gq is not None
    else:
        gP = torch.empty((B, N, N), dtype=torch.float64, device=dev) if need_P else None
        gq = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if need_q else None
    if need_P or need_q:
        L = _lib.load()
        with torch.cuda.device(dev):
            rc = L.dq_qp_backward_ex(_ptr(P), _ptr(q), _ptr(x), _ptr(grad_x), _ptr(state), _ptr(gP), _ptr(gq), B, N,
                                     _stream_ptr(dev))
        _lib.check(rc, "dq_qp_backward")
    return gP, gq


def qcqp_forward(P, q, l_n, mu, eps, max_iter, mu_prox=1e-7, adaptative_rho=True, return_iters=False, warm_start=None,
                 state=None, out=None):
    dev = P.device
    B, N = P.size(0), P.size(1)
    x = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if out is None else out
    iters = torch.empty((B,), dtype=torch.int32, device=dev) if return_iters else None
    L = _lib.load()
    with torch.cuda.device(dev):
        rc = L.dq_qcqp_forward_ex(_ptr(P), _ptr(q), _ptr(l_n), _ptr(mu), _ptr(warm_start), _ptr(x), _ptr(iters),
                                  _ptr(state), B, N, float(eps), float(mu_prox), int(max_iter),
                                  _flags(adaptative_rho, warm_start), _stream_ptr(dev))
    _lib.check(rc, "dq_qcqp_forward")
    return (x, iters) if return_iters else x


def qcqp_backward(P, q, l_n, mu, x, grad_x, need=(True, True, True, True), state=None, out=None):
    dev = P.device
    B, N = P.size(0), P.size(1)
    nc = N // 2
    if out is not None:
        gP, gq, gl, gm = out
        need = tuple(t is not None for t in out)
    else:
        gP = torch.empty((B, N, N), dtype=torch.float64, device=dev) if need[0] else None
        gq = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if need[1] else None
        gl = torch.empty((B, nc, 1), dtype=torch.float64, device=dev) if need[2] else None
        gm = torch.empty((B, nc, 1), dtype=torch.float64, device=dev) if need[3] else None
    if any(need):
        L = _lib.load()
        with torch.cuda.device(dev):
            rc = L.dq_qcqp_backward_ex2(_ptr(P), _ptr(q), _ptr(l_n), _ptr(mu), _ptr(x), _ptr(grad_x), _ptr(state), _ptr(gP),
                                        _ptr(gq), _ptr(gl), _ptr(gm), None, None, B, N, _stream_ptr(dev))
        _lib.check(rc, "dq_qcqp_backward")
    return gP, gq, gl, gm


def boxqp_forward(P, q, l_min, l_max, eps, max_iter, mu_prox=1e-7, adaptative_rho=True, v=None, return_iters=False,
                  warm_start=None):
    """Batched solveBoxQP (v is None) / solveSignedBoxQP on CUDA tensors: l_min, l_max[, v] are (B,N,1)."""
    dev = P.device
    B, N = P.size(0), P.size(1)
    x = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    iters = torch.empty((B,), dtype=torch.int32, device=dev) if return_iters else None
    L = _lib.load()
    with torch.cuda.device(dev):
        rc = L.dq_boxqp_forward(_ptr(P), _ptr(q), _ptr(l_min), _ptr(l_max), _ptr(v), _ptr(warm_start), _ptr(x), _ptr(iters),
                                B, N, float(eps), float(mu_prox), int(max_iter), _flags(adaptative_rho, warm_start),
                                _stream_ptr(dev))
    _lib.check(rc, "dq_boxqp_forward")
    return (x, iters) if return_iters else x


def boxqp_backward(P, q, l_min, l_max, x, grad_x, need=(True, True, True, True)):
    dev = P.device
    B, N = P.size(0), P.size(1)
    gP = torch.empty((B, N, N), dtype=torch.float64, device=dev) if need[0] else None
    gq, glo, ghi = (torch.empty((B, N, 1), dtype=torch.float64, device=dev) if n else None for n in need[1:4])
    if any(need):
        L = _lib.load()
        with torch.cuda.device(dev):
            rc = L.dq_boxqp_backward(_ptr(P), _ptr(q), _ptr(l_min), _ptr(l_max), _ptr(x), _ptr(grad_x), _ptr(gP), _ptr(gq),
                                     _ptr(glo), _ptr(ghi), B, N, _stream_ptr(dev))
        _lib.check(rc, "dq_boxqp_backward")
    return gP, gq, glo, ghi


def _back_to(t, device):
    if t is None or t.device == device:
        return t
    return t.to(device)


# --------------------------------------------------------------------------- CPU tensors: pipelined copies
# A user of the reference holds CPU tensors.  For batches worth it the layers move P to the device in chunks on a copy
# stream while the forward kernel already solves the chunks that have arrived, and in backward read grad_P back chunk by
# chunk on another copy stream while later chunks are still being differentiated (the bulk of the traffic is P in and
# grad_P out: 8 N^2 bytes per problem each; the vectors are 1/N of that and move whole).  Pinned (page-locked) input
# tensors make the copies truly asynchronous; pageable ones work, staged by the driver.  Outputs are pinned CPU tensors.
HOST_PIPE_MIN_BATCH = 4096  # below this a plain copy is as fast
HOST_PIPE_CHUNKS = int(os.environ.get("DQ_PIPE_CHUNKS", "2"))  # the kernels are a small fraction of the copy time: few chunks, little launch overhead
_side_streams = {}


def _pipe_streams(dev):
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _side_streams[key]


def _chunk_bounds(B, n=None):
    n = n or HOST_PIPE_CHUNKS
    step = -(-B // n)
    step = max(4, (step + 3) & ~3)  # chunk starts stay 32-byte aligned for every N
    return [(c0, min(B, c0 + step)) for c0 in range(0, B, step)]


def _use_host_pipe(P):
    return (not P.is_cuda) and P.size(0) >= HOST_PIPE_MIN_BATCH


def _cpu_f64(t):
    t = t.detach()
    if t.dtype != torch.float64:
        t = t.double()
    return t if t.is_contiguous() else t.contiguous()


def _pipe_forward(dev, P, small, launch):
    """P (CPU, (B,N,N)) to the device chunk by chunk; launch(c0, c1, Pd, smalls_dev, xd) enqueues the forward of a chunk on
    the current stream; x comes back chunk by chunk into a pinned CPU tensor.  Returns (Pd, smalls_dev, xd, x_cpu)."""
    B, N = P.size(0), P.size(1)
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _pipe_streams(dev)
    Pc = _cpu_f64(P)
    smalls = [None if t is None else _cpu_f64(t).to(dev, non_blocking=True) for t in small]
    Pd = torch.empty((B, N, N), dtype=torch.float64, device=dev)
    xd = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    x_cpu = torch.empty((B, N, 1), dtype=torch.float64, pin_memory=True)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    for c0, c1 in _chunk_bounds(B):
        with torch.cuda.stream(s_in):
            Pd[c0:c1].copy_(Pc[c0:c1], non_blocking=True)
            e_in = torch.cuda.Event()
            e_in.record(s_in)
        cur.wait_event(e_in)
        launch(c0, c1, Pd, smalls, xd)
        e_k = torch.cuda.Event()
        e_k.record(cur)
        s_out.wait_event(e_k)
        with torch.cuda.stream(s_out):
            x_cpu[c0:c1].copy_(xd[c0:c1], non_blocking=True)
    s_out.synchronize()  # x_cpu is complete (and with it every copy and kernel issued above)
    return Pd, smalls, xd, x_cpu


def _pipe_backward(dev, B, N, grad_l, need_P, launch, small_shapes):
    """launch(c0, c1, gd, gPd, smalls) enqueues the backward of a chunk; grad_P goes back chunk by chunk into a pinned CPU
    tensor, the small gradients (shapes in small_shapes, None = not needed) whole.  Returns (grad_P_cpu, [small grads cpu])."""
    cur = torch.cuda.current_stream(dev)
    gd = _cpu_f64(grad_l).to(dev, non_blocking=True)
    gPd = torch.empty((B, N, N), dtype=torch.float64, device=dev) if need_P else None
    gP_cpu = torch.empty((B, N, N), dtype=torch.float64, pin_memory=True) if need_P else None
    smalls = [None if sh is None else torch.empty(sh, dtype=torch.float64, device=dev) for sh in small_shapes]
    # (Letting the kernel store grad_P straight into the page-locked output -- it is mapped into the device's address space
    # -- was measured and is slower on this platform: 2.28 ms instead of 1.73 ms per B=65536 N=8 step; copy engines it is.)
    _, s_out = _pipe_streams(dev)
    s_out.wait_stream(cur)
    for c0, c1 in _chunk_bounds(B):
        launch(c0, c1, gd, gPd, smalls)
        if need_P:
            e_k = torch.cuda.Event()
            e_k.record(cur)
            s_out.wait_event(e_k)
            with torch.cuda.stream(s_out):
                gP_cpu[c0:c1].copy_(gPd[c0:c1], non_blocking=True)
    outs = []
    for t in smalls:
        if t is None:
            outs.append(None)
        else:
            o = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
            o.copy_(t, non_blocking=True)  # on the current stream, after the last chunk's kernel
            outs.append(o)
    cur.synchronize()
    s_out.synchronize()
    return gP_cpu, outs


def _sl(t, c0, c1):
    return None if t is None else t[c0:c1]


# --------------------------------------------------------------------------- autograd surface
class QPFn2(Function):
    """min 1/2 l'Pl + q'l  s.t. l >= 0, batched.  Mirrors qcqp.py:22-52."""

    @staticmethod
    def forward(ctx, P, q, warm_start, eps, max_iter, mu_prox=1e-7):
        B, N = _check_shapes(P, q)
        dev = _compute_device(P, q)
        want_grad = any(ctx.needs_input_grad[:2])
        ctx.out_device = q.device
        ctx.host_pipe = _use_host_pipe(P) and not q.is_cuda
        if ctx.host_pipe:
            state = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if want_grad else None
            ws = _layer_warm(warm_start, dev, q)

            def launch(c0, c1, Pd, smalls, xd):
                qp_forward(Pd[c0:c1], smalls[0][c0:c1], eps, max_iter, mu_prox, True, warm_start=_sl(ws, c0, c1),
                           state=_sl(state, c0, c1), out=xd[c0:c1])

            Pd, (qd,), x, x_cpu = _pipe_forward(dev, P, [q], launch)
            ctx.save_for_backward(Pd, qd, x, state)
            return x_cpu
        Pd, qd = _as_dev(P, dev, "P"), _as_dev(q, dev, "q")
        # forward -> backward hand-off: diag(P) of the problems solved on the diagonal path (the backward then skips P)
        state = torch.empty_like(qd) if want_grad else None
        x = qp_forward(Pd, qd, eps, max_iter, mu_prox, True, warm_start=_layer_warm(warm_start, dev, q), state=state)
        ctx.save_for_backward(Pd, qd, x, state)
        return _back_to(x, q.device)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_l):
        Pd, qd, x, state = ctx.saved_tensors
        need_P, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if ctx.host_pipe and not grad_l.is_cuda and (need_P or need_q):
            B, N = Pd.size(0), Pd.size(1)

            def launch(c0, c1, gd, gPd, smalls):
                qp_backward(Pd[c0:c1], qd[c0:c1], x[c0:c1], gd[c0:c1], state=_sl(state, c0, c1),
                            out=(_sl(gPd, c0, c1), _sl(smalls[0], c0, c1)))

            gP, (gq,) = _pipe_backward(Pd.device, B, N, grad_l, need_P, launch, [(B, N, 1) if need_q else None])
            return gP, gq, None, None, None, None
        g = _as_dev(grad_l, Pd.device, "grad_l")
        gP, gq = qp_backward(Pd, qd, x, g, need_P, need_q, state=state)
        return _back_to(gP, ctx.out_device), _back_to(gq, ctx.out_device), None, None, None, None


class QCQPFn2(Function):
    """min 1/2 l'Pl + q'l  s.t. |(l_2i, l_2i+1)| <= l_n[i] mu[i], batched.  Mirrors qcqp.py:141-181."""

    @staticmethod
    def forward(ctx, P, q, l_n, mu, warm_start, eps, max_iter, mu_prox=1e-7):
        B, N = _check_shapes(P, q, l_n, mu)
        dev = _compute_device(P, q, l_n, mu)
        want_grad = any(ctx.needs_input_grad[:4])
        ctx.out_device = q.device
        ctx.host_pipe = _use_host_pipe(P) and not (q.is_cuda or l_n.is_cuda or mu.is_cuda)
        if ctx.host_pipe:
            state = torch.empty((B, N, 1), dtype=torch.float64, device=dev) if want_grad else None
            ws = _layer_warm(warm_start, dev, q)

            def launch(c0, c1, Pd, smalls, xd):
                qcqp_forward(Pd[c0:c1], smalls[0][c0:c1], smalls[1][c0:c1], smalls[2][c0:c1], eps, max_iter, mu_prox, True,
                             warm_start=_sl(ws, c0, c1), state=_sl(state, c0, c1), out=xd[c0:c1])

            Pd, (qd, ld, md), x, x_cpu = _pipe_forward(dev, P, [q, l_n, mu], launch)
            ctx.save_for_backward(Pd, qd, ld, md, x, state)
            return x_cpu
        Pd, qd = _as_dev(P, dev, "P"), _as_dev(q, dev, "q")
        ld, md = _as_dev(l_n, dev, "l_n"), _as_dev(mu, dev, "mu")
        state = torch.empty_like(qd) if want_grad else None  # forward -> backward hand-off
        x = qcqp_forward(Pd, qd, ld, md, eps, max_iter, mu_prox, True, warm_start=_layer_warm(warm_start, dev, q),
                         state=state)
        ctx.save_for_backward(Pd, qd, ld, md, x, state)
        return _back_to(x, q.device)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_l):
        Pd, qd, ld, md, x, state = ctx.saved_tensors
        need = tuple(ctx.needs_input_grad[:4])
        if ctx.host_pipe and not grad_l.is_cuda and any(need):
            B, N = Pd.size(0), Pd.size(1)
            nc = N // 2

            def launch(c0, c1, gd, gPd, smalls):
                qcqp_backward(Pd[c0:c1], qd[c0:c1], ld[c0:c1], md[c0:c1], x[c0:c1], gd[c0:c1], state=_sl(state, c0, c1),
                              out=(_sl(gPd, c0, c1), _sl(smalls[0], c0, c1), _sl(smalls[1], c0, c1), _sl(smalls[2], c0, c1)))

            shapes = [(B, N, 1) if need[1] else None, (B, nc, 1) if need[2] else None, (B, nc, 1) if need[3] else None]
            gP, (gq, gl, gm) = _pipe_backward(Pd.device, B, N, grad_l, need[0], launch, shapes)
            return gP, gq, gl, gm, None, None, None, None
        g = _as_dev(grad_l, Pd.device, "grad_l")
        gP, gq, gl, gm = qcqp_backward(Pd, qd, ld, md, x, g, need, state=state)
        o = ctx.out_device
        return _back_to(gP, o), _back_to(gq, o), _back_to(gl, o), _back_to(gm, o), None, None, None, None


def _check_box(P, q, *bounds):
    B, N = _check_shapes(P, q)
    for nm, t in bounds:
        if tuple(t.shape) != (B, N, 1):
            raise ValueError(f"{nm} must have shape (B,N,1)=({B},{N},1), got {tuple(t.shape)}")


class BoxQPFn2(Function):
    """min 1/2 l'Pl + q'l  s.t. l_min <= l <= l_max, batched.  Forward mirrors qcqp.py:56-66
    (solveBoxQP); backward is what qcqp.py:68-94 sets out to compute from solveDerivativesBoxQP -- that code cannot
    run as shipped (it unpacks six names from four values, reads l_min/l_max swapped and calls Tensor.asDiagonal), so
    the gradients are defined from the C++ (Solver.cpp:263-371): grad_P = -dl l', grad_q = -dl,
    grad_l_min = -dgamma_lower gamma_lower, grad_l_max = +dgamma_upper gamma_upper (finite-difference checked)."""

    @staticmethod
    def forward(ctx, P, q, l_min, l_max, warm_start, eps, max_iter, mu_prox=1e-7):
        _check_box(P, q, ("l_min", l_min), ("l_max", l_max))
        dev = _compute_device(P, q, l_min, l_max)
        t = [_as_dev(a, dev, n) for a, n in ((P, "P"), (q, "q"), (l_min, "l_min"), (l_max, "l_max"))]
        x = boxqp_forward(*t, eps, max_iter, mu_prox, True, warm_start=_layer_warm(warm_start, dev, q))
        ctx.save_for_backward(*t, x)
        ctx.out_device = q.device
        return _back_to(x, q.device)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_l):
        Pd, qd, lo, hi, x = ctx.saved_tensors
        g = _as_dev(grad_l, Pd.device, "grad_l")
        grads = boxqp_backward(Pd, qd, lo, hi, x, g, tuple(ctx.needs_input_grad[:4]))
        o = ctx.out_device
        return tuple(_back_to(t, o) for t in grads) + (None, None, None, None)


class SignedBoxQPFn2(Function):
    """Box QP with the extra constraint sign(v_i) l_i <= 0.  Forward mirrors qcqp.py:99-108 (solveSignedBoxQP);
    the reference has no backward for it (qcqp.py:111 'npt implemented', no solveDerivativesSignedBoxQP exists)."""

    @staticmethod
    def forward(ctx, P, q, l_min, l_max, v, warm_start, eps, max_iter, mu_prox=1e-7):
        _check_box(P, q, ("l_min", l_min), ("l_max", l_max), ("v", v))
        dev = _compute_device(P, q, l_min, l_max, v)
        t = [_as_dev(a, dev, n) for a, n in ((P, "P"), (q, "q"), (l_min, "l_min"), (l_max, "l_max"))]
        x = boxqp_forward(*t, eps, max_iter, mu_prox, True, v=_as_dev(v, dev, "v"),
                          warm_start=_layer_warm(warm_start, dev, q))
        return _back_to(x, q.device)

    @staticmethod
    def backward(ctx, grad_l):
        raise NotImplementedError("SignedBoxQPFn2 has no backward in the reference either (qcqp.py:111)")

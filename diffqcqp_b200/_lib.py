"""ctypes binding of the C ABI in include/diffqcqp_b200.h.

There is no CPU fallback: if the shared library is missing this raises, and every compute call on
a machine without a CUDA device returns DQ_ERR_CUDA which is raised as RuntimeError.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DQ_LIB_PATH selects another build of the same library (kernel-tuning experiments); the default is the in-tree build
LIB_PATH = os.environ.get("DQ_LIB_PATH") or os.path.join(_HERE, "libdiffqcqp_b200.so")

# every symbol include/diffqcqp_b200.h declares (tests check the .so exports each one)
SYMBOLS = [
    "dq_version", "dq_build_arch", "dq_error_string", "dq_last_cuda_error", "dq_max_n",
    "dq_qp_forward", "dq_qp_backward", "dq_qcqp_forward", "dq_qcqp_backward",
    "dq_qp_solve_host", "dq_qcqp_solve_host", "dq_launch_count", "dq_host_release", "dq_qcqp_backward_ex", "dq_boxqp_forward", "dq_boxqp_backward", "dq_boxqp_backward_ex",
    "dq_set_forward_path", "dq_set_forward_tuning", "dq_selftest_inverse", "dq_qp_forward_ex", "dq_qp_backward_ex",
    "dq_qcqp_forward_ex", "dq_qcqp_backward_ex2",
]

_vp = ctypes.c_void_p
_i32, _i64, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_lib = None


class DiffQCQPError(RuntimeError):
    pass


def load():
    """Load libdiffqcqp_b200.so (building is the job of diffqcqp_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DiffQCQPError(
            f"{LIB_PATH} is missing: build it with `python -m diffqcqp_b200.build` "
            "(there is no CPU fallback for the CUDA path)")
    L = ctypes.CDLL(LIB_PATH)
    L.dq_version.restype = ctypes.c_int
    L.dq_build_arch.restype = ctypes.c_char_p
    L.dq_error_string.restype = ctypes.c_char_p
    L.dq_error_string.argtypes = [ctypes.c_int]
    L.dq_last_cuda_error.restype = ctypes.c_int
    L.dq_max_n.restype = ctypes.c_int
    L.dq_launch_count.restype = _i64
    L.dq_set_forward_path.restype = ctypes.c_int
    L.dq_set_forward_path.argtypes = [ctypes.c_int]
    L.dq_set_forward_tuning.restype = _i64
    L.dq_set_forward_tuning.argtypes = [_i32, _i64]
    L.dq_selftest_inverse.restype = ctypes.c_int
    L.dq_selftest_inverse.argtypes = [_vp, _i64, _vp, _vp]
    L.dq_qp_forward.restype = ctypes.c_int
    L.dq_qp_forward.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _f64, _f64, _i32, _i32, _vp]
    L.dq_qp_backward.restype = ctypes.c_int
    L.dq_qp_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]
    L.dq_qp_forward_ex.restype = ctypes.c_int
    L.dq_qp_forward_ex.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f64, _f64, _i32, _i32, _vp]
    L.dq_qp_backward_ex.restype = ctypes.c_int
    L.dq_qp_backward_ex.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]
    L.dq_qcqp_forward.restype = ctypes.c_int
    L.dq_qcqp_forward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f64, _f64, _i32, _i32, _vp]
    L.dq_qcqp_backward.restype = ctypes.c_int
    L.dq_qcqp_backward.argtypes = [_vp] * 10 + [_i64, _i32, _vp]
    L.dq_boxqp_forward.restype = ctypes.c_int
    L.dq_boxqp_forward.argtypes = [_vp] * 8 + [_i64, _i32, _f64, _f64, _i32, _i32, _vp]
    L.dq_boxqp_backward.restype = ctypes.c_int
    L.dq_boxqp_backward.argtypes = [_vp] * 10 + [_i64, _i32, _vp]
    L.dq_boxqp_backward_ex.restype = ctypes.c_int
    L.dq_boxqp_backward_ex.argtypes = [_vp] * 12 + [_i64, _i32, _vp]
    L.dq_qcqp_backward_ex.restype = ctypes.c_int
    L.dq_qcqp_backward_ex.argtypes = [_vp] * 12 + [_i64, _i32, _vp]
    L.dq_qcqp_forward_ex.restype = ctypes.c_int
    L.dq_qcqp_forward_ex.argtypes = [_vp] * 8 + [_i64, _i32, _f64, _f64, _i32, _i32, _vp]
    L.dq_qcqp_backward_ex2.restype = ctypes.c_int
    L.dq_qcqp_backward_ex2.argtypes = [_vp] * 13 + [_i64, _i32, _vp]
    L.dq_qp_solve_host.restype = ctypes.c_int
    L.dq_qp_solve_host.argtypes = [_vp] * 6 + [_i64, _i32, _f64, _f64, _i32, _i32]
    L.dq_qcqp_solve_host.restype = ctypes.c_int
    L.dq_qcqp_solve_host.argtypes = [_vp] * 10 + [_i64, _i32, _f64, _f64, _i32, _i32]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = load()
        msg = L.dq_error_string(rc).decode()
        extra = f" (cudaError {L.dq_last_cuda_error()})" if rc == 4 else ""
        raise DiffQCQPError(f"{what} failed: {msg}{extra}")


def launch_count() -> int:
    return int(load().dq_launch_count())


def set_forward_path(path: int) -> int:
    """Forward kernel selection (process-wide): 0 = automatic (N == 8 QP / Box QP: thread-per-problem kernel for batches of
    >= 65536 problems, persistent-CTA tile kernel below that), 1 = generic kernel only (e.g. for batches known to have dense
    P at N == 8), 2 = persistent tile kernel wherever it applies, 3 = thread-per-problem kernel wherever it applies.
    Returns the previous setting.  Results do not depend on it (bit-identical on all-diagonal / all-dense batches)."""
    return int(load().dq_set_forward_path(int(path)))


def set_forward_tuning(key: int, value: int) -> int:
    """dq_set_forward_tuning: key 0 = park threshold (iterations) of the thread-per-problem kernel, key 1 = smallest batch
    the automatic path gives to it.  Returns the previous value."""
    return int(load().dq_set_forward_tuning(int(key), int(value)))

"""diffqcqp_b200 -- B200 (sm_100a) batched differentiable ADMM QP/QCQP solver.

Drop-in for the QPFn2 / QCQPFn2 hot path of quentinll/diffqcqp (qcqp.py); see DESIGN.md.
"""
from .qcqp import (QPFn2, QCQPFn2, BoxQPFn2, SignedBoxQPFn2, qp_forward, qp_backward, qcqp_forward,  # noqa: F401
                   qcqp_backward, boxqp_forward, boxqp_backward, use_warm_start)
from ._lib import DiffQCQPError, launch_count, set_forward_path  # noqa: F401

__version__ = "0.1.0"

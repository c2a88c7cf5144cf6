"""`from qcqp import QPFn2, QCQPFn2` -- the reference's import line keeps working.

Like the reference module (qcqp.py:13) importing this sets torch's default dtype to double.
"""
import torch

torch.set_default_dtype(torch.double)

from diffqcqp_b200.qcqp import QPFn2, QCQPFn2, BoxQPFn2, SignedBoxQPFn2  # noqa: E402,F401

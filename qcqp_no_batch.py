"""`from qcqp_no_batch import QPFn2, QCQPFn2` -- the reference's unbatched layers (qcqp_no_batch.py:23-108).
Like the reference module, importing this sets torch's default dtype to double."""
import torch

torch.set_default_dtype(torch.double)

from diffqcqp_b200.legacy import QPFn2, QCQPFn2  # noqa: E402,F401

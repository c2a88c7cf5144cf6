"""`from diffqcqp import solveQP, solveQCQP, solveDerivativesQP, solveDerivativesQCQP` -- the import line of the
reference's qcqp.py:17 / qcqp_no_batch.py:16 keeps working (per-problem numpy API of pybindings.cpp:76-82, run on the
batched sm_100a kernels; see diffqcqp_b200/legacy.py)."""
from diffqcqp_b200.legacy import (solveQP, solveQCQP, solveDerivativesQP, solveDerivativesQCQP,  # noqa: F401
                                  solveBoxQP, solveSignedBoxQP, solveDerivativesBoxQP)

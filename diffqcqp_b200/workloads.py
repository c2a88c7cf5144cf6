"""Seeded synthetic batches for BASELINE.json's configs.  CPU fp64 tensors."""
from __future__ import annotations

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def qp_diag(B, N, seed=0, shift=0.0, signed_q=True):
    """cfg1/cfg2: P = diag_embed(rand(B,N) + shift), q = 2 rand - 1 (signed) or rand (README verbatim)."""
    g = _gen(seed)
    p = torch.rand(B, N, generator=g, dtype=torch.float64) + shift
    P = torch.diag_embed(p)
    q = torch.rand(B, N, 1, generator=g, dtype=torch.float64)
    if signed_q:
        q = 2 * q - 1
    grad = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
    return P, q, grad


def qp_dense(B, N, seed=0):
    """Dense QP: P = S S^T / N + 0.1 I, S = 2 rand - 1; q = 2 rand - 1."""
    g = _gen(seed)
    S = 2 * torch.rand(B, N, N, generator=g, dtype=torch.float64) - 1
    P = torch.bmm(S, S.transpose(1, 2)) / N + 0.1 * torch.eye(N, dtype=torch.float64)
    q = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
    grad = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
    return P, q, grad


def qcqp_dense(B, N, seed=0, diag=False):
    """cfg3/cfg5: dense P as qp_dense (or diagonal rand+0.1), l_n = 2 rand, mu = rand, nc = N/2."""
    g = _gen(seed)
    nc = N // 2
    if diag:
        P = torch.diag_embed(torch.rand(B, N, generator=g, dtype=torch.float64) + 0.1)
    else:
        S = 2 * torch.rand(B, N, N, generator=g, dtype=torch.float64) - 1
        P = torch.bmm(S, S.transpose(1, 2)) / N + 0.1 * torch.eye(N, dtype=torch.float64)
    q = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
    l_n = 2 * torch.rand(B, nc, 1, generator=g, dtype=torch.float64)
    mu = torch.rand(B, nc, 1, generator=g, dtype=torch.float64)
    grad = 2 * torch.rand(B, N, 1, generator=g, dtype=torch.float64) - 1
    return P, q, l_n, mu, grad


def qcqp_diag(B, N, seed=0):
    """Diagonal-P QCQP (P = diag_embed(rand + 0.1)), otherwise as qcqp_dense."""
    return qcqp_dense(B, N, seed=seed, diag=True)


# algorithmic bytes per solve: every input read once, every output written once.
# warm_start is never read by the kernels (dead in the reference, F2), so its 8N bytes are NOT counted.
def qp_bytes(N, fwd=True, bwd=True):
    b = 0
    if fwd:
        b += 8 * (N * N + 2 * N)          # P, q in; x out
    if bwd:
        b += 8 * (2 * N * N + 4 * N)      # P, q, x, grad_x in; grad_P, grad_q out
    return b


def qcqp_bytes(N, fwd=True, bwd=True):
    b = 0
    if fwd:
        b += 8 * (N * N + 3 * N)          # P, q, l_n, mu in (2*N/2 = N); x out
    if bwd:
        b += 8 * (2 * N * N + 6 * N)      # P, q, l_n, mu, x, grad_x in; grad_P, grad_q, grad_l_n, grad_mu out
    return b

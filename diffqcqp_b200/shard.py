"""Batch sharding of the QP/QCQP hot path across the GPUs of one box (one process per GPU).

Problems are independent (the reference solves them one by one in a Python loop, qcqp.py:29,149), so
the batch axis shards with NO collective inside forward or backward.  Collectives appear only at the
edges, when the caller holds the whole batch on one rank: scatter (P, q[, l_n, mu]) out, gather x* (and
gradients) back -- NCCL over NVLink on the GPU box, gloo in the CPU tests.

    lo, hi = shard_bounds(B, world, rank)            contiguous chunk owned by `rank`
    parts  = scatter_batch([P, q], B, src=0)         root holds (B, ...) tensors; everyone gets its chunk
    full   = gather_batch(x_local, B, dst=0)         inverse; returns the (B, ...) tensor on dst, None elsewhere
    x      = solve_sharded(fn, [P, q], B)            scatter -> fn(*local parts) -> gather

The solver itself is passed in (`fn`) so this module carries no compute path of its own: production passes
``QPFn2.apply``-style callables that run the sm_100a kernels; the gloo tests pass a CPU checker.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_sizes", "scatter_batch", "gather_batch", "solve_sharded"]


def shard_sizes(B: int, world: int) -> List[int]:
    """Chunk sizes: the first B % world ranks get one extra problem (ragged batches are fine)."""
    if B < 0 or world < 1:
        raise ValueError(f"bad shard request B={B} world={world}")
    base, extra = divmod(B, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    sizes = shard_sizes(B, world)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def scatter_batch(tensors: Optional[Sequence[torch.Tensor]], B: int, src: int = 0, device=None,
                  trailing: Optional[Sequence[Tuple[int, ...]]] = None, dtype=torch.float64, group=None):
    """Scatter (B, ...) tensors held by `src` into per-rank contiguous chunks along dim 0.

    Non-root ranks pass ``tensors=None`` and describe what they expect with ``trailing`` (the shapes after
    the batch dim).  Uses point-to-point send/recv (NCCL has no scatter primitive; grouped send/recv is the
    native form) so chunk sizes may be ragged.
    """
    world, rank = _world(group)
    if world == 1:
        return [t if device is None else t.to(device) for t in tensors]
    sizes = shard_sizes(B, world)
    if rank == src:
        if trailing is None:
            trailing = [tuple(t.shape[1:]) for t in tensors]
        device = device if device is not None else tensors[0].device
    elif trailing is None:
        raise ValueError("non-root ranks must pass `trailing` shapes")
    outs = [torch.empty((sizes[rank],) + tuple(tr), dtype=dtype, device=device) for tr in trailing]
    ops = []
    if rank == src:
        offs = 0
        for r in range(world):
            for i, t in enumerate(tensors):
                chunk = t[offs:offs + sizes[r]].contiguous()
                if r == src:
                    outs[i].copy_(chunk)
                elif sizes[r]:
                    ops.append(dist.P2POp(dist.isend, chunk, r, group))
            offs += sizes[r]
    elif sizes[rank]:
        for o in outs:
            ops.append(dist.P2POp(dist.irecv, o, src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return outs


def gather_batch(local: torch.Tensor, B: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of scatter_batch for one tensor: returns the (B, ...) tensor on `dst`, None elsewhere."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = shard_sizes(B, world)
    if local.size(0) != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.size(0)} problems, expected {sizes[rank]}")
    ops, full = [], None
    if rank == dst:
        full = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        offs = 0
        for r in range(world):
            view = full[offs:offs + sizes[r]]
            if r == dst:
                view.copy_(local)
            elif sizes[r]:
                ops.append(dist.P2POp(dist.irecv, view, r, group))
            offs += sizes[r]
    elif sizes[rank]:
        ops.append(dist.P2POp(dist.isend, local.contiguous(), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


def solve_sharded(fn: Callable[..., torch.Tensor], tensors: Optional[Sequence[torch.Tensor]], B: int,
                  src: int = 0, device=None, trailing=None, group=None) -> Optional[torch.Tensor]:
    """scatter -> fn(*local) -> gather.  `fn` is the per-rank solve (e.g. a QPFn2.apply closure)."""
    parts = scatter_batch(tensors, B, src=src, device=device, trailing=trailing, group=group)
    x_local = fn(*parts)
    return gather_batch(x_local, B, dst=src, group=group)

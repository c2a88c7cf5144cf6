"""Batch sharding of the QP/QCQP hot path across the GPUs of one box (one process per GPU).

Problems are independent (the reference solves them one by one in a Python loop, qcqp.py:29,149), so
the batch axis shards with NO collective inside forward or backward.  Collectives appear only at the
edges, when the caller holds the whole batch on one rank: scatter (P, q[, l_n, mu]) out, gather x* (and
gradients) back -- NCCL over NVLink on the GPU box, gloo in the CPU tests.

    lo, hi = shard_bounds(B, world, rank)            contiguous chunk owned by `rank`
    parts  = scatter_batch([P, q], B, src=0)         root holds (B, ...) tensors; everyone gets its chunk
    full   = gather_batch(x_local, B, dst=0)         inverse; returns the (B, ...) tensor on dst, None elsewhere
    x      = solve_sharded(fn, [P, q], B)            scatter -> fn(*local parts) -> gather
    outs   = solve_sharded_pipelined(fn, [P, q], B, chunks=4)
                                                     the same in `chunks` pieces per rank: piece c is solved while piece
                                                     c+1 is still on the wire and piece c-1's results travel back

The solver itself is passed in (`fn`) so this module carries no compute path of its own: production passes
``QPFn2.apply``-style callables that run the sm_100a kernels; the gloo tests pass a CPU checker.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_sizes", "scatter_batch", "gather_batch", "solve_sharded", "solve_sharded_pipelined"]


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


def solve_sharded_pipelined(fn: Callable[..., Sequence[torch.Tensor]], tensors: Optional[Sequence[torch.Tensor]], B: int,
                            chunks: int = 4, src: int = 0, device=None, trailing=None, dtype=torch.float64, group=None
                            ) -> Optional[List[torch.Tensor]]:
    """scatter -> fn -> gather in `chunks` pieces per rank, software-pipelined.

    Rank r's shard is cut into `chunks` contiguous pieces.  Piece c of every rank is scattered as one grouped
    send/recv; ``fn(*piece, lo, hi)`` (the per-rank solve of local problems [lo, hi) of the shard; returns a sequence of
    (hi - lo, ...) tensors) runs as soon as that piece has arrived, while the group of piece c+1 is already in flight,
    and the results of piece c go back to `src` as another grouped send/recv.  The order of the groups is the same on
    every rank (S0 S1 G0 S2 G1 ... G_last), which is all NCCL needs; on the GPU box they run on the communicator's own
    stream, so the root's egress -- the floor of any scatter schedule -- overlaps the solves instead of preceding them.
    Set TORCH_NCCL_HIGH_PRIORITY=1 before init_process_group: the root's send kernels then take SM slots ahead of the
    solve's queued CTAs (8 x B200, 2,097,152 N=16 QCQPs: 12.6 ms unpipelined, 11.8 pipelined, 9.3 pipelined with priority).
    `fn` is also called for an empty piece (hi == lo) and must return empty tensors of the right trailing shapes.
    Returns the list of gathered (B, ...) outputs on `src`, None elsewhere.  Results are those of solve_sharded bit for
    bit (problems are independent; only the launch granularity changes).
    """
    world, rank = _world(group)
    if chunks < 1:
        raise ValueError("chunks must be >= 1")
    sizes = shard_sizes(B, world)
    if rank == src:
        if trailing is None:
            trailing = [tuple(t.shape[1:]) for t in tensors]
        device = device if device is not None else tensors[0].device
    elif trailing is None:
        raise ValueError("non-root ranks must pass `trailing` shapes")
    n_loc = sizes[rank]
    starts = [sum(sizes[:r]) for r in range(world)]
    pieces = [shard_sizes(sizes[r], chunks) for r in range(world)]          # pieces[r][c] = problems of rank r in piece c
    pstart = [[sum(pieces[r][:c]) for c in range(chunks)] for r in range(world)]
    if world == 1:
        loc = [t if device is None else t.to(device) for t in tensors]
    else:
        loc = [torch.empty((n_loc,) + tuple(tr), dtype=dtype, device=device) for tr in trailing]

    def scatter_piece(c):
        ops = []
        if world == 1:
            return ops
        if rank == src:
            for r in range(world):
                n, o = pieces[r][c], starts[r] + pstart[r][c]
                if not n:
                    continue
                for i, t in enumerate(tensors):
                    if r == src:
                        loc[i][pstart[r][c]:pstart[r][c] + n].copy_(t[o:o + n])
                    else:
                        ops.append(dist.P2POp(dist.isend, t[o:o + n], r, group))
        elif pieces[rank][c]:
            lo = pstart[rank][c]
            for o in loc:
                ops.append(dist.P2POp(dist.irecv, o[lo:lo + pieces[rank][c]], src, group))
        return dist.batch_isend_irecv(ops) if ops else []

    full: Optional[List[torch.Tensor]] = None

    def gather_piece(c, res):
        nonlocal full
        ops = []
        if rank == src:
            if full is None:
                full = [torch.empty((B,) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device) for o in res]
            for r in range(world):
                n, o = pieces[r][c], starts[r] + pstart[r][c]
                if not n:
                    continue
                for i, f in enumerate(full):
                    if r == src:
                        f[o:o + n].copy_(res[i])
                    else:
                        ops.append(dist.P2POp(dist.irecv, f[o:o + n], r, group))
        elif pieces[rank][c]:
            for o in res:
                ops.append(dist.P2POp(dist.isend, o.contiguous(), src, group))
        return dist.batch_isend_irecv(ops) if ops else []

    works = [scatter_piece(0)]
    pending = []
    for c in range(chunks):
        if c + 1 < chunks:
            works.append(scatter_piece(c + 1))      # piece c+1 goes on the wire before piece c is solved
        for w in works[c]:
            w.wait()                                # (NCCL: the current stream waits, the host does not)
        lo, n = pstart[rank][c], pieces[rank][c]
        res = list(fn(*[t[lo:lo + n] for t in loc], lo, lo + n))
        pending.append(gather_piece(c, res))
    for ws in pending:
        for w in ws:
            w.wait()
    return full if rank == src else None

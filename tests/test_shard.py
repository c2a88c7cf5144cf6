"""world_size-2 gloo tests (CPU) of the batch-sharding plumbing in diffqcqp_b200/shard.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffqcqp_b200 import shard


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 8, 65536, 2097152 + 3):
        for world in (1, 2, 3, 8):
            sizes = shard.shard_sizes(B, world)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
            prev = 0
            for r in range(world):
                lo, hi = shard.shard_bounds(B, world, r)
                assert lo == prev and hi - lo == sizes[r]
                prev = hi
            assert prev == B
    with pytest.raises(ValueError):
        shard.shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc  # the CPU checker stands in for the per-rank solve (tests only)
        from diffqcqp_b200 import workloads as wl
        P = q = None
        if rank == 0:
            P, q, _ = wl.qp_dense(B, N, seed=3)

        def local_solve(Pl, ql):
            assert Pl.shape[0] == shard.shard_sizes(B, world)[rank]
            return torch.from_numpy(orc.qp_forward(Pl.numpy(), ql.numpy(), None, 1e-7, 1000))

        x = shard.solve_sharded(local_solve, [P, q] if rank == 0 else None, B, src=0, device=torch.device("cpu"),
                                trailing=[(N, N), (N, 1)])
        if rank == 0:
            ref = orc.qp_forward(P.numpy(), q.numpy(), None, 1e-7, 1000)
            np.save(os.path.join(out_dir, "ok.npy"), np.array([float(np.abs(x.numpy() - ref).max()), x.shape[0]]))
        else:
            assert x is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 37])
def test_scatter_solve_gather_world2(tmp_path, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, 8, str(tmp_path)), nprocs=2, join=True)
    err, n = np.load(tmp_path / "ok.npy")
    assert n == B and err == 0.0  # sharding must not change any bit of any problem's result


def _worker_pipelined(rank, world, port, B, N, chunks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc  # the CPU checker stands in for the per-rank solve (tests only)
        from diffqcqp_b200 import workloads as wl
        P = q = None
        if rank == 0:
            P, q, _ = wl.qp_dense(B, N, seed=4)
        seen = []

        def local_solve(Pl, ql, lo, hi):
            assert Pl.shape[0] == hi - lo
            seen.append((lo, hi))
            x = torch.from_numpy(orc.qp_forward(Pl.numpy(), ql.numpy(), None, 1e-7, 1000)) if hi > lo else torch.empty((0, N, 1), dtype=torch.float64)
            return x, 2.0 * x[:, :1]          # two outputs with different trailing shapes

        outs = shard.solve_sharded_pipelined(local_solve, [P, q] if rank == 0 else None, B, chunks=chunks, src=0,
                                             device=torch.device("cpu"), trailing=[(N, N), (N, 1)])
        n_loc = shard.shard_sizes(B, world)[rank]
        assert len(seen) == chunks and seen[0][0] == 0 and seen[-1][1] == n_loc
        assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
        if rank == 0:
            ref = orc.qp_forward(P.numpy(), q.numpy(), None, 1e-7, 1000)
            np.save(os.path.join(out_dir, "okp.npy"), np.array([float(np.abs(outs[0].numpy() - ref).max()),
                                                                 float(np.abs(outs[1].numpy() - 2.0 * ref[:, :1]).max()), outs[0].shape[0]]))
        else:
            assert outs is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,chunks", [(37, 3), (10, 4), (5, 4)])
def test_pipelined_scatter_solve_gather_world2(tmp_path, B, chunks):
    """Ragged shards, ragged pieces (and empty pieces when a shard has fewer problems than pieces): same bits as one solve."""
    port = _free_port()
    mp.spawn(_worker_pipelined, args=(2, port, B, 8, chunks, str(tmp_path)), nprocs=2, join=True)
    e0, e1, n = np.load(tmp_path / "okp.npy")
    assert n == B and e0 == 0.0 and e1 == 0.0

"""World-size-2 gloo test of the multi-GPU host logic: slab bounds, per-rank fit stand-in, gather of the maps."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luci_b200 import dist as ldist


def test_shard_bounds_cover_everything_once():
    for n in (2048, 2047, 10, 7):
        for world in (1, 2, 4, 8):
            for align in (1, 2):
                edges = [ldist.shard_bounds(n, r, world, align) for r in range(world)]
                assert edges[0][0] == 0 and edges[-1][1] == n
                for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                    assert b == c and a <= b and b % align == 0
    assert [ldist.split_even(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    r, w, _ = ldist.init_process_group('gloo')
    nx, ny, K = 13, 5, 7
    lo, hi = ldist.shard_bounds(nx, r, w)
    # stand-in for the per-rank GPU fit: row s of the global (x-major) spaxel list carries its own index
    s = torch.arange(lo * ny, hi * ny, dtype=torch.float32)
    local = s[:, None] * 10 + torch.arange(K, dtype=torch.float32)[None, :]
    got = ldist.gather_rows(local, dst=0)
    # same gather with the shard sizes known up front and a preallocated destination (what bench.py does)
    counts = [(b - a) * ny for a, b in (ldist.shard_bounds(nx, q, w) for q in range(w))]
    out = torch.full((nx * ny, K), -1.0) if r == 0 else None
    got2 = ldist.gather_rows(local, dst=0, counts=counts, out=out)
    if r == 0:
        assert got2 is out and torch.equal(got, got2)
    else:
        assert got2 is None
    tmax = ldist.max_over_ranks(1.0 + r)
    tsum = ldist.sum_over_ranks(float(hi - lo))
    if r == 0:
        q.put((got.numpy(), tmax, tsum))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax, tsum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nx, ny, K = 13, 5, 7
    s = np.arange(nx * ny, dtype=np.float32)
    np.testing.assert_array_equal(got, s[:, None] * 10 + np.arange(K, dtype=np.float32)[None, :])
    assert tmax == 2.0 and tsum == 13.0

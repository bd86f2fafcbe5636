"""world_size-2 (and 3) gloo tests of the slab partition and halo exchange (CPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ncols, col, q):
    sys.path.insert(0, ROOT)
    from stefanproblem_b200.distributed import HaloExchange, allreduce_interface_sums, slab_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = slab_range(ncols, world, rank)
        # global vector: entry (column c, k) = 1000 c + k
        x = (1000.0 * torch.arange(b, e, dtype=torch.float64)[:, None]
             + torch.arange(col, dtype=torch.float64)[None, :]).reshape(-1).contiguous()
        halo = HaloExchange(col, rank, world, x)
        lo, hi = halo.exchange(x)
        ok = True
        if rank > 0:
            ok &= bool(torch.equal(lo, 1000.0 * (b - 1) + torch.arange(col, dtype=torch.float64)))
        else:
            ok &= lo is None
        if rank < world - 1:
            ok &= bool(torch.equal(hi, 1000.0 * e + torch.arange(col, dtype=torch.float64)))
        else:
            ok &= hi is None
        s = allreduce_interface_sums(torch.tensor([float(rank + 1), 2.0, 0.0, 1.0], dtype=torch.float64))
        ok &= bool(torch.equal(s, torch.tensor([world * (world + 1) / 2.0, 2.0 * world, 0.0, float(world)],
                                               dtype=torch.float64)))
        q.put((rank, b, e, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ncols", [(2, 10), (2, 7), (3, 10)])
def test_halo_exchange_and_partition(world, ncols):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + ncols
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncols, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[3] for r in res)
    # slabs tile [0, ncols) without gaps or overlap
    assert res[0][1] == 0 and res[-1][2] == ncols
    for a, b in zip(res[:-1], res[1:]):
        assert a[2] == b[1]
    sizes = [r[2] - r[1] for r in res]
    assert max(sizes) - min(sizes) <= 1


def test_slab_range_properties():
    from stefanproblem_b200.distributed import slab_range
    for n in (1, 5, 5000, 5001):
        for w in (1, 2, 4, 8):
            r = [slab_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))


def _fd_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from oracle import fd as ofd
    from stefanproblem_b200.distributed import HaloExchange, slab_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # an interface two nodes beyond the first cut: the one-sided rows reach across the halo
        cut = slab_range(n, world, 1)[0]
        ph = np.where((np.arange(n) >= n // 5) & (np.arange(n) < cut + 2), 1, 2).astype(np.int8)
        A = ofd.laplace_matrix(ph, 1.0 / (n - 1)).tocsr()
        ug = np.random.default_rng(4).uniform(-1, 1, n)
        b, e = slab_range(n, world, rank)
        u = torch.from_numpy(ug[b:e].copy())
        lo, hi = HaloExchange(3, rank, world, u).exchange(u)
        c0, c1 = (b - 3 if rank > 0 else b), (e + 3 if rank < world - 1 else e)
        padded = np.concatenate([t.numpy() for t in (lo, u, hi) if t is not None])
        rows = A[b:e]
        inside = (rows.indices >= c0) & (rows.indices < c1)   # every row of the node range stays within the halo
        ok = bool(inside.all()) and np.array_equal(rows[:, c0:c1] @ padded, (A @ ug)[b:e])
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fd_three_node_halos_reproduce_the_global_rows(world):
    """The node-range partition of `DistributedFDOperator`: three halo nodes per side carry every
    row of finite-difference.jl:34-85 (the one-sided interface rows included) on world_size 2 and 3."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fd_worker, args=(r, world, 29680 + world, 101, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_balanced_bounds_properties():
    """Cost-balanced column slabs (distributed.balanced_bounds): bounds are monotone, cover [0, ncols], equal costs
    give equal shares, and a rank that was slower per column gets fewer columns."""
    from stefanproblem_b200.distributed import balanced_bounds
    b = balanced_bounds(5000, [40e-6] * 8, [625] * 8, fixed_seconds=8e-6)
    assert b == [0, 625, 1250, 1875, 2500, 3125, 3750, 4375, 5000]
    secs = [39.6e-6, 42e-6, 41e-6, 40.4e-6, 40.4e-6, 41e-6, 42e-6, 39.6e-6]
    b = balanced_bounds(5000, secs, [625] * 8, fixed_seconds=8e-6)
    w = [b[r + 1] - b[r] for r in range(8)]
    assert b[0] == 0 and b[-1] == 5000 and all(x > 0 for x in w)
    assert w[1] < w[3] < w[0] and w[1] == w[6] and w[0] == w[7]
    # the model's predicted times are equal to within one column
    pred = [8e-6 + (secs[r] - 8e-6) / 625 * w[r] for r in range(8)]
    assert max(pred) - min(pred) < 1.5 * max(secs) / 625
    assert balanced_bounds(13, [1.0, 1.0, 5.0], [4, 4, 5])[-1] == 13

"""world_size = 2 on CPU (gloo): the host-side logic of the row-sharded fit - row ownership of the drawn
centroids, the exact SUM reconstruction, partial sums/counts, tie offsets, rank-consistent control flow.
The device kernels are not involved (they are covered single-GPU, incl. an emulation of this protocol,
in tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generative_ranking_recommender_b200 import engine, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = engine.ShardGroup()
        assert g.active and g.world == world and g.rank == rank
        rng = np.random.default_rng(0)
        n, dim, k = 1003, 16, 8
        x = rng.standard_normal((n, dim)).astype(np.float32)
        bounds = [0, 400, n]
        lo, hi = bounds[rank], bounds[rank + 1]
        xl = torch.from_numpy(x[lo:hi])

        # global row bookkeeping
        sizes = g.all_gather(torch.tensor([hi - lo], dtype=torch.int64)).view(-1).tolist()
        n_global, row0 = sharding.global_rows(sizes, rank)
        assert (n_global, row0) == (n, lo)

        # centroid init: same draw everywhere (same seed), owners contribute, SUM reconstructs exactly
        np.random.seed(5)
        idx = np.random.choice(n_global, k, replace=False)
        pos, loc = sharding.owned_rows(idx, row0, hi - lo)
        out = torch.zeros((k, dim))
        sharding.scatter_owned(out, pos, xl[torch.from_numpy(loc)])
        g.all_reduce(out, "sum")
        assert torch.equal(out, torch.from_numpy(x[idx]))

        # host-side draws of a sharded fit: ranks seeded DIFFERENTLY still use rank 0's rows
        np.random.seed(100 + rank)
        mine = np.random.choice(n_global, k, replace=False)
        used = g.broadcast_ints(mine)
        np.random.seed(100)
        assert np.array_equal(used, np.random.choice(n_global, k, replace=False))

        # with replacement (K > N): duplicates survive
        idx2 = np.array([3, 3, 999, 400, 3])
        pos2, loc2 = sharding.owned_rows(idx2, row0, hi - lo)
        out2 = torch.zeros((5, dim))
        sharding.scatter_owned(out2, pos2, xl[torch.from_numpy(loc2)])
        g.all_reduce(out2, "sum")
        assert torch.equal(out2, torch.from_numpy(x[idx2]))

        # centroid update: per-rank partial sums / counts -> all_reduce == global
        a = rng.integers(0, k, n)
        al = a[lo:hi]
        sums = torch.zeros((k, dim), dtype=torch.float64)
        sums.index_add_(0, torch.from_numpy(al), xl.double())
        counts = torch.bincount(torch.from_numpy(al), minlength=k)
        g.all_reduce(sums, "sum")
        g.all_reduce(counts, "sum")
        ref = np.zeros((k, dim))
        np.add.at(ref, a, x.astype(np.float64))
        assert np.allclose(sums.numpy(), ref) and np.array_equal(counts.numpy(), np.bincount(a, minlength=k))

        # eps needs the global extrema
        mm = torch.tensor([100 + rank, 50 - rank], dtype=torch.int32)
        g.all_reduce(mm[0:1], "max")
        g.all_reduce(mm[1:2], "min")
        assert mm.tolist() == [101, 49]

        # tie offsets: rank-major job order
        totals = g.all_gather(torch.tensor([rank + 1, 2 * rank, 7], dtype=torch.int32))
        offs = sharding.rank_tie_offsets(totals, rank)
        assert offs.tolist() == ([0, 0, 0] if rank == 0 else [1, 0, 7])

        # control flow must not diverge: every rank derives stop decisions from reduced values only
        shift = torch.tensor([0.25 * (rank + 1)], dtype=torch.float64)
        g.all_reduce(shift, "max")
        q.put((rank, float(shift.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, 0.5), (1, 0.5)]


def test_no_shard_is_identity():
    g = engine.no_shard()
    t = torch.arange(4)
    assert not g.active and g.world == 1 and torch.equal(g.all_reduce(t.clone()), t)
    assert g.all_gather(t).shape == (1, 4)
    assert sharding.rank_tie_offsets(torch.ones((1, 3), dtype=torch.int32), 0).tolist() == [0, 0, 0]

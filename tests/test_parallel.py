"""world_size-2 gloo tests (CPU) of the batch-sharding host logic used for multi-GPU sampling."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sr3_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 2, 5, 16, 17):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cond = torch.rand(n, 3, 4, 4, generator=g)
        x_T = torch.randn(n, 3, 4, 4, generator=g)

        def fake_sampler(c, xt, first):          # deterministic function of (global index, inputs): stands in for the GPU loop
            idx = torch.arange(first, first + xt.shape[0], dtype=torch.float32).view(-1, 1, 1, 1)
            return c * 2 + xt + idx

        out = parallel.sharded_sample(fake_sampler, cond, x_T)
        ref = cond * 2 + x_T + torch.arange(n, dtype=torch.float32).view(-1, 1, 1, 1)
        ret[rank] = bool(torch.equal(out, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [16, 5])
def test_sharded_sample_two_ranks_gloo(n):
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


class _FakeEngine:
    def __init__(self, n):
        self.n = n

    def p_sample_loop(self, c, xt, noises, seed, first, want_snapshots=True):
        assert c.shape[0] == xt.shape[0] == self.n and noises is None and not want_snapshots
        idx = torch.arange(first, first + self.n, dtype=torch.float32).view(-1, 1, 1, 1)
        return c * 2 + xt + idx + float(seed), None


class _FakeNet:
    """Stands in for GaussianDiffusion on a CPU box: sharded_super_resolution only needs `.betas.device` and `._engine(batch)`."""
    betas = torch.zeros(1)

    def _engine(self, n):
        return _FakeEngine(n)


def _worker_sr(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cond = torch.rand(n, 3, 4, 4, generator=g)
        x_T = torch.randn(n, 3, 4, 4, generator=g)
        out = parallel.sharded_super_resolution(_FakeNet(), cond, x_T=x_T, seed=5)
        ref = cond * 2 + x_T + torch.arange(n, dtype=torch.float32).view(-1, 1, 1, 1) + 5.0
        ret[rank] = bool(torch.equal(out, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [16, 3])
def test_sharded_super_resolution_two_ranks_gloo(n):
    """The public multi-GPU entry point (shard by image -> per-rank engine loop keyed by the global sample index -> all-gather)."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_sr, args=(world, port, n, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


# ---- training: bucketed gradient all-reduce (training row) ------------------------------------------------------
def test_plan_buckets_orders_by_backward_completion():
    # 4 backward blocks (run 3, 2, 1, 0); parameter 6 belongs to no block (FiLM / noise MLP: final after the whole backward)
    blocks = [[0, 1], [2], [3, 4], [5]]
    numels = [10, 10, 10, 10, 10, 30, 7]
    b = parallel.plan_buckets(blocks, numels, 25)
    assert b == [(3, [5]), (1, [3, 4, 2]), (0, [0, 1]), (-1, [6])]
    assert sorted(sum((ix for _, ix in b), [])) == list(range(7))
    # one huge bucket: everything except the late parameters is reduced when block 0 has run
    assert parallel.plan_buckets(blocks, numels, 10 ** 9) == [(0, [5, 3, 4, 2, 0, 1]), (-1, [6])]


def _bucket_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = [(4, 3), (5,), (2, 2, 2), (7,), (3, 3), (6,), (2,)]
        params = [torch.zeros(s) for s in shapes]
        blocks = [[0, 1], [2], [3, 4], [5]]
        gb = parallel.GradientBuckets(params, blocks, 12)
        gb.begin()
        for i in range(len(blocks) - 1, -1, -1):          # the "backward": block i writes its parameters' gradients, then reports
            for pi in blocks[i]:
                gb.views[pi].copy_(torch.full(shapes[pi], float((rank + 1) * (pi + 1))))
            gb.ready(i)
        gb.views[6].copy_(torch.full(shapes[6], float((rank + 1) * 7)))
        gb.finish()
        ok = gb.n_reduced == len(gb.slices)
        for pi in range(7):
            ok = ok and bool(torch.equal(gb.views[pi], torch.full(shapes[pi], float(3 * (pi + 1)))))     # (1 + 2) * (pi + 1)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_two_ranks_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_bucket_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

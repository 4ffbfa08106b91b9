"""Multi-GPU sampling: the reverse trajectories of different images are independent (GroupNorm and attention are per
sample), so the batch is sharded across ranks with NO per-step communication and the finished images are collected with one
all-gather.  The reference only offers nn.DataParallel for training and samples on GPU 0
(model/model.py:60-78); this is the B200-native replacement for that path: one process per GPU, NCCL over NVLink.

Noise comes from Philox streams keyed by the GLOBAL sample index, so the result does not depend on the number of ranks.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of the batch owned by `rank` (first `global_batch % world_size` ranks get one more)."""
    if global_batch < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local: torch.Tensor, global_batch: int, group=None) -> torch.Tensor:
    """All-gather row shards of possibly unequal length into the [global_batch, ...] tensor (same on every rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(global_batch, world, r)[1] - shard_bounds(global_batch, world, r)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = torch.empty((global_batch,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def sharded_sample(sample_fn: Callable[[Optional[torch.Tensor], torch.Tensor, int], torch.Tensor], cond: Optional[torch.Tensor],
                   x_T: torch.Tensor, group=None) -> torch.Tensor:
    """Run `sample_fn(cond_shard, x_T_shard, first_global_index)` on this rank's slice and all-gather the finished images.

    `cond` / `x_T` are the GLOBAL tensors (every rank holds or can build them, e.g. from a shared seed)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = x_T.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = sample_fn(None if cond is None else cond[lo:hi], x_T[lo:hi], lo)
    if world == 1:
        return local
    return gather_shards(local, n, group)


def sharded_super_resolution(netG, x_in: torch.Tensor, x_T: Optional[torch.Tensor] = None, seed: int = 0, group=None) -> torch.Tensor:
    """Batch-sharded `GaussianDiffusion.super_resolution` (reference: model/sr3_modules/diffusion.py:176-210 run on GPU 0 only,
    model/model.py:60-78): returns the [B,3,H,W] finished images x_0 on every rank.

    `x_in` / `x_T` are the GLOBAL (host or device) tensors; each rank copies and samples only its slice; the Philox noise streams are keyed
    by the global sample index (`first_index`), so the images do not depend on the number of ranks."""
    dev = netG.betas.device
    if x_T is None:
        g = torch.Generator().manual_seed(seed)
        x_T = torch.randn(tuple(x_in.shape), generator=g)

    def fn(c, xt, first):
        if c.shape[0] == 0:
            return torch.empty((0,) + tuple(xt.shape[1:]), device=dev)
        eng = netG._engine(c.shape[0])
        final, _ = eng.p_sample_loop(c.to(dev, non_blocking=True), xt.to(dev, non_blocking=True), None, seed, first, want_snapshots=False)
        return final

    return sharded_sample(fn, x_in, x_T, group)

"""Multi-GPU sharding of the fitting batch.

Bodies are independent optimisation problems given a read-only scene and model, so the batch
is partitioned contiguously over ranks with NO data-path communication; every rank keeps its
own copy of the model constants and of the scene(s) it needs.  One all-gather of the fitted
[B/G,72] vectors ends the job (153.6 kB at B=512).  Backend: NCCL on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int):
    """Contiguous, balanced partition: the first (total % world) ranks get one extra body."""
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(x, world_size: int, rank: int):
    s, e = shard_bounds(x.shape[0], world_size, rank)
    return x[s:e]


def gather_rows(local, total: int, group=None):
    """All-gather variable-length row shards back into [total, ...] (same order as shard_rows)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def fit_sharded(make_fitting_op, xh_all, cam_ext_all, num_iter=None, group=None):
    """Each rank fits its contiguous shard with its own FittingOP, then all ranks receive the
    full fitted [B,72].  `make_fitting_op(local_batch)` builds the rank-local operator."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    xh = shard_rows(xh_all, world, rank)
    cam = cam_ext_all if cam_ext_all.shape[0] == 1 else shard_rows(cam_ext_all, world, rank)
    op = make_fitting_op(xh.shape[0])
    out = op.fit(xh.to(op.device), cam.to(op.device), num_iter) if xh.shape[0] else xh.new_zeros((0, 72)).to(op.device)
    return gather_rows(out, xh_all.shape[0], group)

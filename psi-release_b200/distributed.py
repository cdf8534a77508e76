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


def plan_scene_shards(bodies_per_scene, world_size):
    """BASELINE config 3 (512 bodies over 4 scenes on 8 GPUs): bodies are laid out scene after
    scene and cut into `world_size` contiguous, balanced runs, so a rank touches the fewest
    scenes possible (whole scenes / halves of scenes when the sizes divide).  Returns, per rank,
    a list of (scene_id, first_body_in_scene, last_body_in_scene_exclusive)."""
    total = int(sum(bodies_per_scene))
    starts = [0]
    for n in bodies_per_scene:
        starts.append(starts[-1] + int(n))
    plan = []
    for r in range(world_size):
        lo, hi = shard_bounds(total, world_size, r)
        items = []
        for s, n in enumerate(bodies_per_scene):
            a, b = max(lo, starts[s]), min(hi, starts[s + 1])
            if a < b:
                items.append((s, a - starts[s], b - starts[s]))
        plan.append(items)
    return plan


def fit_scenes(make_fitting_op, xh_per_scene, cam_ext_per_scene, num_iter=None, group=None):
    """Fit several scenes' bodies over all ranks.  `make_fitting_op(scene_id, batch)` builds the
    rank-local operator for one scene (model + that scene's SDF / points / index on this GPU).
    Every rank returns the fitted [sum(B_s), 72] in scene order (one all-gather at the end)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    counts = [int(x.shape[0]) for x in xh_per_scene]
    mine = plan_scene_shards(counts, world)[rank]
    outs, dev = [], None
    for s, a, b in mine:
        op = make_fitting_op(s, b - a)
        dev = op.device
        cam = cam_ext_per_scene[s]
        cam = cam if cam.shape[0] == 1 else cam[a:b]
        outs.append(op.fit(xh_per_scene[s][a:b].to(dev), cam.to(dev), num_iter))
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    local = torch.cat(outs, dim=0) if outs else torch.zeros(0, 72, device=dev)
    return gather_rows(local, sum(counts), group)

"""Generate -> decode -> fit -> score over several rooms (BASELINE config 5: MP3D-R style, 7 rooms x 128
samples), sharded over the ranks of a torch.distributed job.

The reference does this with three scripts per room: test_habitat_s2.py:215-229 samples bodies from the
CVAE and writes one pickle per body, fitting_habitat.py:201-218 refines every pickle (one process, one
body at a time) and utils_eval_collision_habitat.py:91-140 scores the results.  Here a room's samples
form ONE batch on one GPU (or a slice of it: `distributed.plan_scene_shards`), fitted by the fused
loop, scored with the same kernels, and written in the reference's per-body pickle format.  The CVAE
itself stays stock torch (out of scope): `generate` is any callable
`(room_index, n) -> xh [n,72]`; the synthetic stand-in samples the CVAE's output distribution shape
(`synthetic.make_body_params`).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import evaluate, io
from .distributed import gather_rows, plan_scene_shards
from .fitting import FittingOP


def fit_rooms(rooms, generate, samples_per_room, fittingconfig, lossconfig, out_dir=None, group=None):
    """rooms: list of scene objects (.sdf, .grid_min, .grid_max, .points, .cam_ext); generate(room, n) ->
    numpy/torch [n,72]; fittingconfig: FittingOP config WITHOUT scene / batch_size.
    Returns dict(fitted [R*n,72], non_collision [R*n], contact [R*n]) in room order on every rank;
    with out_dir, rank-local results are also written as <out_dir>/room_<r>/body_gen_<i>.pkl."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    counts = [int(samples_per_room)] * len(rooms)
    outs, ncs, cts, dev = [], [], [], None
    if "body_mesh_model" not in fittingconfig:
        # one device copy of the SMPL-X constants for every room this rank touches
        from . import body_model
        fittingconfig = dict(fittingconfig, body_mesh_model=body_model.create(
            fittingconfig.get("human_model_path"), model_type="smplx", gender="neutral", ext="npz", num_pca_comps=12,
            batch_size=int(samples_per_room), model_data=fittingconfig.get("model_data")))
    for r, a, b in plan_scene_shards(counts, world)[rank]:
        scene = rooms[r]
        xh = torch.as_tensor(generate(r, counts[r]), dtype=torch.float32)[a:b]     # same samples whatever the sharding
        op = FittingOP(dict(fittingconfig, scene=scene, batch_size=b - a), lossconfig)
        dev = op.device
        cam = torch.as_tensor(scene.cam_ext, dtype=torch.float32, device=dev).unsqueeze(0)
        fitted = op.fit(xh.to(dev), cam)
        nc, ct = evaluate.collision_scores(op, fitted, cam)
        outs.append(fitted); ncs.append(nc); cts.append(ct)
        if out_dir is not None:
            d = os.path.join(out_dir, "room_%03d" % r)
            os.makedirs(d, exist_ok=True)
            fh = fitted.cpu().numpy()
            for i in range(a, b):
                io.write_body_pickle(os.path.join(d, "body_gen_%06d.pkl" % i), fh[i - a], scene.cam_ext[None],
                                     getattr(scene, "cam_int", np.eye(3, dtype=np.float32)[None]))
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    cat = lambda xs, w: torch.cat(xs, 0) if xs else torch.zeros((0,) + w, device=dev)
    total = sum(counts)
    return {"fitted": gather_rows(cat(outs, (72,)), total, group),
            "non_collision": gather_rows(cat(ncs, ()), total, group),
            "contact": gather_rows(cat(cts, ()), total, group)}

"""Physical-plausibility scores of fitted / generated bodies.

Mirrors the scoring block of utils/utils_eval_collision_habitat.py:91-140 (printed as
`--collision_mean` / `--contact_mean` at :232-233): per body, with body_sdf = SDF at the 10 475 vertices,

    contact score        = 1 if any vertex has sdf < 0 else 0
    non-collision score  = (#vertices with sdf > 0) / V      if any vertex has sdf < 0
                           1.0 (= V/V)                        otherwise

The reference evaluates one body at a time (batch_size 1, :211); here a whole batch runs through the
same LBS and SDF kernels as the fitting loop (no gradients), and the counts stay per body.
"""
from __future__ import annotations

import torch

from . import sdf as sdf_mod


@torch.no_grad()
def collision_scores(op, xh, cam_ext):
    """op: a FittingOP (body model, VPoser, scene SDF); xh [B,72]; cam_ext [B|1,4,4].
    Returns (non_collision [B], contact [B]) float32 tensors on the device."""
    verts = op.body_verts(xh, cam_ext)                                   # scene frame, [B,V,3]
    sdf, _, _ = sdf_mod.sdf_forward(op.scene_sdf, verts.contiguous(), want_grad=False)
    V = sdf.shape[1]
    any_neg = (sdf < 0).any(dim=1)
    pos = (sdf > 0).sum(dim=1).to(torch.float32) / float(V)
    non_collision = torch.where(any_neg, pos, torch.ones_like(pos))
    return non_collision, any_neg.to(torch.float32)

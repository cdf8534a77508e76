"""Scene SDF lookup -- replaces the normalise + F.grid_sample block of
source/fitting_habitat.py:145-152 (also fitting_proxe.py:144-151, train_s2.py:182-189,
utils/utils_eval_collision_habitat.py:121-128) with torch-1.2 semantics (align_corners=True,
padding_mode='border').

`SceneSDF` keeps ONE grid per scene on the device (the reference replicates the 64 MiB grid
once per body, fitting_proxe.py:90).  `grid_sample_sdf` is the call-compatible shim for
unchanged reference code: it takes the reference's normalised, swizzled grid tensor.
"""
from __future__ import annotations

import json

import numpy as np
import torch
from torch.autograd import Function

from . import _lib


MAX_SCENES = 16     # kMaxScenes in csrc/sdf.cu


class SceneSDF:
    """S scenes: sdf [S,D,D,D] on the device, grid_min / grid_max [S,3] on the host."""

    def __init__(self, sdf, grid_min, grid_max, device="cuda"):
        sdf = torch.as_tensor(np.asarray(sdf) if not torch.is_tensor(sdf) else sdf, dtype=torch.float32)
        if sdf.dim() == 3:
            sdf = sdf.unsqueeze(0)
        if sdf.dim() != 4 or not (sdf.shape[1] == sdf.shape[2] == sdf.shape[3]):
            raise ValueError(f"sdf must be [S,D,D,D], got {tuple(sdf.shape)}")
        self.sdf = sdf.to(device).contiguous()
        self.num_scenes, self.dim = int(sdf.shape[0]), int(sdf.shape[1])
        self.grid_min = np.ascontiguousarray(np.asarray(grid_min, dtype=np.float32).reshape(self.num_scenes, 3))
        self.grid_max = np.ascontiguousarray(np.asarray(grid_max, dtype=np.float32).reshape(self.num_scenes, 3))
        self._gmin_t = torch.from_numpy(self.grid_min)
        self._gmax_t = torch.from_numpy(self.grid_max)

    @classmethod
    def from_files(cls, prefix: str, device="cuda"):
        """<prefix>.json {min,max,dim} + <prefix>_sdf.npy (fitting_habitat.py:80-90)."""
        with open(prefix + ".json") as f:
            d = json.load(f)
        dim = int(d["dim"])
        sdf = np.load(prefix + "_sdf.npy").reshape(dim, dim, dim)
        return cls(sdf, np.array(d["min"]), np.array(d["max"]), device=device)

    def lookup(self, verts, body_scene=None, with_partials=False):
        return sdf_lookup(self, verts, body_scene, with_partials)


def sdf_forward(scene: SceneSDF, verts, body_scene=None, want_grad=True, want_partials=False):
    """verts [B,V,3] scene frame -> values [B,V], grad [B,V,3] | None, partial [B,np,2] | None."""
    _lib.require_cuda(verts, scene.sdf, body_scene)
    if verts.dtype != torch.float32 or verts.dim() != 3 or verts.shape[-1] != 3:
        raise TypeError("verts must be float32 [B,V,3]")
    verts = verts.contiguous()
    B, V, _ = verts.shape
    out = torch.empty(B, V, dtype=torch.float32, device=verts.device)
    grad = torch.empty(B, V, 3, dtype=torch.float32, device=verts.device) if want_grad else None
    L = _lib.lib()
    partial = None
    if want_partials:
        partial = torch.empty(B, L.psi_sdf_num_partials(V), 2, dtype=torch.float32, device=verts.device)
    if body_scene is not None and body_scene.dtype != torch.int32:
        raise TypeError("body_scene must be int32")
    with torch.cuda.device(verts.device):
        rc = L.psi_sdf_fwd(_lib.ptr(scene.sdf), scene.num_scenes, scene.dim, _lib.ptr(scene._gmin_t),
                           _lib.ptr(scene._gmax_t), _lib.ptr(verts), B, V, _lib.ptr(body_scene),
                           _lib.ptr(out), _lib.ptr(grad), _lib.ptr(partial), _lib.stream_ptr())
    _lib.check(rc, "psi_sdf_fwd")
    return out, grad, partial


class _SdfLookup(Function):
    @staticmethod
    def forward(ctx, verts, scene, body_scene, with_partials):
        out, grad, partial = sdf_forward(scene, verts, body_scene, want_grad=True,
                                         want_partials=with_partials)
        ctx.save_for_backward(grad)
        if with_partials:
            ctx.mark_non_differentiable(partial)
            return out, partial
        return out, None

    @staticmethod
    def backward(ctx, gout, _gpartial):
        (grad,) = ctx.saved_tensors
        gout = gout.contiguous()
        gv = torch.empty_like(grad)
        with torch.cuda.device(grad.device):
            rc = _lib.lib().psi_sdf_bwd(_lib.ptr(gout), _lib.ptr(grad), gout.numel(), _lib.ptr(gv),
                                        _lib.stream_ptr())
        _lib.check(rc, "psi_sdf_bwd")
        return gv, None, None, None


def sdf_lookup(scene: SceneSDF, verts, body_scene=None, with_partials=False):
    """Differentiable body_sdf [B,V] (== `body_sdf_batch` of fitting_habitat.py:150 reshaped).
    with_partials=True also returns the per-chunk (sum of -sdf, count) over sdf<0."""
    out, partial = _SdfLookup.apply(verts, scene, body_scene, with_partials)
    return (out, partial) if with_partials else out


def grid_sample_sdf(input, grid, padding_mode="border", grid_min=None, grid_max=None):
    """Shim with F.grid_sample's call shape for unchanged reference code:
        F.grid_sample(s_sdf.unsqueeze(1), norm_verts[:,:,[2,1,0]].view(-1,V,1,1,3), padding_mode='border')
    `input` [B,1,D,D,D]: one grid per batch entry (train_s2.py:182-189 samples a different scene per body); a
    batch-expanded single scene (stride 0, what `.expand` gives) is read once.  The reference's `.repeat`ed
    copies (fitting_proxe.py:90) work too, as B grids.  `grid` [B,V,1,1,3] already normalised to [-1,1] in
    (z,y,x) order; gradients flow to `grid`.  Returns [B,1,V,1,1]."""
    if padding_mode != "border":
        raise ValueError("only padding_mode='border' is used on the PSI path")
    B, V = grid.shape[0], grid.shape[1]
    D = input.shape[-1]
    shared = input.shape[0] == 1 or input.stride(0) == 0
    verts = grid.reshape(B, V, 3)[:, :, [2, 1, 0]].contiguous()   # back to (x,y,z), still in [-1,1]

    def unit_bank(vol):                                            # [S,D,D,D] sampled over [-1,1]^3
        scene = SceneSDF.__new__(SceneSDF)
        scene.sdf = vol.contiguous()
        S = int(vol.shape[0])
        scene.num_scenes, scene.dim = S, D
        scene.grid_min = np.full((S, 3), -1.0, dtype=np.float32)
        scene.grid_max = np.full((S, 3), 1.0, dtype=np.float32)
        scene._gmin_t = torch.from_numpy(scene.grid_min)
        scene._gmax_t = torch.from_numpy(scene.grid_max)
        return scene

    if shared:
        return sdf_lookup(unit_bank(input[0:1, 0]), verts).view(B, 1, V, 1, 1)
    outs = []
    for b0 in range(0, B, MAX_SCENES):                             # psi_sdf_fwd addresses up to 16 grids per call
        nb = min(MAX_SCENES, B - b0)
        ids = torch.arange(nb, dtype=torch.int32, device=grid.device)
        outs.append(sdf_lookup(unit_bank(input[b0:b0 + nb, 0]), verts[b0:b0 + nb], ids))
    return torch.cat(outs, 0).view(B, 1, V, 1, 1)

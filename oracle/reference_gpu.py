"""The REFERENCE's own GPU path on the same B200 -- MEASUREMENT INFRASTRUCTURE, not a product path.

Only bench.py (`reference_gpu` block) and tools/ may import this module.  It is the "fairest beat-THAT-kernel
comparison" of SURVEY.md 8(d): the fitting loop exactly as source/fitting_habitat.py:103-164,177-191 runs it on
a GPU --

  * Chamfer: the reference's chamfer.cu / chamfer_cuda.cpp compiled UNMODIFIED for sm_100a into
    oracle/_ref/chamfer_ref.so (oracle/build_ref.py), driven the way dist_chamfer.py:13-46 drives it:
    both directions computed, outputs pre-zeroed, backward accumulates with atomics;
  * SMPL-X: the lbs.py:34-262 arithmetic as stock torch CUDA ops (oracle.lbs, ~25 kernels + the 54-step chain);
  * SDF: torch's grid_sampler_3d with the torch-1.2 default spelled out (align_corners=True);
  * VPoser decode, rotation conversions, Adam: stock torch.

Kinder to the reference than its scripts in two ways: the 64 MiB SDF is expanded, not replicated, over the batch
(fitting_proxe.py:90 makes B copies) and the contact ids are not re-read from JSON every iteration (cvae.py:99-115).
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

from . import oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = None


def ref_ext():
    """The prebuilt reference extension, or None when oracle/_ref/chamfer_ref.so is absent."""
    global _REF
    if _REF is None:
        if _HERE not in sys.path:
            sys.path.insert(0, _HERE)
        import build_ref
        _REF = build_ref.load_ref() or False
    return _REF or None


class RefChamfer(torch.autograd.Function):
    """dist_chamfer.py:13-46 over the reference kernels (legacy default stream: call with torch's default stream
    current)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        ext = ref_ext()
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dist1 = torch.zeros(B, n, device=xyz1.device)
        dist2 = torch.zeros(B, m, device=xyz1.device)
        idx1 = torch.zeros(B, n, dtype=torch.int32, device=xyz1.device)
        idx2 = torch.zeros(B, m, dtype=torch.int32, device=xyz1.device)
        ext.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        g1 = torch.zeros_like(xyz1)
        g2 = torch.zeros_like(xyz2)
        ref_ext().backward(xyz1, xyz2, g1, g2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
        return g1, g2


class ReferenceGpuLoop:
    """cal_loss + backward + Adam of fitting_habitat.py on CUDA tensors (loss_mode 'independent' = the sum over
    bodies of the B = 1 loss, the same objective the psi loop optimises)."""

    def __init__(self, model, vposer_weights, scene, contact_ids, weights, device, robust_c=1.0):
        t = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
        self.smplx = oracle.SMPLXOracle(model).to(device)
        self.vposer = oracle.VPoserDecoderOracle(vposer_weights).to(device)
        self.sdf, self.gmin, self.gmax = t(scene.sdf), t(scene.grid_min), t(scene.grid_max)
        self.points = t(scene.points)
        self.cid = torch.as_tensor(contact_ids, dtype=torch.long, device=device)
        self.w, self.c, self.device = weights, robust_c, device

    def cal_loss(self, xhr, xhr_rec, cam_ext):
        B = xhr_rec.shape[0]
        red = lambda x: x.reshape(B, -1).mean(dim=1).sum()
        loss_rec = self.w["weight_loss_rec"] * red((xhr - xhr_rec).abs())
        xh = oracle.convert_to_3D_rot(xhr_rec)
        z = xh[:, 16:48]
        loss_vposer = self.w["weight_loss_vposer"] * red(z ** 2)
        verts, _ = self.smplx(body_pose=self.vposer.decode(z), transl=xh[:, :3], global_orient=xh[:, 3:6],
                              betas=xh[:, 6:16], left_hand_pose=xh[:, 48:60], right_hand_pose=xh[:, 60:])
        verts = oracle.verts_transform(verts, cam_ext)
        contact = verts[:, self.cid, :].contiguous()
        d, _ = RefChamfer.apply(contact, self.points.unsqueeze(0).repeat(B, 1, 1))     # fitting_habitat.py:93-96,137-139
        s = torch.sqrt(d + 1e-4)
        loss_contact = self.w["weight_contact"] * red(s / (s + self.c))
        body_sdf = oracle.sdf_lookup_torch(self.sdf, self.gmin, self.gmax, verts)
        neg = body_sdf < 0
        cnt = neg.sum(dim=1).clamp(min=1).to(body_sdf.dtype)
        loss_collision = self.w["weight_collision"] * ((-body_sdf * neg).sum(dim=1) / cnt).sum()
        return loss_rec, loss_vposer, loss_contact, loss_collision

    def fit(self, xh, cam_ext, num_iter, lr=0.1):
        xhr = oracle.convert_to_6D_rot(xh)
        xhr_rec = xhr.clone().requires_grad_(True)
        opt = torch.optim.Adam([xhr_rec], lr=lr)
        cam = cam_ext.expand(xh.shape[0], -1, -1)
        for _ in range(num_iter):
            opt.zero_grad()
            sum(self.cal_loss(xhr, xhr_rec, cam)).backward()
            opt.step()
        return oracle.convert_to_3D_rot(xhr_rec.detach())

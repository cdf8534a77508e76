"""The geometry block of CVAE training (BASELINE config 4) on the psi kernels.

Mirrors the scene-aware part of TrainOP.cal_loss, source/train_s2.py:136-202 (train_s1.py:136-204 is
the same block): VPoser L2, contact (Chamfer of the contact vertices against the scene vertices,
robustifier constant 1.0) and SDF penetration, with the reference's batch-mean semantics and its
epoch gating (contact / penetration active only for ep > 0.75*epochs, :174-176,198-200).  The CVAE,
its optimiser, KL and reconstruction terms stay stock torch (out of scope).

Differences by design: scenes live in a BANK (S grids, S point clouds + indices) addressed by a
per-sample scene id instead of being replicated per sample (batch_gen_hdf5.py:222-257); the block
is skipped entirely while both gates are 0 (the reference computes it and multiplies by 0.0);
the `.item()` sync of the penetration term is replaced by a branch-free expression.
"""
from __future__ import annotations

import numpy as np
import torch

from . import chamfer, sdf as sdf_mod
from .geometry import BodyParamParser


class SceneLossBlock:
    def __init__(self, body_mesh_model, vposer, scenes, contact_ids, weight_loss_vposer, weight_contact,
                 weight_collision, device="cuda"):
        """scenes: list of objects with .sdf [D,D,D], .grid_min, .grid_max, .points [M,3]."""
        self.device = torch.device(device)
        self.body_mesh_model = body_mesh_model.to(self.device)
        self.vposer = vposer.to(self.device)
        self.scene_sdf = sdf_mod.SceneSDF(np.stack([s.sdf for s in scenes]), np.stack([s.grid_min for s in scenes]),
                                          np.stack([s.grid_max for s in scenes]), device=self.device)
        self.scene_index = [chamfer.SceneIndex(torch.tensor(np.asarray(s.points), dtype=torch.float32, device=self.device))
                            for s in scenes]
        self.contact_ids = torch.as_tensor(np.asarray(contact_ids), dtype=torch.long, device=self.device)
        self.weight_loss_vposer, self.weight_contact, self.weight_collision = weight_loss_vposer, weight_contact, weight_collision

    def __call__(self, xh_rec, cam_ext, scene_ids, ep, epochs):
        """xh_rec [B,72] (the CVAE reconstruction after recover_global_T, train_s2.py:116),
        cam_ext [B,4,4], scene_ids: sequence of B ints.  Returns (loss_contact, loss_vposer,
        loss_sdf_pene) as in train_s2.py:204."""
        loss_vposer = self.weight_loss_vposer * torch.mean(xh_rec[:, 16:48] ** 2)
        active = ep > 0.75 * epochs                      # fcc = fsp (train_s2.py:174-176,198-200)
        zero = torch.zeros((), dtype=torch.float32, device=xh_rec.device)
        if not active:
            return zero, loss_vposer, zero
        B = xh_rec.shape[0]
        bp = BodyParamParser.body_params_encapsulate_batch(xh_rec)
        joint_rot = self.vposer.decode(bp["body_pose_vp"], output_type="aa").view(B, -1)
        verts = self.body_mesh_model(return_verts=True, body_pose=joint_rot, transl=bp["transl"],
                                     global_orient=bp["global_orient"], betas=bp["betas"],
                                     left_hand_pose=bp["left_hand_pose"], right_hand_pose=bp["right_hand_pose"],
                                     cam_ext=cam_ext).vertices
        contact = verts[:, self.contact_ids, :]
        ids = torch.as_tensor(list(scene_ids), dtype=torch.int64)
        dists = torch.empty(B, contact.shape[1], dtype=torch.float32, device=verts.device)
        for s in sorted(set(ids.tolist())):              # one exact NN query per scene present
            rows = torch.nonzero(ids == s).flatten().to(verts.device)
            d, _ = chamfer.nn_distance(contact[rows].contiguous(), self.scene_index[s])
            dists = dists.index_copy(0, rows, d)
        sq = torch.sqrt(dists + 1e-4)
        loss_contact = self.weight_contact * torch.mean(sq / (sq + 1.0))
        body_sdf = self.scene_sdf.lookup(verts, body_scene=ids.to(torch.int32).to(verts.device))
        neg = body_sdf < 0
        cnt = neg.sum().clamp(min=1).to(body_sdf.dtype)
        loss_sdf_pene = self.weight_collision * (((-body_sdf) * neg).sum() / cnt)
        return loss_contact, loss_vposer, loss_sdf_pene


class TrainStep:
    """One optimisation step of CVAE stage-2 training (BASELINE config 4): the loss assembly of
    TrainOP.cal_loss, source/train_s2.py:102-204, around ANY `model_h(xhnr, eps_g, eps_l, xs) ->
    (xhnr_rec, mu_g, logsigma2_g, mu_l, logsigma2_l)` (the CVAE and its ResNet18 scene encoder stay stock
    torch modules: out of scope).  The network runs under bf16 autocast (north_star:
    "bf16, 1->8 B200 data-parallel"), the geometry block (VPoser decode, SMPL-X, contact NN, SDF) in
    FP32 on the psi kernels.  Data parallelism is plain DDP around `model_h`: the geometry block has no
    trainable parameters (train_s2.py:66-67), so every rank evaluates it on its own samples and only the
    CVAE gradients are all-reduced."""

    def __init__(self, model_h, block: SceneLossBlock, optimizer, weight_loss_rec_h=1.0, weight_loss_kl=1.0,
                 epochs=30, loss_weight_anealing=True, autocast_dtype=torch.bfloat16):
        self.model_h, self.block, self.optimizer = model_h, block, optimizer
        self.weight_loss_rec_h, self.weight_loss_kl = weight_loss_rec_h, weight_loss_kl
        self.epochs, self.loss_weight_anealing, self.autocast_dtype = epochs, loss_weight_anealing, autocast_dtype

    def cal_loss(self, xs, xh, eps_g, eps_l, cam_ext, cam_int, max_d, scene_ids, ep):
        """Returns the seven terms of train_s2.py:204 in its order."""
        from .geometry import GeometryTransformer
        import torch.nn.functional as F
        xhn = GeometryTransformer.normalize_global_T(xh, cam_int, max_d)
        xhnr = GeometryTransformer.convert_to_6D_rot(xhn)
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            xhnr_rec, mu_g, logsigma2_g, mu_l, logsigma2_l = self.model_h(xhnr, eps_g, eps_l, xs)
        xhnr_rec, mu_g, logsigma2_g, mu_l, logsigma2_l = (t.float() for t in (xhnr_rec, mu_g, logsigma2_g, mu_l, logsigma2_l))
        xhn_rec = GeometryTransformer.convert_to_3D_rot(xhnr_rec)
        xh_rec = GeometryTransformer.recover_global_T(xhn_rec, cam_int, max_d)
        loss_rec_t = self.weight_loss_rec_h * (0.5 * F.l1_loss(xhnr_rec[:, :3], xhnr[:, :3]) +
                                               0.5 * F.l1_loss(xh_rec[:, :3], xh[:, :3]))
        loss_rec_p = self.weight_loss_rec_h * F.l1_loss(xhnr_rec[:, 3:], xhnr[:, 3:])
        fca = min(1.0, max(float(ep) / (self.epochs * 0.75), 0)) if self.loss_weight_anealing else 1.0
        kl = lambda mu, ls: 0.5 * torch.mean(torch.exp(ls) + mu ** 2 - 1.0 - ls)
        loss_kl_g = fca ** 2 * self.weight_loss_kl * kl(mu_g, logsigma2_g)
        loss_kl_l = fca ** 2 * self.weight_loss_kl * kl(mu_l, logsigma2_l)
        loss_contact, loss_vposer, loss_sdf_pene = self.block(xh_rec, cam_ext, scene_ids, ep, self.epochs)
        return loss_rec_t, loss_rec_p, loss_kl_g, loss_kl_l, loss_contact, loss_vposer, loss_sdf_pene

    def step(self, *batch, ep):
        """forward + backward + optimizer.step (train_s2.py:253-262); returns the detached terms."""
        self.optimizer.zero_grad(set_to_none=True)
        terms = self.cal_loss(*batch, ep)
        sum(terms).backward()
        self.optimizer.step()
        return torch.stack([t.detach() for t in terms])

"""SMPL-X body model -- drop-in for `smplx.create(...)` as the reference calls it.

Reference call sites: source/fitting_habitat.py:57-71 (create) and :126-129 (forward),
fitting_proxe.py:55-69,125-128, train_s2.py:72-87,156-159.  The reference imports the pinned
third-party `smplx==0.1.13` (requirements.txt:95); its LBS math is vendored at
human_body_prior/body_model/lbs.py:34-262 and the wrapper's concatenation order / shape split
/ translation are mirrored at human_body_prior/body_model/body_model.py:223-247.  Hand PCA and
`pose_mean` follow the published smplx 0.1.x behaviour.

The vertex path runs in psi_lbs_fwd / psi_lbs_bwd (lib/libpsi_b200.so); the few [B,45]-sized
parameter assembly ops (hand PCA, concatenation, pose_mean) are host-orchestrated torch ops.
"""
from __future__ import annotations

import ctypes
import os
from collections import namedtuple

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

from . import _lib

ModelOutput = namedtuple("ModelOutput", ["vertices", "joints", "full_pose", "betas", "global_orient",
                                         "body_pose", "expression", "left_hand_pose",
                                         "right_hand_pose", "jaw_pose"])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)


class _ModelHandle:
    """Owns a psi_lbs_model (device constants in the kernels' layout)."""

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, weights, parents, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PsiError("the SMPL-X kernels need a CUDA device (no CPU fallback)")
        f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        v_template, shapedirs, posedirs = f32(v_template), f32(shapedirs), f32(posedirs)
        J_regressor, weights = f32(J_regressor), f32(weights)
        par = np.ascontiguousarray(np.asarray(parents, dtype=np.int32))
        self.V, self.J, self.NB = int(v_template.shape[0]), int(J_regressor.shape[0]), int(shapedirs.shape[2])
        assert posedirs.shape == ((self.J - 1) * 9, self.V * 3), posedirs.shape
        assert weights.shape == (self.V, self.J) and J_regressor.shape == (self.J, self.V)
        h = ctypes.c_void_p()
        hp = lambda a: ctypes.c_void_p(a.ctypes.data)
        with torch.cuda.device(self.device):
            rc = _lib.lib().psi_lbs_model_create(ctypes.byref(h), self.V, self.J, self.NB, hp(v_template),
                                                 hp(shapedirs), hp(posedirs), hp(J_regressor),
                                                 hp(weights), hp(par), _lib.stream_ptr())
        _lib.check(rc, "psi_lbs_model_create")
        self.h = h

    def nbytes(self) -> int:
        return int(_lib.lib().psi_lbs_model_bytes(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().psi_lbs_model_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _LBS(Function):
    """verts, joints = lbs(betas[B,NB], pose[B,J*3]) (+transl, + rigid cam transform)."""

    @staticmethod
    def forward(ctx, betas, pose, transl, cam, handle, want_joints):
        _lib.require_cuda(betas, pose, transl, cam)
        L = _lib.lib()
        betas, pose = betas.contiguous().float(), pose.contiguous().float()
        B = betas.shape[0]
        dev = betas.device
        if pose.shape != (B, handle.J * 3) or betas.shape[1] != handle.NB:
            raise ValueError(f"lbs: betas {tuple(betas.shape)} pose {tuple(pose.shape)}")
        transl_c = transl.contiguous().float() if transl is not None else None
        cam_c, cam_stride = None, 0
        if cam is not None:
            cam_c = cam.reshape(-1, 16)[:, :12].contiguous().float() if cam.shape[-2:] == (4, 4) \
                else cam.reshape(-1, 12).contiguous().float()
            cam_stride = 12 if cam_c.shape[0] == B else 0
            if cam_c.shape[0] not in (1, B):
                raise ValueError("cam must be [B,4,4], [1,4,4] or [4,4]")
        verts = torch.empty(B, handle.V, 3, dtype=torch.float32, device=dev)
        joints = torch.empty(B, handle.J, 3, dtype=torch.float32, device=dev) if want_joints else None
        saved = torch.empty(int(L.psi_lbs_saved_floats(handle.h, B)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.psi_lbs_fwd(handle.h, B, _lib.ptr(betas), _lib.ptr(pose), _lib.ptr(transl_c),
                               _lib.ptr(cam_c), cam_stride, None, 0, _lib.ptr(verts), _lib.ptr(joints),
                               _lib.ptr(saved), _lib.stream_ptr())
        _lib.check(rc, "psi_lbs_fwd")
        ctx.handle = handle
        ctx.cam_stride = cam_stride
        ctx.has_transl = transl is not None
        ctx.want_joints = want_joints
        ctx.save_for_backward(betas, pose, cam_c, saved)
        if want_joints:
            return verts, joints
        return verts, None

    @staticmethod
    def backward(ctx, gverts, gjoints):
        betas, pose, cam_c, saved = ctx.saved_tensors
        handle = ctx.handle
        L = _lib.lib()
        B = betas.shape[0]
        dev = betas.device
        if gverts is None:
            gverts = torch.zeros(B, handle.V, 3, dtype=torch.float32, device=dev)
        gverts = gverts.contiguous().float()
        gj = gjoints.contiguous().float() if (gjoints is not None and ctx.want_joints) else None
        gbetas = torch.empty_like(betas)
        gpose = torch.empty_like(pose)
        gtransl = torch.empty(B, 3, dtype=torch.float32, device=dev) if ctx.has_transl else None
        nbytes = int(L.psi_lbs_bwd_workspace_bytes(handle.h, B))
        ws = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.psi_lbs_bwd(handle.h, B, _lib.ptr(betas), _lib.ptr(pose), _lib.ptr(cam_c),
                               ctx.cam_stride, _lib.ptr(saved), _lib.ptr(gverts), _lib.ptr(gj),
                               _lib.ptr(gbetas), _lib.ptr(gpose), _lib.ptr(gtransl), None, 0, _lib.ptr(ws),
                               ws.numel() * 4, _lib.stream_ptr())
        _lib.check(rc, "psi_lbs_bwd")
        return gbetas, gpose, gtransl, None, None, None


def lbs(betas, pose, handle, transl=None, cam=None, want_joints=False):
    """Functional LBS.  cam: optional rigid [B|1,4,4] transform fused after the translation
    (GeometryTransformer.verts_transform, source/cvae.py:141-149); not differentiated."""
    return _LBS.apply(betas, pose, transl, cam, handle, want_joints)


# npz keys read from an SMPL-X model file (human_body_prior/body_model/body_model.py:86-135 plus smplx's hand PCA)
MODEL_KEYS = ("v_template", "shapedirs", "posedirs", "J_regressor", "kintree_table", "weights", "f",
              "hands_componentsl", "hands_componentsr", "hands_meanl", "hands_meanr")


class SMPLX(nn.Module):
    """`smplx.create(model_path, model_type='smplx', ...)` result.  Same keyword surface as the
    reference's call; every `create_*` flag makes a zero nn.Parameter of shape [batch_size, .]."""

    NUM_BODY_JOINTS = 21
    NUM_HAND_JOINTS = 15

    def __init__(self, model_path=None, gender="neutral", ext="npz", num_betas=10, num_expression_coeffs=10,
                 num_pca_comps=6, use_pca=True, flat_hand_mean=False, batch_size=1,
                 create_global_orient=True, create_body_pose=True, create_betas=True,
                 create_left_hand_pose=True, create_right_hand_pose=True, create_expression=True,
                 create_jaw_pose=True, create_leye_pose=True, create_reye_pose=True, create_transl=True,
                 dtype=torch.float32, model_data=None, **kwargs):
        super().__init__()
        if dtype != torch.float32:
            raise TypeError("the B200 kernels compute in float32")
        if model_data is None:
            if os.path.isdir(model_path):
                model_path = os.path.join(model_path, f"SMPLX_{gender.upper()}.{ext}")
            if not os.path.exists(model_path):
                raise FileNotFoundError(model_path)
            # only the keys the path reads: the published SMPLX_*.npz also carries object-dtype
            # entries (joint2num, part2num, ...) that np.load refuses with allow_pickle=False and that smplx
            # itself unpickles -- they are never touched here, so nothing has to be unpickled
            with np.load(model_path, allow_pickle=False) as z:
                missing = [k for k in MODEL_KEYS if k not in z.files]
                if missing:
                    raise KeyError(f"{model_path}: missing SMPL-X keys {missing}")
                model_data = {k: z[k] for k in MODEL_KEYS}
        self._model_data = {k: np.asarray(v) for k, v in model_data.items() if k in MODEL_KEYS}
        md = self._model_data
        self.batch_size = batch_size
        self.num_betas, self.num_expression_coeffs = num_betas, num_expression_coeffs
        self.use_pca, self.num_pca_comps, self.flat_hand_mean = use_pca, num_pca_comps, flat_hand_mean
        sd = md["shapedirs"]
        begin = 300 if sd.shape[-1] > 300 else 10                     # body_model.py:103-105
        self._shapedirs = np.concatenate([sd[:, :, :num_betas], sd[:, :, begin:begin + num_expression_coeffs]], -1)
        pd = md["posedirs"]
        self._posedirs = pd.reshape(pd.shape[0] * 3, -1).T            # body_model.py:123-125
        parents = md["kintree_table"][0].astype(np.int64).copy()
        parents[0] = -1
        self.register_buffer("parents", torch.tensor(parents))
        self.register_buffer("faces_tensor", torch.tensor(md["f"].astype(np.int64)))
        self.faces = md["f"]
        nj = md["J_regressor"].shape[0]
        self.num_joints = nj
        self.register_buffer("left_hand_components", torch.tensor(md["hands_componentsl"][:num_pca_comps], dtype=dtype))
        self.register_buffer("right_hand_components", torch.tensor(md["hands_componentsr"][:num_pca_comps], dtype=dtype))
        mean = np.zeros(nj * 3, dtype=np.float32)
        if not flat_hand_mean:
            mean[(nj - 30) * 3:(nj - 15) * 3] = md["hands_meanl"]
            mean[(nj - 15) * 3:] = md["hands_meanr"]
        self.register_buffer("pose_mean", torch.tensor(mean, dtype=dtype))
        hand_dim = num_pca_comps if use_pca else 45

        def param(flag, name, dim):
            if flag:
                self.register_parameter(name, nn.Parameter(torch.zeros(batch_size, dim, dtype=dtype)))

        param(create_global_orient, "global_orient", 3)
        param(create_body_pose, "body_pose", self.NUM_BODY_JOINTS * 3)
        param(create_betas, "betas", num_betas)
        param(create_left_hand_pose, "left_hand_pose", hand_dim)
        param(create_right_hand_pose, "right_hand_pose", hand_dim)
        param(create_expression, "expression", num_expression_coeffs)
        param(create_jaw_pose, "jaw_pose", 3)
        param(create_leye_pose, "leye_pose", 3)
        param(create_reye_pose, "reye_pose", 3)
        param(create_transl, "transl", 3)
        self._handle = None

    # device constants are created lazily on first use (after .to(device))
    def handle(self, device=None) -> _ModelHandle:
        device = torch.device(device) if device is not None else self.pose_mean.device
        if self._handle is None or self._handle.device != device:
            md = self._model_data
            self._handle = _ModelHandle(md["v_template"], self._shapedirs, self._posedirs,
                                        md["J_regressor"], md["weights"], self.parents.cpu().numpy(), device)
        return self._handle

    def coherent_handle(self, device=None) -> _ModelHandle:
        """The same model with its vertices RE-ORDERED so that consecutive vertices are skinned to the same joints
        (dominant joint, joints depth-first along the tree, kd order of the template inside a joint).  The kernels
        take one thread per vertex and gather the joints' transforms from a shared-memory table: with a vertex order
        that scatters the joints (nothing in the file format promises otherwise) every warp hits most of the table
        and its banks collide; in this order a warp reads one or two rows.  Per-vertex results come out in the new
        order: `handle.perm[k]` = original index of new vertex k, `handle.inv` = its inverse.  Used by the fused
        fitting loop, where vertices never leave the library (fused.py)."""
        device = torch.device(device) if device is not None else self.pose_mean.device
        h = getattr(self, "_chandle", None)
        if h is None or h.device != device:
            from .fused import _spatial_order
            md = self._model_data
            V = md["v_template"].shape[0]
            parents = self.parents.cpu().numpy()
            perm = _spatial_order(np.arange(V), md["v_template"], md["weights"], parents).astype(np.int64)
            assert perm.shape == (V,) and np.array_equal(np.sort(perm), np.arange(V))
            P = self._posedirs.shape[0]
            pd = self._posedirs.reshape(P, V, 3)[:, perm, :].reshape(P, V * 3)
            h = _ModelHandle(md["v_template"][perm], self._shapedirs[perm], pd, md["J_regressor"][:, perm],
                             md["weights"][perm], parents, device)
            h.perm = perm
            h.inv = np.argsort(perm)
            self._chandle = h
        return h

    def _default(self, value, name, B, dim):
        if value is not None:
            return value
        p = getattr(self, name, None)
        if p is not None:
            # smplx fixes the batch size at construction; a module shared by operators of several batch sizes
            # (pipeline.fit_rooms) broadcasts the first row of its default parameter instead of failing
            return p if p.shape[0] == B else p[:1].expand(B, -1)
        return torch.zeros(B, dim, dtype=torch.float32, device=self.pose_mean.device)

    def assemble(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                 right_hand_pose=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None):
        """-> (shape [B,NB], full_pose [B,J*3]) in the order body_model.py:223-235 uses."""
        B = max(x.shape[0] for x in (betas, global_orient, body_pose) if x is not None) \
            if any(x is not None for x in (betas, global_orient, body_pose)) else self.batch_size
        hd = self.num_pca_comps if self.use_pca else 45
        go = self._default(global_orient, "global_orient", B, 3)
        bp = self._default(body_pose, "body_pose", B, 63)
        be = self._default(betas, "betas", B, self.num_betas)
        lh = self._default(left_hand_pose, "left_hand_pose", B, hd)
        rh = self._default(right_hand_pose, "right_hand_pose", B, hd)
        ex = self._default(expression, "expression", B, self.num_expression_coeffs)
        jp = self._default(jaw_pose, "jaw_pose", B, 3)
        le = self._default(leye_pose, "leye_pose", B, 3)
        re = self._default(reye_pose, "reye_pose", B, 3)
        if self.use_pca:
            lh = torch.einsum("bi,ij->bj", lh, self.left_hand_components)
            rh = torch.einsum("bi,ij->bj", rh, self.right_hand_components)
        full_pose = torch.cat([go, bp, jp, le, re, lh, rh], dim=1) + self.pose_mean
        shape = torch.cat([be, ex], dim=-1)
        return shape, full_pose, (be, go, bp, ex, lh, rh, jp)

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                right_hand_pose=None, transl=None, expression=None, jaw_pose=None, leye_pose=None,
                reye_pose=None, return_verts=True, return_full_pose=False, cam_ext=None, **kwargs):
        shape, full_pose, (be, go, bp, ex, lh, rh, jp) = self.assemble(
            betas, global_orient, body_pose, left_hand_pose, right_hand_pose, expression, jaw_pose,
            leye_pose, reye_pose)
        if transl is None:
            transl = getattr(self, "transl", None)
        verts, joints = lbs(shape, full_pose, self.handle(shape.device), transl=transl, cam=cam_ext,
                            want_joints=True)
        return ModelOutput(vertices=verts if return_verts else None, joints=joints,
                           full_pose=full_pose if return_full_pose else None, betas=be,
                           global_orient=go, body_pose=bp, expression=ex, left_hand_pose=lh,
                           right_hand_pose=rh, jaw_pose=jp)


def create(model_path=None, model_type="smplx", **kwargs):
    """smplx.create: `<model_path>/smplx/SMPLX_<GENDER>.npz` (layout: train_s2.py:89-91)."""
    if model_type.lower() != "smplx":
        raise ValueError("only model_type='smplx' is on the PSI path")
    if model_path is not None and os.path.isdir(model_path):
        model_path = os.path.join(model_path, "smplx")
    return SMPLX(model_path, **kwargs)

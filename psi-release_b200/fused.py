"""The fused fitting loop (psi_fit_* in include/psi_b200.h): one C call runs all iterations of
source/fitting_habitat.py:177-191 for a batch of bodies -- VPoser decode, rotation chain, SMPL-X,
contact NN, SDF, losses, backward and Adam as 11 kernel launches per iteration from a CUDA graph.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


class FusedFit:
    """Owns a psi_fit_ctx bound to (body model handle, scene index, scene SDF, VPoser decoder)."""

    def __init__(self, batch_size, model_handle, body_model, scene_index, scene_sdf, vposer, contact_ids,
                 weights, robust_c, lr, use_graph=True):
        if scene_sdf.num_scenes != 1:
            raise ValueError("the fused loop fits one scene per context")
        self.device = model_handle.device
        self.B = int(batch_size)
        self._keep = (model_handle, scene_index, scene_sdf)     # borrowed device objects
        f32 = lambda t: np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))
        sd = {k: f32(v) for k, v in vposer.state_dict().items()}
        W1, b1 = sd["bodyprior_dec_fc1.weight"], sd["bodyprior_dec_fc1.bias"]
        W2, b2 = sd["bodyprior_dec_fc2.weight"], sd["bodyprior_dec_fc2.bias"]
        W3, b3 = sd["bodyprior_dec_out.weight"], sd["bodyprior_dec_out.bias"]
        hl, hr = f32(body_model.left_hand_components), f32(body_model.right_hand_components)
        pm = f32(body_model.pose_mean)
        # Query order = order of first appearance in this list.  Sorting the ids along a Morton curve
        # of the template mesh makes the 32 queries of a warp neighbours on the body, which the
        # thread-per-query NN kernel needs to be fast; the result does not depend on the order.
        cid = np.asarray(contact_ids, dtype=np.int64)
        vt = np.asarray(body_model._model_data["v_template"], dtype=np.float64)
        cell = ((vt - vt.min(0)) / np.maximum(vt.max(0) - vt.min(0), 1e-12) * 1023.0 + 0.5).astype(np.int64)

        def _spread(v):
            v = v & 0x3FF
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            return (v | (v << 2)) & 0x09249249

        rank = _spread(cell[:, 0]) | (_spread(cell[:, 1]) << 1) | (_spread(cell[:, 2]) << 2)
        cid = np.ascontiguousarray(cid[np.argsort(rank[cid], kind="stable")].astype(np.int32))
        cfg = _lib.FitConfig(B=self.B, use_graph=1 if use_graph else 0,
                             w_rec=float(weights["weight_loss_rec"]), w_vposer=float(weights["weight_loss_vposer"]),
                             w_contact=float(weights["weight_contact"]), w_collision=float(weights["weight_collision"]),
                             robust_c=float(robust_c), lr=float(lr), beta1=0.9, beta2=0.999, eps=1e-8)
        hp = lambda a: ctypes.c_void_p(a.ctypes.data)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = _lib.lib().psi_fit_create(
                ctypes.byref(h), model_handle.h, model_handle.V, model_handle.J, model_handle.NB,
                scene_index.h, _lib.ptr(scene_index.points), _lib.ptr(scene_sdf.sdf), scene_sdf.dim,
                hp(scene_sdf.grid_min), hp(scene_sdf.grid_max), hp(W1), hp(b1), hp(W2), hp(b2), hp(W3), hp(b3),
                W1.shape[1], W1.shape[0], W3.shape[0] // 6, hp(hl), hp(hr), hp(pm), hl.shape[0],
                hp(cid), cid.shape[0], ctypes.byref(cfg), _lib.stream_ptr())
        _lib.check(rc, "psi_fit_create")
        self.h = h
        self.xdim = 19 + W1.shape[1] + 2 * hl.shape[0]

    def run(self, xhr, cam_ext, num_iter):
        """xhr [B,75] device; cam_ext [B|1,4,4] -> (fitted xhr [B,75], losses [B,4])."""
        _lib.require_cuda(xhr, cam_ext)
        if xhr.shape != (self.B, self.xdim):
            raise ValueError(f"xhr must be [{self.B},{self.xdim}], got {tuple(xhr.shape)}")
        xhr = xhr.contiguous().float()
        cam = cam_ext.reshape(-1, 16)[:, :12].contiguous().float()
        if cam.shape[0] not in (1, self.B):
            raise ValueError("cam_ext must be [B,4,4] or [1,4,4]")
        out = torch.empty_like(xhr)
        losses = torch.empty(self.B, 4, dtype=torch.float32, device=xhr.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().psi_fit_run(self.h, _lib.ptr(xhr), _lib.ptr(cam), 12 if cam.shape[0] == self.B else 0,
                                        int(num_iter), _lib.ptr(out), _lib.ptr(losses), _lib.stream_ptr())
        _lib.check(rc, "psi_fit_run")
        return out, losses

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().psi_fit_destroy(self.h)
                self.h = None
        except Exception:
            pass

"""Chamfer / nearest-neighbour distance -- drop-in for the reference's chamfer_pytorch package.

Mirrors (same names, argument meaning, return order):
  chamfer_pytorch/dist_chamfer.py:13-53      chamferFunction / chamferDist  -> (dist1, dist2)
  chamfer_pytorch/dist_chamfer_idx.py:9-52   the *_idx variant               -> (+ idx1, idx2)
over the C ABI psi_chamfer_fwd / psi_chamfer_bwd (include/psi_b200.h), which replaces the
pybind module `chamfer` (chamfer_pytorch/chamfer_cuda.cpp:30-33).

Differences by design: outputs are allocated on the device (the reference builds them on
the CPU and copies, dist_chamfer.py:19-28); kernels run on torch's current stream (the
reference uses the legacy default stream); return codes are checked (the reference ignores
them, dist_chamfer.py:30).  `nn_distance` is the one-direction form the fitting loss needs
(fitting_habitat.py:138 discards dist2) with a scene shared by the whole batch.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from . import _lib


def _check_cloud(x, name):
    if x.dim() != 3 or x.shape[-1] != 3:
        raise ValueError(f"{name} must be [B,N,3], got {tuple(x.shape)}")
    if x.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {x.dtype}")


def _workspace(B, n, m, device):
    nbytes = _lib.lib().psi_nn_workspace_bytes(B, n, m)
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device)


def chamfer_forward(xyz1, xyz2, both_directions=True):
    """chamfer.forward: returns dist1 [B,N], dist2 [B,M] (or None), idx1, idx2 (int32)."""
    _check_cloud(xyz1, "xyz1")
    _check_cloud(xyz2, "xyz2")
    _lib.require_cuda(xyz1, xyz2)
    xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
    B, n, _ = xyz1.shape
    B2, m, _ = xyz2.shape
    if B != B2:
        raise ValueError(f"batch mismatch {B} vs {B2}")
    dev = xyz1.device
    dist1 = torch.zeros(B, n, dtype=torch.float32, device=dev)
    idx1 = torch.zeros(B, n, dtype=torch.int32, device=dev)
    dist2 = idx2 = None
    if both_directions:
        dist2 = torch.zeros(B, m, dtype=torch.float32, device=dev)
        idx2 = torch.zeros(B, m, dtype=torch.int32, device=dev)
    ws = _workspace(B, n, m, dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().psi_chamfer_fwd(_lib.ptr(xyz1), _lib.ptr(xyz2), B, n, m, _lib.ptr(dist1),
                                        _lib.ptr(dist2), _lib.ptr(idx1), _lib.ptr(idx2),
                                        _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "psi_chamfer_fwd")
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """chamfer.backward: returns gradxyz1 [B,N,3], gradxyz2 [B,M,3]."""
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = torch.empty_like(xyz1)
    g2 = torch.empty_like(xyz2)
    with torch.cuda.device(xyz1.device):
        rc = _lib.lib().psi_chamfer_bwd(_lib.ptr(xyz1), _lib.ptr(xyz2), B, n, m,
                                        _lib.ptr(graddist1), _lib.ptr(graddist2), _lib.ptr(idx1),
                                        _lib.ptr(idx2), _lib.ptr(g1), _lib.ptr(g2), _lib.stream_ptr())
    _lib.check(rc, "psi_chamfer_bwd")
    return g1, g2


class chamferFunction(Function):
    """dist_chamfer.py:13-46."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        dist1, dist2, idx1, idx2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        return chamfer_backward(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)


class chamferFunctionIdx(Function):
    """dist_chamfer_idx.py:9-46."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        dist1, dist2, idx1, idx2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        return chamfer_backward(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)


class chamferDist(nn.Module):
    """`chamferDist()(a, b) -> (dist1, dist2)`; construction is free (a new one is built every
    iteration at fitting_habitat.py:137)."""

    def __init__(self, return_idx: bool = False):
        super().__init__()
        self.return_idx = return_idx

    def forward(self, input1, input2):
        if self.return_idx:
            return chamferFunctionIdx.apply(input1, input2)
        return chamferFunction.apply(input1, input2)


# ------------------------------------------------------------------------------------------------
class SceneIndex:
    """Cluster index over a STATIC scene cloud (psi_nn_index_*): nn_forward / nn_distance against
    it return exactly what the brute-force kernel returns (same distances, same indices into
    the ORIGINAL point order), skipping clusters that provably cannot hold the minimum."""

    def __init__(self, points):
        import ctypes
        pts = points.detach().to("cpu", torch.float32).contiguous()
        if pts.dim() != 2 or pts.shape[1] != 3:
            raise ValueError(f"scene points must be [M,3], got {tuple(pts.shape)}")
        if not points.is_cuda:
            raise _lib.PsiError("SceneIndex needs the scene points on a CUDA device")
        self.points = points.contiguous()          # original order, used by the backward gather
        self.device = points.device
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = _lib.lib().psi_nn_index_create(ctypes.byref(h), _lib.ptr(pts), pts.shape[0], _lib.stream_ptr())
        _lib.check(rc, "psi_nn_index_create")
        self.h = h

    def nbytes(self) -> int:
        return int(_lib.lib().psi_nn_index_bytes(self.h))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().psi_nn_index_destroy(self.h)
                self.h = None
        except Exception:
            pass


def nn_forward(query, scene):
    """query [B,N,3]; scene [M,3] (shared by the batch), [B,M,3] or a SceneIndex
    -> dist [B,N], idx [B,N]."""
    _check_cloud(query, "query")
    if isinstance(scene, SceneIndex):
        _lib.require_cuda(query)
        query = query.contiguous()
        B, n, _ = query.shape
        dist = torch.empty(B, n, dtype=torch.float32, device=query.device)
        idx = torch.empty(B, n, dtype=torch.int32, device=query.device)
        with torch.cuda.device(query.device):
            rc = _lib.lib().psi_nn_index_query(scene.h, _lib.ptr(query), n * 3, B, n, None, _lib.ptr(dist),
                                               _lib.ptr(idx), _lib.stream_ptr())
        _lib.check(rc, "psi_nn_index_query")
        return dist, idx
    _lib.require_cuda(query, scene)
    query = query.contiguous()
    scene = scene.contiguous()
    if scene.dtype != torch.float32:
        raise TypeError("scene must be float32")
    B, n, _ = query.shape
    shared = scene.dim() == 2
    m = scene.shape[-2]
    if not shared and scene.shape[0] != B:
        raise ValueError("batch mismatch")
    dist = torch.zeros(B, n, dtype=torch.float32, device=query.device)
    idx = torch.zeros(B, n, dtype=torch.int32, device=query.device)
    ws = _workspace(B, n, m, query.device)
    with torch.cuda.device(query.device):
        rc = _lib.lib().psi_nn_fwd(_lib.ptr(query), n * 3, B, n, _lib.ptr(scene), 0 if shared else m * 3,
                                   m, _lib.ptr(dist), _lib.ptr(idx), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr())
    _lib.check(rc, "psi_nn_fwd")
    return dist, idx


class _NNDistance(Function):
    @staticmethod
    def forward(ctx, query, scene):
        query = query.contiguous()
        if isinstance(scene, SceneIndex):
            dist, idx = nn_forward(query, scene)
            scene = scene.points
        else:
            scene = scene.contiguous()
            dist, idx = nn_forward(query, scene)
        ctx.save_for_backward(query, scene, idx)
        ctx.mark_non_differentiable(idx)
        return dist, idx

    @staticmethod
    def backward(ctx, graddist, _gradidx):
        query, scene, idx = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise _lib.PsiError("nn_distance does not differentiate w.r.t. the scene; use chamferDist")
        B, n, _ = query.shape
        shared = scene.dim() == 2
        m = scene.shape[-2]
        gq = torch.empty_like(query)
        with torch.cuda.device(query.device):
            rc = _lib.lib().psi_nn_bwd(_lib.ptr(query), n * 3, B, n, _lib.ptr(scene),
                                       0 if shared else m * 3, m, _lib.ptr(graddist.contiguous()),
                                       _lib.ptr(idx), _lib.ptr(gq), _lib.stream_ptr())
        _lib.check(rc, "psi_nn_bwd")
        return gq, None


def nn_distance(query, scene):
    """One-direction squared NN distance (dist1 of chamferDist) and index; differentiable in
    `query` with the reference's gradient 2*g*(p - q_idx) (chamfer.cu:165-168)."""
    return _NNDistance.apply(query, scene)

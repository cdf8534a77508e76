"""CPU ORACLE for the PSI fitting hot path -- TEST INFRASTRUCTURE, not a product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (psi-release_b200) never does.

What is restated here and which reference lines it follows:

  * nn_fwd / chamfer_fwd / chamfer_bwd / sdf_fwd  -> ctypes into oracle/libpsi_oracle.so
    (psi_oracle.c; chamfer_pytorch/chamfer.cu:12-195, fitting_habitat.py:145-152)
  * lbs()            human_body_prior/body_model/lbs.py:34-118 (+ helpers :121-262)
  * smplx_forward()  the pinned third-party smplx==0.1.13 SMPLX.forward (requirements.txt:95;
                     call sites fitting_habitat.py:57-71,126-129); in-tree mirror of the
                     concatenation order / translation: body_model.py:223-247.  Hand PCA
                     and pose_mean are stated from the published smplx 0.1.x source --
                     PARITY UNPINNED for those two details (no in-tree test or fixture).
  * rot6d / aa conversions  source/cvae.py:46-55,117-137 on top of torchgeometry==0.1.2
                     (requirements.txt:103, third-party, absent) -- PARITY UNPINNED, restated
                     from its published algorithm.
  * cal_loss()       source/fitting_habitat.py:103-164 with `.cuda()` removed.

Pins: LBS is checked against the reference's own lbs.py imported by file path
(tests/golden/make_golden.py generated tests/golden/lbs_*.npz here; /root/reference does
not travel to the GPU box).  SDF is checked against torch's F.grid_sample with
align_corners=True (torch 1.2 default).  Chamfer is checked on the GPU box
against the reference chamfer.cu built unmodified into oracle/_ref/.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpsi_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libpsi_oracle.so missing: run `make -C oracle`")
        L = ctypes.CDLL(path)
        fp, ip, lg, it = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int),
                          ctypes.c_long, ctypes.c_int)
        L.psi_oracle_nn_fwd.argtypes = [fp, lg, it, it, fp, lg, it, fp, ip, it]
        L.psi_oracle_chamfer_fwd.argtypes = [fp, fp, it, it, it, fp, ip, fp, ip]
        L.psi_oracle_chamfer_bwd.argtypes = [fp, fp, it, it, it, fp, ip, fp, ip, fp, fp]
        L.psi_oracle_sdf_fwd.argtypes = [fp, it, fp, fp, fp, lg, fp, fp]
        L.psi_oracle_set_threads.argtypes = [it]
        L.psi_oracle_set_threads.restype = None
        for f in (L.psi_oracle_nn_fwd, L.psi_oracle_chamfer_fwd, L.psi_oracle_chamfer_bwd,
                  L.psi_oracle_sdf_fwd, L.psi_oracle_num_threads, L.psi_oracle_simd_width):
            f.restype = it
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads() -> int:
    return int(lib().psi_oracle_num_threads())


def set_threads(n: int) -> None:
    lib().psi_oracle_set_threads(int(n))


def simd_width() -> int:
    return int(lib().psi_oracle_simd_width())


def nn_fwd(q, s, scalar: bool = False):
    """q [B,n,3], s [B,m,3] or [m,3] (shared) -> (dist [B,n] f32, idx [B,n] i32)."""
    q = _f32(q)
    s = _f32(s)
    B, n, _ = q.shape
    shared = s.ndim == 2
    m = s.shape[-2]
    dist = np.zeros((B, n), dtype=np.float32)
    idx = np.zeros((B, n), dtype=np.int32)
    rc = lib().psi_oracle_nn_fwd(_fp(q), n * 3, B, n, _fp(s), 0 if shared else m * 3, m,
                                 _fp(dist), _ip(idx), 1 if scalar else 0)
    assert rc == 0, rc
    return dist, idx


def chamfer_fwd(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((B, n), np.float32)
    d2 = np.zeros((B, m), np.float32)
    i1 = np.zeros((B, n), np.int32)
    i2 = np.zeros((B, m), np.int32)
    rc = lib().psi_oracle_chamfer_fwd(_fp(xyz1), _fp(xyz2), B, n, m, _fp(d1), _ip(i1), _fp(d2), _ip(i2))
    assert rc == 0, rc
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, gd1, gd2, idx1, idx2):
    xyz1, xyz2, gd1, gd2 = _f32(xyz1), _f32(xyz2), _f32(gd1), _f32(gd2)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.zeros_like(xyz1)
    g2 = np.zeros_like(xyz2)
    rc = lib().psi_oracle_chamfer_bwd(_fp(xyz1), _fp(xyz2), B, n, m, _fp(gd1), _ip(idx1),
                                      _fp(gd2), _ip(idx2), _fp(g1), _fp(g2))
    assert rc == 0, rc
    return g1, g2


def sdf_fwd(sdf, gmin, gmax, verts, want_grad: bool = True):
    """sdf [D,D,D]; verts [...,3] scene frame -> (values [...], grad [...,3] or None)."""
    sdf, gmin, gmax, verts = _f32(sdf), _f32(gmin), _f32(gmax), _f32(verts)
    D = sdf.shape[0]
    flat = verts.reshape(-1, 3)
    out = np.zeros(flat.shape[0], np.float32)
    grad = np.zeros_like(flat) if want_grad else None
    rc = lib().psi_oracle_sdf_fwd(_fp(sdf), D, _fp(gmin), _fp(gmax), _fp(flat), flat.shape[0],
                                  _fp(out), _fp(grad) if want_grad else None)
    assert rc == 0, rc
    return out.reshape(verts.shape[:-1]), (grad.reshape(verts.shape) if want_grad else None)


# --------------------------------------------------------------------------- torch restatements
def sdf_lookup_torch(sdf, gmin, gmax, verts):
    """fitting_habitat.py:145-152 with the torch-1.2 default spelled out (align_corners=True).
    sdf [S?,D,D,D] or [D,D,D] (one grid for the whole batch), verts [B,V,3] -> [B,V]."""
    B, V, _ = verts.shape
    norm = (verts - gmin.view(1, 1, 3)) / (gmax.view(1, 1, 3) - gmin.view(1, 1, 3)) * 2 - 1
    grid = sdf.view(1, 1, *sdf.shape[-3:]).expand(B, -1, -1, -1, -1)
    out = F.grid_sample(grid, norm[:, :, [2, 1, 0]].view(-1, V, 1, 1, 3),
                        padding_mode="border", align_corners=True)
    return out.view(B, V)


def batch_rodrigues(aa):
    """lbs.py:165-192 -- note eps is added to the VECTOR before the norm (:177)."""
    angle = torch.norm(aa + 1e-8, dim=1, keepdim=True)
    n = aa / angle
    c = torch.cos(angle).unsqueeze(1)
    s = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = n[:, 0:1], n[:, 1:2], n[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """lbs.py:34-118.  betas [B,NB], pose [B,J*3], v_template [V,3], shapedirs [V,3,NB],
    posedirs [P,V*3], J_regressor [J,V], parents [J], lbs_weights [V,J] -> verts [B,V,3], joints."""
    B = betas.shape[0]
    nj = J_regressor.shape[0]
    v_shaped = v_template.unsqueeze(0) + torch.einsum("bl,mkl->bmk", betas, shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor).contiguous()
    R = batch_rodrigues(pose.reshape(-1, 3)).view(B, nj, 3, 3)
    pose_feature = (R[:, 1:] - torch.eye(3, dtype=betas.dtype, device=betas.device)).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    # kinematic chain (lbs.py:207-262)
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    Tm = torch.zeros(B, nj, 4, 4, dtype=betas.dtype, device=betas.device)
    Tm[:, :, :3, :3] = R
    Tm[:, :, :3, 3] = rel
    Tm[:, :, 3, 3] = 1
    chain = [Tm[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], Tm[:, i]))
    G = torch.stack(chain, dim=1)
    posed_joints = G[:, :, :3, 3]
    Jh = torch.cat([J, torch.zeros(B, nj, 1, dtype=betas.dtype, device=betas.device)], dim=2).unsqueeze(-1)
    init_bone = F.pad(torch.matmul(G, Jh), [3, 0])
    A = G - init_bone
    T = torch.matmul(lbs_weights.unsqueeze(0).expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype, device=betas.device)], dim=2)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed_joints


class SMPLXOracle:
    """smplx.create(..., num_pca_comps=12)(...) restated on CPU tensors."""

    def __init__(self, model: dict, num_betas=10, num_expression=10, num_pca_comps=12,
                 dtype=torch.float32):
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        sd = np.asarray(model["shapedirs"])
        begin = 300 if sd.shape[-1] > 300 else 10           # body_model.py:103-105
        self.shapedirs = t(np.concatenate([sd[:, :, :num_betas], sd[:, :, begin:begin + num_expression]], -1))
        self.v_template = t(model["v_template"])
        pd = np.asarray(model["posedirs"])
        self.posedirs = t(pd.reshape(pd.shape[0] * 3, -1).T)  # body_model.py:123-125
        self.J_regressor = t(model["J_regressor"])
        self.weights = t(model["weights"])
        parents = np.asarray(model["kintree_table"])[0].astype(np.int64).copy()
        parents[0] = -1
        self.parents = parents
        self.lh = t(np.asarray(model["hands_componentsl"])[:num_pca_comps])
        self.rh = t(np.asarray(model["hands_componentsr"])[:num_pca_comps])
        nj = self.J_regressor.shape[0]
        mean = np.zeros(nj * 3, dtype=np.float32)
        mean[(nj - 30) * 3:(nj - 15) * 3] = np.asarray(model["hands_meanl"], dtype=np.float32)
        mean[(nj - 15) * 3:] = np.asarray(model["hands_meanr"], dtype=np.float32)
        self.pose_mean = t(mean)
        self.dtype = dtype
        self.num_expression = num_expression

    def to(self, device):
        """The same restatement with its constants on `device` (oracle/reference_gpu.py: the reference's
        GPU path timed on the B200 next to ours)."""
        for k in ("shapedirs", "v_template", "posedirs", "J_regressor", "weights", "lh", "rh", "pose_mean"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def full_pose(self, global_orient, body_pose, left_hand_pose, right_hand_pose):
        B = global_orient.shape[0]
        z3 = torch.zeros(B, 3, dtype=self.dtype, device=global_orient.device)
        lh = torch.einsum("bi,ij->bj", left_hand_pose, self.lh)
        rh = torch.einsum("bi,ij->bj", right_hand_pose, self.rh)
        full = torch.cat([global_orient, body_pose, z3, z3, z3, lh, rh], dim=1)
        return full + self.pose_mean

    def __call__(self, body_pose, transl, global_orient, betas, left_hand_pose, right_hand_pose):
        B = betas.shape[0]
        full = self.full_pose(global_orient, body_pose, left_hand_pose, right_hand_pose)
        shape = torch.cat([betas, torch.zeros(B, self.num_expression, dtype=self.dtype, device=betas.device)], dim=-1)
        verts, joints = lbs(shape, full, self.v_template, self.shapedirs, self.posedirs,
                            self.J_regressor, self.parents, self.weights)
        return verts + transl.unsqueeze(1), joints + transl.unsqueeze(1)


# --- rotation representation chain (cvae.py:46-55,117-137 over torchgeometry 0.1.2) -------------
def rot6d_to_matrix(x6):
    x = x6.view(-1, 3, 2)
    b1 = F.normalize(x[:, :, 0], dim=1)
    dot = torch.sum(b1 * x[:, :, 1], dim=1, keepdim=True)
    b2 = F.normalize(x[:, :, 1] - dot * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1)


def rotation_matrix_to_quaternion(R, eps=1e-6):
    """torchgeometry 0.1.2 rotation_matrix_to_quaternion on [N,3,3] (its input is the 3x4 pad)."""
    rt = R.transpose(1, 2)
    m22, m00, m11 = rt[:, 2, 2], rt[:, 0, 0], rt[:, 1, 1]
    mask_d2 = m22 < eps
    mask_d0_d1 = m00 > m11
    mask_d0_nd1 = m00 < -m11
    t0 = 1 + m00 - m11 - m22
    q0 = torch.stack([rt[:, 1, 2] - rt[:, 2, 1], t0, rt[:, 0, 1] + rt[:, 1, 0], rt[:, 2, 0] + rt[:, 0, 2]], -1)
    t1 = 1 - m00 + m11 - m22
    q1 = torch.stack([rt[:, 2, 0] - rt[:, 0, 2], rt[:, 0, 1] + rt[:, 1, 0], t1, rt[:, 1, 2] + rt[:, 2, 1]], -1)
    t2 = 1 - m00 - m11 + m22
    q2 = torch.stack([rt[:, 0, 1] - rt[:, 1, 0], rt[:, 2, 0] + rt[:, 0, 2], rt[:, 1, 2] + rt[:, 2, 1], t2], -1)
    t3 = 1 + m00 + m11 + m22
    q3 = torch.stack([t3, rt[:, 1, 2] - rt[:, 2, 1], rt[:, 2, 0] - rt[:, 0, 2], rt[:, 0, 1] - rt[:, 1, 0]], -1)
    c0 = (mask_d2 & mask_d0_d1).to(R.dtype).unsqueeze(-1)
    c1 = (mask_d2 & ~mask_d0_d1).to(R.dtype).unsqueeze(-1)
    c2 = (~mask_d2 & mask_d0_nd1).to(R.dtype).unsqueeze(-1)
    c3 = (~mask_d2 & ~mask_d0_nd1).to(R.dtype).unsqueeze(-1)
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0.unsqueeze(-1) * c0 + t1.unsqueeze(-1) * c1 + t2.unsqueeze(-1) * c2 + t3.unsqueeze(-1) * c3)
    return q * 0.5


def quaternion_to_angle_axis(q):
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    sin_sq = q1 * q1 + q2 * q2 + q3 * q3
    sin_t = torch.sqrt(sin_sq)
    cos_t = q[..., 0]
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(sin_sq > 0.0, two_theta / sin_t, 2.0 * torch.ones_like(sin_t))
    return torch.stack([q1 * k, q2 * k, q3 * k], dim=-1)


def matrix_to_aa(R):
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(R))


def aa_to_matrix(aa, eps=1e-6):
    """torchgeometry 0.1.2 angle_axis_to_rotation_matrix (Taylor branch for small angles)."""
    theta2 = torch.sum(aa * aa, dim=1)
    theta = torch.sqrt(theta2)
    w = aa / (theta.unsqueeze(1) + eps)
    wx, wy, wz = w[:, 0], w[:, 1], w[:, 2]
    c, s = torch.cos(theta), torch.sin(theta)
    k1 = 1.0
    normal = torch.stack([
        c + wx * wx * (k1 - c), wx * wy * (k1 - c) - wz * s, wy * s + wx * wz * (k1 - c),
        wz * s + wx * wy * (k1 - c), c + wy * wy * (k1 - c), -wx * s + wy * wz * (k1 - c),
        -wy * s + wx * wz * (k1 - c), wx * s + wy * wz * (k1 - c), c + wz * wz * (k1 - c)], dim=1).view(-1, 3, 3)
    rx, ry, rz = aa[:, 0], aa[:, 1], aa[:, 2]
    one = torch.ones_like(rx)
    taylor = torch.stack([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)
    mask = (theta2 > eps).view(-1, 1, 1).to(aa.dtype)
    return mask * normal + (1 - mask) * taylor


def convert_to_6D_rot(x):
    """cvae.py:117-126: [t3 | aa3 | rest] -> [t3 | 6D | rest]."""
    R = aa_to_matrix(x[:, 3:6])
    return torch.cat([x[:, :3], R[:, :, :-1].reshape(-1, 6), x[:, 6:]], dim=-1)


def convert_to_3D_rot(x):
    """cvae.py:128-137: [t3 | 6D | rest] -> [t3 | aa3 | rest]."""
    return torch.cat([x[:, :3], matrix_to_aa(rot6d_to_matrix(x[:, 3:9])), x[:, 9:]], dim=-1)


class VPoserDecoderOracle:
    """VPoser.decode(z, output_type='aa') (vposer_smpl.py:107-121), eval mode (dropout off)."""

    def __init__(self, weights: dict):
        self.w = {k: torch.tensor(v) for k, v in weights.items()}

    def to(self, device):
        self.w = {k: v.to(device) for k, v in self.w.items()}
        return self

    def decode(self, z):
        w = self.w
        x = F.leaky_relu(F.linear(z, w["bodyprior_dec_fc1.weight"], w["bodyprior_dec_fc1.bias"]), 0.2)
        x = F.leaky_relu(F.linear(x, w["bodyprior_dec_fc2.weight"], w["bodyprior_dec_fc2.bias"]), 0.2)
        x = F.linear(x, w["bodyprior_dec_out.weight"], w["bodyprior_dec_out.bias"])
        R = rot6d_to_matrix(x)                       # [B*21,3,3]
        return matrix_to_aa(R).view(z.shape[0], -1)  # [B,63]


def verts_transform(verts, cam_ext):
    """cvae.py:141-149."""
    vh = F.pad(verts, (0, 1), mode="constant", value=1)
    return torch.matmul(vh, cam_ext.permute(0, 2, 1))[:, :, :-1]


class _NNFunction(torch.autograd.Function):
    """dist1 of chamferDist (one direction; the fitting loss discards dist2,
    fitting_habitat.py:138) over the C oracle, with the reference backward for xyz1."""

    @staticmethod
    def forward(ctx, q, s):
        d, i = nn_fwd(q.detach().numpy(), s.detach().numpy())
        ctx.save_for_backward(q, s, torch.from_numpy(i))
        return torch.from_numpy(d)

    @staticmethod
    def backward(ctx, g):
        q, s, i = ctx.saved_tensors
        sb = s if s.dim() == 3 else s.unsqueeze(0).expand(q.shape[0], -1, -1)
        nearest = torch.gather(sb, 1, i.long().unsqueeze(-1).expand(-1, -1, 3))
        return (g * 2).unsqueeze(-1) * (q - nearest), None


def cal_loss(xhr, xhr_rec, cam_ext, smplx_model: SMPLXOracle, vposer: VPoserDecoderOracle,
             sdf, gmin, gmax, scene_points, contact_ids, weights: dict, robust_c: float = 1.0,
             loss_mode: str = "batch"):
    """fitting_habitat.py:103-164 on CPU tensors.  Returns the four weighted loss terms.
    loss_mode 'batch' = the reference code as written (means over the whole batch, the
    collision mean over all negative entries of the batch);
    'independent' = the SUM over bodies of that same loss evaluated at B=1, i.e. what the
    shipped scripts compute body by body (batch_size 1, fitting_habitat.py:254).  Both coincide
    at B=1."""
    B = xhr_rec.shape[0]
    indep = loss_mode == "independent"
    red = (lambda t: t.reshape(B, -1).mean(dim=1).sum()) if indep else (lambda t: t.mean())
    loss_rec = weights["weight_loss_rec"] * red((xhr - xhr_rec).abs())
    xh = convert_to_3D_rot(xhr_rec)
    z = xh[:, 16:48]
    loss_vposer = weights["weight_loss_vposer"] * red(z ** 2)
    body_pose = vposer.decode(z)
    verts, _ = smplx_model(body_pose=body_pose, transl=xh[:, :3], global_orient=xh[:, 3:6],
                           betas=xh[:, 6:16], left_hand_pose=xh[:, 48:60], right_hand_pose=xh[:, 60:])
    verts = verts_transform(verts, cam_ext)
    contact = verts[:, torch.as_tensor(contact_ids, dtype=torch.long), :].contiguous()
    d = _NNFunction.apply(contact, scene_points)
    s = torch.sqrt(d + 1e-4)
    loss_contact = weights["weight_contact"] * red(s / (s + robust_c))
    body_sdf = sdf_lookup_torch(sdf, gmin, gmax, verts)
    neg = body_sdf < 0
    if not indep:
        if int(neg.sum()) < 1:
            pene = torch.tensor(0.0)
        else:
            pene = body_sdf[neg].abs().mean()
    else:
        cnt = neg.sum(dim=1).clamp(min=1).to(body_sdf.dtype)
        pene = ((-body_sdf * neg).sum(dim=1) / cnt).sum()
    loss_collision = weights["weight_collision"] * pene
    return loss_rec, loss_vposer, loss_contact, loss_collision


def fit_loop(xh, cam_ext, num_iter, lr, smplx_model, vposer, sdf, gmin, gmax, scene_points,
             contact_ids, weights, robust_c=1.0, loss_mode="independent"):
    """fitting_habitat.py:169-197 on the CPU: Adam(lr) on the 75-D vector, `num_iter` iterations.
    xh [B,72] -> fitted [B,72]."""
    xhr = convert_to_6D_rot(xh)
    xhr_rec = xhr.clone().requires_grad_(True)
    opt = torch.optim.Adam([xhr_rec], lr=lr)
    for _ in range(num_iter):
        opt.zero_grad()
        terms = cal_loss(xhr, xhr_rec, cam_ext, smplx_model, vposer, sdf, gmin, gmax, scene_points,
                         contact_ids, weights, robust_c, loss_mode)
        sum(terms).backward()
        opt.step()
    return convert_to_3D_rot(xhr_rec.detach())

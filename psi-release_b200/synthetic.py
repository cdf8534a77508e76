"""Deterministic synthetic world for the PSI fitting hot path.

Nothing the reference needs at run time is in its tree: no SMPL-X
model file, no VPoser checkpoint, no PROX-E / MP3D-R scenes.  Everything here is
generated from fixed seeds with numpy's PCG64 generator so that the CPU box, the
GPU box and the committed golden fixtures all see the same bytes.

Shapes and npz key names follow what the reference loads:
  * SMPL-X npz keys  -- /root/reference/human_body_prior/body_model/body_model.py:86-135
    (v_template, shapedirs, posedirs, J_regressor, kintree_table, weights, f) plus
    smplx's hands_componentsl/r, hands_meanl/r.
  * scene SDF  = {min[3], max[3], dim} + dim^3 float grid
    -- /root/reference/source/fitting_habitat.py:80-90
  * body parameter vector [transl 3 | global_orient 3 | betas 10 | vposer z 32 |
    lhand 12 | rhand 12] -- /root/reference/source/cvae.py:238-249
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np

# Public SMPL-X kinematic tree (55 joints: 22 body, jaw, 2 eyes, 2 x 15 hand).
SMPLX_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
     15, 15, 15,
     20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
     21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53], dtype=np.int64)

NUM_VERTS = 10475
NUM_JOINTS = 55
NUM_POSE_BASIS = (NUM_JOINTS - 1) * 9  # 486
NUM_FACES = 20908


def make_smplx_model(seed: int = 1234, num_verts: int = NUM_VERTS,
                     num_shape_file: int = 20, parents: np.ndarray | None = None,
                     max_skin_nnz: int = 4) -> dict:
    """SMPL-X-shaped model dictionary (float32 arrays, npz key names)."""
    rng = np.random.default_rng(seed)
    parents = SMPLX_PARENTS if parents is None else np.asarray(parents, dtype=np.int64)
    nj = len(parents)
    V = num_verts

    # a tree-shaped blob: every joint sits a short bone away from its parent
    jpos = np.zeros((nj, 3), dtype=np.float32)
    for j in range(1, nj):
        step = rng.standard_normal(3, dtype=np.float32) * np.float32(0.10)
        step[1] += np.float32(0.08) if j < 22 else np.float32(0.02)
        jpos[j] = jpos[parents[j]] + step
    jpos -= jpos.mean(axis=0, keepdims=True)

    owner = rng.integers(0, nj, size=V)
    v_template = (jpos[owner] + rng.standard_normal((V, 3), dtype=np.float32) * np.float32(0.05)).astype(np.float32)

    shapedirs = (rng.standard_normal((V, 3, num_shape_file), dtype=np.float32) * np.float32(0.01)).astype(np.float32)
    posedirs = (rng.standard_normal((V, 3, (nj - 1) * 9), dtype=np.float32) * np.float32(1e-3)).astype(np.float32)

    # skinning weights: owner joint dominant, the rest spread over its ancestors
    weights = np.zeros((V, nj), dtype=np.float32)
    w_raw = rng.random((V, max_skin_nnz), dtype=np.float32) + np.float32(0.05)
    w_raw[:, 0] += np.float32(1.5)
    w_raw /= w_raw.sum(axis=1, keepdims=True)
    cur = owner.copy()
    for k in range(max_skin_nnz):
        np.add.at(weights, (np.arange(V), cur), w_raw[:, k])
        nxt = parents[cur]
        cur = np.where(nxt < 0, cur, nxt)
    weights = (weights / weights.sum(axis=1, keepdims=True)).astype(np.float32)

    # joint regressor: ~20 vertices per joint, non-negative rows summing to one
    J_regressor = np.zeros((nj, V), dtype=np.float32)
    for j in range(nj):
        mine = np.nonzero(owner == j)[0]
        if len(mine) < 20:
            mine = np.concatenate([mine, rng.integers(0, V, size=20 - len(mine))])
        pick = rng.choice(mine, size=20, replace=False) if len(mine) > 20 else mine[:20]
        w = rng.random(20, dtype=np.float32) + np.float32(0.1)
        np.add.at(J_regressor[j], pick, w / w.sum())
    J_regressor = J_regressor.astype(np.float32)

    kintree = np.stack([parents.copy(), np.arange(nj)]).astype(np.int64)
    kintree[0, 0] = -1

    nf = NUM_FACES if V == NUM_VERTS else max(4, 2 * V - 4)
    faces = rng.integers(0, V, size=(nf, 3)).astype(np.uint32)

    def hand_basis():
        a = rng.standard_normal((45, 45))
        q, _ = np.linalg.qr(a)
        return q.T.astype(np.float32)  # orthonormal rows

    return {
        "v_template": v_template,
        "shapedirs": shapedirs,
        "posedirs": posedirs,
        "J_regressor": J_regressor,
        "kintree_table": kintree,
        "weights": weights,
        "f": faces,
        "hands_componentsl": hand_basis(),
        "hands_componentsr": hand_basis(),
        "hands_meanl": (rng.standard_normal(45, dtype=np.float32) * np.float32(0.1)).astype(np.float32),
        "hands_meanr": (rng.standard_normal(45, dtype=np.float32) * np.float32(0.1)).astype(np.float32),
    }


def write_smplx_model(model_path: str, model: dict | None = None, **kw) -> str:
    """Write <model_path>/smplx/SMPLX_NEUTRAL.npz (layout: train_s2.py:89-91)."""
    model = make_smplx_model(**kw) if model is None else model
    d = os.path.join(model_path, "smplx")
    os.makedirs(d, exist_ok=True)
    fn = os.path.join(d, "SMPLX_NEUTRAL.npz")
    np.savez(fn, **model)
    return fn


# ----------------------------------------------------------------------------- scene
@dataclass
class Scene:
    grid_min: np.ndarray   # [3] float32
    grid_max: np.ndarray   # [3] float32
    dim: int
    sdf: np.ndarray        # [dim,dim,dim] float32, index [x][y][z]
    points: np.ndarray     # [M,3] float32 on the zero level set
    cam_ext: np.ndarray    # [4,4] float32 camera -> scene rigid transform
    boxes: np.ndarray      # [K,6] solid boxes (centre, half extent); row 0 is the room (free inside)


def _sd_box(px, py, pz, c, h):
    qx = np.abs(px - c[0]) - h[0]
    qy = np.abs(py - c[1]) - h[1]
    qz = np.abs(pz - c[2]) - h[2]
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    inside = np.minimum(np.maximum(np.maximum(qx, qy), qz), 0)
    return (outside + inside).astype(np.float32)


def make_scene(seed: int = 0, dim: int = 256, num_points: int = 50000,
               extent: float = 3.0, num_furniture: int = 4) -> Scene:
    """Box room with box furniture: SDF > 0 in free space, < 0 inside solids."""
    rng = np.random.default_rng(10_000 + seed)
    gmin = np.full(3, -extent, dtype=np.float32)
    gmax = np.full(3, extent, dtype=np.float32)
    room_c = np.zeros(3, dtype=np.float32)
    room_h = np.array([2.4, 2.4, 1.4], dtype=np.float32)
    boxes = [np.concatenate([room_c, room_h])]
    for _ in range(num_furniture):
        h = rng.uniform([0.25, 0.25, 0.2], [0.8, 0.8, 0.6]).astype(np.float32)
        c = rng.uniform(-1.6, 1.6, size=3).astype(np.float32)
        c[2] = -room_h[2] + h[2]          # resting on the floor
        boxes.append(np.concatenate([c, h]).astype(np.float32))
    boxes = np.stack(boxes).astype(np.float32)

    ax = np.linspace(-extent, extent, dim, dtype=np.float32)
    px, py, pz = ax[:, None, None], ax[None, :, None], ax[None, None, :]
    sdf = -_sd_box(px, py, pz, boxes[0, :3], boxes[0, 3:])       # inside room = free (>0)
    sdf = np.broadcast_to(sdf, (dim, dim, dim)).copy()
    for b in boxes[1:]:
        np.minimum(sdf, _sd_box(px, py, pz, b[:3], b[3:]), out=sdf)
    sdf = np.ascontiguousarray(sdf, dtype=np.float32)

    # surface samples: faces of every box, area-weighted
    faces = []
    for b in boxes:
        c, h = b[:3], b[3:]
        for axis in range(3):
            o = [a for a in range(3) if a != axis]
            area = 4.0 * h[o[0]] * h[o[1]]
            for sgn in (-1.0, 1.0):
                faces.append((c, h, axis, sgn, area))
    areas = np.array([f[4] for f in faces], dtype=np.float64)
    which = rng.choice(len(faces), size=num_points, p=areas / areas.sum())
    uv = rng.uniform(-1.0, 1.0, size=(num_points, 3)).astype(np.float32)
    pts = np.empty((num_points, 3), dtype=np.float32)
    for i, (c, h, axis, sgn, _) in enumerate(faces):
        m = which == i
        p = c[None, :] + uv[m] * h[None, :]
        p[:, axis] = c[axis] + sgn * h[axis]
        pts[m] = p
    pts = pts.astype(np.float32)

    # camera -> scene rigid transform
    a = rng.standard_normal((3, 3))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))[None, :]
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    cam = np.eye(4, dtype=np.float32)
    cam[:3, :3] = q.astype(np.float32)
    cam[:3, 3] = rng.uniform(-0.5, 0.5, size=3).astype(np.float32)
    return Scene(gmin, gmax, dim, sdf, pts, cam, boxes)


def write_scene(prefix: str, scene: Scene) -> None:
    """<prefix>.json + <prefix>_sdf.npy + <prefix>_verts.npy (fitting_habitat.py:80-96)."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    with open(prefix + ".json", "w") as f:
        json.dump({"min": scene.grid_min.tolist(), "max": scene.grid_max.tolist(),
                   "dim": int(scene.dim)}, f)
    np.save(prefix + "_sdf.npy", scene.sdf.reshape(-1))
    np.save(prefix + "_verts.npy", scene.points)


# ----------------------------------------------------------------------------- bodies
def make_body_params(scene: Scene, batch: int, seed: int = 0) -> np.ndarray:
    """[batch,72] generator-style body vectors placed inside the room (camera frame)."""
    rng = np.random.default_rng(20_000 + seed)
    room_h = scene.boxes[0, 3:]
    p_scene = rng.uniform(-0.75, 0.75, size=(batch, 3)).astype(np.float32) * room_h[None, :]
    # half of the bodies are pushed towards the floor so that sdf < 0 is non-empty
    low = np.arange(batch) % 2 == 0
    p_scene[low, 2] = -room_h[2] + rng.uniform(-0.05, 0.25, size=int(low.sum())).astype(np.float32)
    R, t = scene.cam_ext[:3, :3], scene.cam_ext[:3, 3]
    transl = (p_scene - t[None, :]) @ R          # R^T (p - t)
    x = np.zeros((batch, 72), dtype=np.float32)
    x[:, 0:3] = transl
    x[:, 3:6] = rng.standard_normal((batch, 3)) * 0.5
    x[:, 6:16] = rng.standard_normal((batch, 10))
    x[:, 16:48] = rng.standard_normal((batch, 32))
    x[:, 48:72] = rng.standard_normal((batch, 24)) * 0.5
    return x.astype(np.float32)


def make_contact_ids(num_verts: int = NUM_VERTS, mode: str = "parts", seed: int = 5) -> np.ndarray:
    """Contact vertex ids. 'full' = every vertex (BASELINE's 10k-vert body);
    'parts' = 7 sets, ~2500 ids with a few duplicates across parts (cvae.py:99-115)."""
    if mode == "full":
        return np.arange(num_verts, dtype=np.int64)
    rng = np.random.default_rng(seed)
    sizes = [520, 410, 300, 300, 330, 330, 310]
    if num_verts < 4 * max(sizes):
        sizes = [max(1, num_verts // 12)] * 7
    parts = []
    for s in sizes:
        start = int(rng.integers(0, max(1, num_verts - 2 * s)))
        ids = start + rng.choice(2 * s, size=s, replace=False)
        parts.append(np.sort(ids))
    return np.concatenate(parts).astype(np.int64)


def make_vposer_weights(seed: int = 7, latent: int = 32, hidden: int = 512, out: int = 126) -> dict:
    """VPoser-decoder-shaped MLP 32->512->512->126 (vposer_smpl.py:83-89,110-113)."""
    rng = np.random.default_rng(seed)

    def lin(i, o):
        b = 1.0 / np.sqrt(i)
        return (rng.uniform(-b, b, size=(o, i)).astype(np.float32),
                rng.uniform(-b, b, size=(o,)).astype(np.float32))

    w1, b1 = lin(latent, hidden)
    w2, b2 = lin(hidden, hidden)
    w3, b3 = lin(hidden, out)
    return {"bodyprior_dec_fc1.weight": w1, "bodyprior_dec_fc1.bias": b1,
            "bodyprior_dec_fc2.weight": w2, "bodyprior_dec_fc2.bias": b2,
            "bodyprior_dec_out.weight": w3, "bodyprior_dec_out.bias": b3}

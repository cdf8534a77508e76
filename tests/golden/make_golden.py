"""Generate golden vectors from the REFERENCE's own code (run in the build container only).

    python tests/golden/make_golden.py

/root/reference does not travel to the GPU box, so the outputs are committed as small
fixtures next to this script:

  lbs_small.npz / lbs_full.npz
      reference `lbs()` (human_body_prior/body_model/lbs.py:34-118, imported BY FILE PATH,
      unmodified; B>1 needs `vertices2joints` made contiguous -- a torch>=1.5 stride issue
      at lbs.py:242) on the synthetic SMPL-X-shaped model from
      psi-release_b200/synthetic.py (seeded; regenerated at test time, not stored), with the
      smplx-style full pose / shape vectors stored as inputs.  Stored outputs: verts, posed
      joints and the autograd gradients of sum(verts * probe) w.r.t. betas and pose.
  chamfer_matmul_4x100.npz
      the reference test's independent formula (chamfer_pytorch/test_chamfer.py:22-32,
      ||x||^2+||y||^2-2xy, CPU form) on seeded rand(4,100,3) clouds: distances to 1e-8
      summed squared error are the only Chamfer results the reference's own test pins.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_ref_lbs():
    spec = importlib.util.spec_from_file_location(
        "ref_lbs", os.path.join(REF, "human_body_prior/body_model/lbs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    orig = mod.vertices2joints
    mod.vertices2joints = lambda J, v: orig(J, v).contiguous()
    return mod


def lbs_case(ref, name, num_verts, batch, seed):
    from psi_release_b200 import synthetic
    model = synthetic.make_smplx_model(seed=1234, num_verts=num_verts)
    rng = np.random.default_rng(seed)
    betas = rng.standard_normal((batch, 20)).astype(np.float32)
    betas[:, 10:] *= 0.3
    pose = (rng.standard_normal((batch, 165)) * 0.4).astype(np.float32)
    pose[0, 3:9] = 0.0          # exercise the tiny-angle Rodrigues path (eps inside the norm)
    probe = rng.standard_normal((batch, num_verts, 3)).astype(np.float32)

    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    sd = model["shapedirs"]
    shapedirs = t(np.concatenate([sd[:, :, :10], sd[:, :, 10:20]], -1))
    pd = model["posedirs"]
    posedirs = t(pd.reshape(pd.shape[0] * 3, -1).T)
    parents = torch.tensor(model["kintree_table"][0].astype(np.int64))
    parents[0] = -1
    b = t(betas).requires_grad_(True)
    p = t(pose).requires_grad_(True)
    verts, joints = ref.lbs(betas=b, pose=p, v_template=t(model["v_template"]).unsqueeze(0).repeat(batch, 1, 1),
                            shapedirs=shapedirs, posedirs=posedirs,
                            J_regressor=t(model["J_regressor"]), parents=parents,
                            lbs_weights=t(model["weights"]), num_joints=55)
    (verts * t(probe)).sum().backward()
    np.savez_compressed(os.path.join(HERE, name), model_seed=1234, num_verts=num_verts,
                        betas=betas, pose=pose, probe_seed=seed,
                        verts=verts.detach().numpy(), joints=joints.detach().numpy(),
                        grad_betas=b.grad.numpy(), grad_pose=p.grad.numpy())
    print(name, verts.shape, float(verts.abs().max()))


def chamfer_matmul_case():
    g = torch.Generator().manual_seed(4100)
    a = torch.rand(4, 100, 3, generator=g)
    b = torch.rand(4, 100, 3, generator=g)
    xx = torch.bmm(a, a.transpose(2, 1))
    yy = torch.bmm(b, b.transpose(2, 1))
    zz = torch.bmm(a, b.transpose(2, 1))
    di = torch.arange(0, 100)
    rx = xx[:, di, di].unsqueeze(1).expand_as(xx)
    ry = yy[:, di, di].unsqueeze(1).expand_as(yy)
    P = rx.transpose(2, 1) + ry - 2 * zz
    np.savez_compressed(os.path.join(HERE, "chamfer_matmul_4x100.npz"), a=a.numpy(), b=b.numpy(),
                        dist1=P.min(2)[0].numpy(), dist2=P.min(1)[0].numpy())
    print("chamfer_matmul_4x100.npz")


if __name__ == "__main__":
    assert os.path.isdir(REF), "reference tree not present: goldens can only be regenerated in the build container"
    ref = load_ref_lbs()
    lbs_case(ref, "lbs_small.npz", 431, 3, 11)
    lbs_case(ref, "lbs_full.npz", 10475, 2, 12)
    chamfer_matmul_case()

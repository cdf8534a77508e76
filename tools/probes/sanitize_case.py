"""A tiny pass over every kernel family of libpsi_b200 (the fused loop in its eager, replayed and whole-loop
forms, batch-coupled loss, the generic LBS / SDF / NN / Chamfer operators) -- small enough to run under
compute-sanitizer (tests/test_gpu_sanitizer.py, profiles/*_sanitizer.txt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from psi_release_b200 import chamfer, synthetic  # noqa: E402
from psi_release_b200.fitting import FittingOP  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
model = synthetic.make_smplx_model(seed=1234, num_verts=431)
scene = synthetic.make_scene(seed=1, dim=16, num_points=700)
B = 3
xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
base = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
            contact_ids=synthetic.make_contact_ids(431, "parts"), init_lr_h=0.1, num_iter=2, batch_size=B, device="cuda")
parts = sys.argv[1].split(",") if len(sys.argv) > 1 else ["eager", "replay", "whole", "batch", "autograd", "bruteforce", "chamfer"]
cases = {"eager": dict(use_cuda_graph=False), "replay": dict(loop_mode="replay"), "whole": dict(loop_mode="whole"),
         "batch": dict(loss_mode="batch", use_cuda_graph=False), "autograd": dict(engine="autograd", use_cuda_graph=False),
         "bruteforce": dict(engine="autograd", nn="bruteforce", use_cuda_graph=False)}
outs = {}
for name in parts:
    if name in cases:
        op = FittingOP(dict(base, **cases[name]), W)
        outs[name] = op.fit(xh, cam, num_iter=17 if name == "whole" else 2)
if "chamfer" in parts:
    a, b = torch.rand(2, 70, 3, device="cuda", requires_grad=True), torch.rand(2, 90, 3, device="cuda", requires_grad=True)
    d1, d2 = chamfer.chamferDist()(a, b)
    (d1.sum() + d2.sum()).backward()
torch.cuda.synchronize()
assert all(torch.isfinite(o).all() for o in outs.values())
if "eager" in outs and "replay" in outs:
    assert torch.equal(outs["eager"], outs["replay"])
print("sanitize_case ok:", ",".join(parts))

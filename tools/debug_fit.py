import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psi_release_b200 import synthetic
from psi_release_b200.fitting import FittingOP
W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
small = synthetic.make_smplx_model(seed=1234, num_verts=431)
for contact in ("parts", "full"):
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    xh = synthetic.make_body_params(scene, 3, seed=3)
    cid = synthetic.make_contact_ids(431, contact)
    cfg = dict(model_data=small, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
               contact_ids=cid, init_lr_h=0.1, num_iter=4, batch_size=3, device="cuda")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    fused = FittingOP(dict(cfg, engine="fused", num_iter=6), W)
    auto = FittingOP(dict(cfg, engine="autograd", num_iter=6), W)
    f1 = fused.fit(torch.tensor(xh).cuda(), cam.cuda(), num_iter=1)
    a1 = auto.fit(torch.tensor(xh).cuda(), cam.cuda(), num_iter=1)
    d = (f1 - a1).abs()
    print(contact, "max", float(d.max()), "at", np.unravel_index(int(d.argmax()), d.shape))
    print((d > 1e-4).nonzero().tolist())
    print("fused", f1[d > 1e-4].tolist(), "auto", a1[d > 1e-4].tolist(), "x0", torch.tensor(xh).cuda()[d > 1e-4].tolist())

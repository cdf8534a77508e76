import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle
from psi_release_b200 import synthetic
from psi_release_b200.fitting import FittingOP
from psi_release_b200.geometry import GeometryTransformer
W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
model = synthetic.make_smplx_model(seed=1234, num_verts=431)
scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
t = torch.tensor
def cpu_grad(x0, x_eval, cam, kw, B):
    xr = x_eval.clone().requires_grad_(True)
    terms = oracle.cal_loss(x0, xr, cam.expand(B, -1, -1), loss_mode="independent", **kw)
    (go,) = torch.autograd.grad(sum(terms), xr)
    return go
for B, contact in ((65, "full"), (65, "parts"), (66, "full"), (129, "full")):
    xh = t(synthetic.make_body_params(scene, B, seed=3))
    cid = synthetic.make_contact_ids(431, contact)
    cam = t(scene.cam_ext).unsqueeze(0)
    kw = dict(smplx_model=oracle.SMPLXOracle(model), vposer=oracle.VPoserDecoderOracle(synthetic.make_vposer_weights()), sdf=t(scene.sdf),
              gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points), contact_ids=cid, weights=W)
    cfg = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(), contact_ids=cid, init_lr_h=0.1, num_iter=3,
               batch_size=B, device="cuda", engine="fused")
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()
    for name, over in (("graph", {}), ("eager", dict(use_cuda_graph=False))):
        op = FittingOP(dict(cfg, **over), W)
        for k in (2, 3, 4):
            op.fit(xh.cuda(), cam.cuda(), num_iter=k)
            x_eval = op.trace("x_eval").cpu(); g = op.trace("grad_x").cpu()
            go = cpu_grad(x0, x_eval, cam, kw, B)
            err = (g - go).abs().amax(1) / go.abs().amax(1)
            bad = torch.nonzero(err > 2e-4).flatten().tolist()
            print("B=%d %s %s k=%d: worst %.2e bad bodies %s" % (B, contact, name, k, float(err.max()), bad[:8]))
            if bad and k == 3 and name == "graph":
                # a fresh context evaluating ONE iteration at x_eval (x0 := x_eval: no rec gradient on either side)
                xh_eval = GeometryTransformer.convert_to_3D_rot(x_eval.cuda())
                op2 = FittingOP(cfg, W)
                op2.fit(xh_eval, cam.cuda(), num_iter=1)
                xe2 = op2.trace("x_eval").cpu(); g2 = op2.trace("grad_x").cpu()
                go2 = cpu_grad(xe2, xe2, cam, kw, B)
                err2 = (g2 - go2).abs().amax(1) / go2.abs().amax(1)
                print("    fresh single iteration at (approximately) that point: worst %.2e, bad %s; |xe2-x_eval| %.2e"
                      % (float(err2.max()), torch.nonzero(err2 > 2e-4).flatten().tolist()[:8], float((xe2 - x_eval).abs().max())))
                # per-body loss terms
                lg = op.trace("losses").cpu()
                for b in bad[:2]:
                    xr = x_eval[b:b+1].clone()
                    tb = oracle.cal_loss(x0[b:b+1], xr, cam, loss_mode="independent", **kw)
                    print("    body %d losses fused %s cpu %s" % (b, [round(float(v), 7) for v in lg[b]], [round(float(v), 7) for v in tb]))
                    sv = op.trace("sdf").cpu()[b]; print("    neg count", int((sv < 0).sum()), "nn_dist sum", float(op.trace("nn_dist").cpu()[b].sum()))

"""Diagnostic: per-component dL/dx of the fused loop vs the CPU oracle and vs the autograd engine."""
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
for B in (64, 65):
    xh = t(synthetic.make_body_params(scene, B, seed=3))
    cid = synthetic.make_contact_ids(431, "full")
    cam = t(scene.cam_ext).unsqueeze(0)
    kw = dict(smplx_model=oracle.SMPLXOracle(model), vposer=oracle.VPoserDecoderOracle(synthetic.make_vposer_weights()), sdf=t(scene.sdf),
              gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points), contact_ids=cid, weights=W)
    cfg = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(), contact_ids=cid, init_lr_h=0.1, num_iter=3,
               batch_size=B, device="cuda", engine="fused")
    op = FittingOP(cfg, W)
    auto = FittingOP(dict(cfg, engine="autograd", use_cuda_graph=False), W)
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()
    for k in (1, 3):
        op.fit(xh.cuda(), cam.cuda(), num_iter=k)
        x_eval = op.trace("x_eval").cpu()
        g = op.trace("grad_x").cpu()
        xr = x_eval.clone().requires_grad_(True)
        terms = oracle.cal_loss(x0, xr, cam.expand(B, -1, -1), loss_mode="independent", **kw)
        (go,) = torch.autograd.grad(sum(terms), xr)
        # autograd engine on the GPU at the same point
        with torch.no_grad():
            auto.xhr_rec.copy_(x_eval.cuda())
        at = auto.cal_loss(x0.cuda(), cam.cuda())
        (ga,) = torch.autograd.grad(sum(at), auto.xhr_rec)
        ga = ga.cpu()
        err = (g - go).abs().amax(1) / go.abs().amax(1)
        erra = (g - ga).abs().amax(1) / ga.abs().amax(1)
        erroa = (go - ga).abs().amax(1) / go.abs().amax(1)
        worst = int(err.argmax())
        c = int((g[worst] - go[worst]).abs().argmax())
        print("B=%d k=%d: fused-vs-cpu worst %.2e (body %d comp %d: fused %.6f cpu %.6f auto %.6f), fused-vs-auto worst %.2e, cpu-vs-auto worst %.2e"
              % (B, k, float(err.max()), worst, c, float(g[worst, c]), float(go[worst, c]), float(ga[worst, c]), float(erra.max()), float(erroa.max())))
        bad = torch.nonzero(err > 2e-4).flatten().tolist()
        print("   bodies over 2e-4:", bad[:10], "their errs", [round(float(err[b]), 5) for b in bad[:10]])
        verts = op.trace("verts").cpu().view(B, -1, 3)
        sv = op.trace("sdf").cpu().numpy()
        xh_e = oracle.convert_to_3D_rot(x_eval)
        vo, _ = kw["smplx_model"](body_pose=kw["vposer"].decode(xh_e[:, 16:48]), transl=xh_e[:, :3], global_orient=xh_e[:, 3:6],
                                  betas=xh_e[:, 6:16], left_hand_pose=xh_e[:, 48:60], right_hand_pose=xh_e[:, 60:])
        vo = oracle.verts_transform(vo, cam.expand(B, -1, -1))
        svc, _ = oracle.sdf_fwd(scene.sdf, scene.grid_min, scene.grid_max, vo.numpy(), want_grad=False)
        for b in bad[:4]:
            flips = np.nonzero((sv[b] < 0) != (svc[b] < 0))[0]
            print("   body %d: neg gpu %d cpu %d, sign flips at %s, min|sdf| %.2e, comps over: %s" % (
                b, int((sv[b] < 0).sum()), int((svc[b] < 0).sum()), flips.tolist()[:5], float(np.abs(svc[b]).min()),
                torch.nonzero((g[b] - go[b]).abs() > 2e-4 * go[b].abs().max()).flatten().tolist()[:12]))

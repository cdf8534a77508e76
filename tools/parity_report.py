"""Measured deviations of the fused fitting loop from the CPU oracle at BASELINE configs[1] size, as one JSON object
(committed under profiles/ next to the tests that assert the bounds: tests/test_gpu_trace.py).

    python tools/parity_report.py > profiles/rNN_parity_report.json        (needs a GPU)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (a measurement tool, like the tests: the checker, never the product)
from psi_release_b200 import synthetic  # noqa: E402
from psi_release_b200.fitting import FittingOP  # noqa: E402
from psi_release_b200.geometry import GeometryTransformer  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)


def main():
    B, V, M, D = 64, 10475, 50000, 256
    model = synthetic.make_smplx_model(seed=1234, num_verts=V)
    scene = synthetic.make_scene(seed=0, dim=D, num_points=M)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=0))
    cid = synthetic.make_contact_ids(V, "full")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    vw = synthetic.make_vposer_weights()
    op = FittingOP(dict(model_data=model, scene=scene, vposer_weights=vw, contact_ids=cid, init_lr_h=0.1, num_iter=3,
                        batch_size=B, device="cuda"), W)
    t = torch.tensor
    kw = dict(smplx_model=oracle.SMPLXOracle(model), vposer=oracle.VPoserDecoderOracle(vw), sdf=t(scene.sdf),
              gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points), contact_ids=cid, weights=W)
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()      # the loop's own starting vector
    out = {"config": "B=64, V=10475, D=256, M=50000, full contact, Adam lr 0.1", "iterations": {}}
    for k in (1, 3, 10):
        op.fit(xh.cuda(), cam.cuda(), num_iter=k)
        x_eval = op.trace("x_eval").cpu()
        verts = op.trace("verts").cpu().view(B, V, 3)
        xh_e = oracle.convert_to_3D_rot(x_eval)
        t0 = time.time()
        vo, _ = kw["smplx_model"](body_pose=kw["vposer"].decode(xh_e[:, 16:48]), transl=xh_e[:, :3], global_orient=xh_e[:, 3:6],
                                  betas=xh_e[:, 6:16], left_hand_pose=xh_e[:, 48:60], right_hand_pose=xh_e[:, 60:])
        vo = oracle.verts_transform(vo, cam.expand(B, -1, -1))
        qid = op.trace("query_ids").cpu().long()
        do, io = oracle.nn_fwd(verts[:, qid].contiguous().numpy(), scene.points)
        nnd, nni = op.trace("nn_dist").cpu().numpy(), op.trace("nn_idx").cpu().numpy()
        svo, sgo = oracle.sdf_fwd(scene.sdf, scene.grid_min, scene.grid_max, verts.numpy())
        sv, sg = op.trace("sdf").cpu().numpy(), op.trace("sdf_grad").cpu().numpy().reshape(B, V, 3)
        xr = x_eval.clone().requires_grad_(True)
        terms = oracle.cal_loss(x0, xr, cam.expand(B, -1, -1), loss_mode="independent", **kw)
        (go,) = torch.autograd.grad(sum(terms), xr)
        g = op.trace("grad_x").cpu()
        per_body = ((g - go).abs().amax(1) / go.abs().amax(1)).numpy()
        svc, _ = oracle.sdf_fwd(scene.sdf, scene.grid_min, scene.grid_max, vo.numpy(), want_grad=False)
        _, ioc = oracle.nn_fwd(vo[:, qid].contiguous().numpy(), scene.points)
        crossing = ((sv < 0).sum(1) != (svc < 0).sum(1)) | (nni != ioc).any(1)
        f = (verts.numpy().astype(np.float64) - scene.grid_min) / (scene.grid_max - scene.grid_min) * (D - 1)
        interior = (np.abs(f - np.round(f)) > 2e-3).all(-1)
        losses = op.trace("losses").cpu().sum(0)
        out["iterations"][str(k - 1)] = {
            "verts_rel_err": float((verts - vo).abs().max() / vo.abs().max()),
            "verts_abs_err_m": float((verts - vo).abs().max()),
            "nn_idx_mismatches": int((nni != io).sum()), "nn_dist_bit_mismatches": int((nnd.view(np.uint32) != do.view(np.uint32)).sum()),
            "nn_queries": int(nni.size),
            "sdf_abs_err": float(np.abs(sv - svo).max()), "sdf_grad_abs_err_away_from_voxel_faces": float(np.abs(sg - sgo)[interior].max()),
            "sdf_grad_abs_err_all_vertices": float(np.abs(sg - sgo).max()),
            "bodies_on_a_loss_discontinuity": int(crossing.sum()),
            "sdf_scale": float(np.abs(svo).max()), "sdf_grad_scale": float(np.abs(sgo).max()),
            "penetrating_vertices": int((svo < 0).sum()),
            "grad_x_err_over_max_per_body_worst": float(per_body[~crossing].max()), "grad_x_err_over_max_per_body_median": float(np.median(per_body)),
            "loss_terms_gpu": [float(x) for x in losses], "loss_terms_cpu": [float(x) for x in terms],
            "cpu_oracle_seconds": round(time.time() - t0, 2)}
    # the 6D-direct deviation on vertices, iteration 0
    op.fit(xh.cuda(), cam.cuda(), num_iter=1)
    direct = op.trace("verts").view(B, V, 3)
    rt = op.body_verts(xh.cuda(), cam.cuda())
    out["direct_6d_vs_axis_angle_round_trip_max_abs_m"] = float((direct - rt).abs().max())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

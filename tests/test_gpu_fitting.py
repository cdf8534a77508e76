"""GPU tests of the fitting loop (FittingOP) against the CPU restatement of
source/fitting_habitat.py's loop, plus the loop-level invariants the sharded job relies on."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)


def _world(small_model, B, contact="parts"):
    from psi_release_b200 import synthetic
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    xh = synthetic.make_body_params(scene, B, seed=3)
    cid = synthetic.make_contact_ids(431, contact)
    cfg = dict(model_data=small_model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
               contact_ids=cid, init_lr_h=0.1, num_iter=4, batch_size=B, device="cuda")
    return scene, xh, cid, cfg


def _oracle_kw(small_model, scene, cid):
    from psi_release_b200 import synthetic
    t = torch.tensor
    return dict(smplx_model=oracle.SMPLXOracle(small_model),
                vposer=oracle.VPoserDecoderOracle(synthetic.make_vposer_weights()), sdf=t(scene.sdf),
                gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points),
                contact_ids=cid, weights=W)


@pytest.mark.parametrize("mode", ["independent", "batch"])
def test_cal_loss_and_gradient_match_cpu_restatement(small_model, mode):
    from psi_release_b200.fitting import FittingOP
    from psi_release_b200.geometry import GeometryTransformer
    B = 3
    scene, xh, cid, cfg = _world(small_model, B)
    op = FittingOP(dict(cfg, loss_mode=mode, use_cuda_graph=False), W)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    xhr = GeometryTransformer.convert_to_6D_rot(torch.tensor(xh))
    pert = xhr + 0.01 * torch.randn(xhr.shape, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        op.xhr_rec.copy_(pert.cuda())
    terms = op.cal_loss(xhr.cuda(), cam.cuda())
    (g,) = torch.autograd.grad(sum(terms), op.xhr_rec)
    ref_rec = pert.clone().requires_grad_(True)
    rterms = oracle.cal_loss(xhr, ref_rec, cam.expand(B, -1, -1), loss_mode=mode, **_oracle_kw(small_model, scene, cid))
    (rg,) = torch.autograd.grad(sum(rterms), ref_rec)
    for a, b in zip(terms, rterms):
        assert abs(float(a) - float(b)) <= 1e-4 * max(1.0, abs(float(b)))
    assert float((g.cpu() - rg).abs().max()) <= 2e-4 * float(rg.abs().max())


def test_fit_matches_cpu_fit_loop_and_graph_equals_eager(small_model):
    from psi_release_b200.fitting import FittingOP
    B = 2
    scene, xh, cid, cfg = _world(small_model, B)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    eager = FittingOP(dict(cfg, use_cuda_graph=False), W).fit(torch.tensor(xh).cuda(), cam.cuda())
    op = FittingOP(dict(cfg, use_cuda_graph=True), W)
    graph = op.fit(torch.tensor(xh).cuda(), cam.cuda())
    assert torch.equal(eager, graph)                       # deterministic kernels, same order
    again = op.fit(torch.tensor(xh).cuda(), cam.cuda())    # graph reuse resets the optimiser state
    assert torch.equal(graph, again)
    ref = oracle.fit_loop(torch.tensor(xh), cam.expand(B, -1, -1), 4, 0.1, loss_mode="independent",
                          **_oracle_kw(small_model, scene, cid))
    # 4 Adam steps of size 0.1: sign-level agreement of every update -> a tight absolute band
    assert float((graph.cpu() - ref).abs().max()) < 5e-3
    host = op.fit_host(torch.tensor(xh).pin_memory(), cam.pin_memory())
    assert torch.equal(host, graph.cpu())


def test_independent_mode_is_shard_invariant(small_model):
    """Fitting bodies [0:4] together == fitting [0:2] and [2:4] separately: the property that
    lets the batch be sharded over GPUs with no collective in the loop."""
    from psi_release_b200.fitting import FittingOP
    scene, xh, cid, cfg = _world(small_model, 4)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    full = FittingOP(dict(cfg, batch_size=4), W).fit(torch.tensor(xh).cuda(), cam)
    lo = FittingOP(dict(cfg, batch_size=2), W).fit(torch.tensor(xh[:2]).cuda(), cam)
    hi = FittingOP(dict(cfg, batch_size=2), W).fit(torch.tensor(xh[2:]).cuda(), cam)
    # fused engine: every kernel is per-body deterministic -> bit-identical whatever the batch split
    assert torch.equal(full, torch.cat([lo, hi]))
    # autograd engine: the torch/cuBLAS VPoser MLP picks batch-dependent tilings -> tight tolerance
    a_full = FittingOP(dict(cfg, batch_size=4, engine="autograd"), W).fit(torch.tensor(xh).cuda(), cam)
    a_lo = FittingOP(dict(cfg, batch_size=2, engine="autograd"), W).fit(torch.tensor(xh[:2]).cuda(), cam)
    assert float((a_full[:2] - a_lo).abs().max()) < 1e-4


def test_fused_engine_matches_autograd_engine_and_cpu_loop(small_model):
    """psi_fit_run vs torch autograd over the psi ops vs the CPU restatement.
    The fused path hands rotation matrices to LBS instead of the reference's matrix -> axis-angle ->
    Rodrigues round trip (identity on SO(3) up to rounding), hence tolerances, not equality."""
    from psi_release_b200.fitting import FittingOP
    for contact in ("parts", "full"):
        scene, xh, cid, cfg = _world(small_model, 3, contact)
        cam = torch.tensor(scene.cam_ext).unsqueeze(0)
        fused = FittingOP(dict(cfg, engine="fused", num_iter=6), W)
        auto = FittingOP(dict(cfg, engine="autograd", num_iter=6), W)
        # one iteration: the Adam step is lr*sign(g) -> any sign disagreement shows up as 0.2
        f1 = fused.fit(torch.tensor(xh).cuda(), cam.cuda(), num_iter=1)
        a1 = auto.fit(torch.tensor(xh).cuda(), cam.cuda(), num_iter=1)
        # (components whose gradient is at Adam's eps scale, |g| ~ 1e-8, take a partial step that
        # amplifies rounding noise: they get a looser band)
        full_step = ((f1 - torch.tensor(xh).cuda()).abs() > 0.09) | ((a1 - torch.tensor(xh).cuda()).abs() > 0.09)
        d1 = (f1 - a1).abs()
        assert float(d1[full_step].max()) < 2e-3 and float(d1.max()) < 3e-2
        ff = fused.fit(torch.tensor(xh).cuda(), cam.cuda())
        aa = auto.fit(torch.tensor(xh).cuda(), cam.cuda())
        assert float((ff - aa).abs().max()) < 2e-2
        ref = oracle.fit_loop(torch.tensor(xh), cam.expand(3, -1, -1), 6, 0.1, loss_mode="independent",
                              **_oracle_kw(small_model, scene, cid))
        assert float((ff.cpu() - ref).abs().max()) < 2e-2
        # loss values of the last iteration agree with the autograd engine's
        np.testing.assert_allclose(fused.last_losses.cpu().numpy(), auto.last_losses.cpu().numpy(), rtol=2e-2, atol=1e-4)


def test_reference_style_single_body_pickle_flow(small_model, tmp_path):
    """fitting(pickle) -> save_result(pickle) as fitting_habitat.py:__main__ drives it."""
    from psi_release_b200 import io
    from psi_release_b200.fitting import FittingOP
    scene, xh, cid, cfg = _world(small_model, 1)
    fn = str(tmp_path / "gen" / "body_gen_000000.pkl")
    (tmp_path / "gen").mkdir()
    io.write_body_pickle(fn, xh[0], scene.cam_ext[None], np.eye(3, dtype=np.float32)[None])
    op = FittingOP(cfg, W)
    out = op.fitting(fn)
    assert out.shape == (1, 72) and torch.isfinite(out).all()
    op.save_result(out, str(tmp_path / "fit" / "body_gen_000000.pkl"))
    d = io.read_body_pickle(str(tmp_path / "fit" / "body_gen_000000.pkl"))
    assert set(d) == {"transl", "global_orient", "betas", "body_pose", "left_hand_pose", "right_hand_pose", "cam_ext", "cam_int"}
    assert d["body_pose"].shape == (1, 32)


def test_evaluation_scores_match_reference_definition(small_model):
    """utils_eval_collision_habitat.py:121-138 evaluated on the CPU restatement."""
    from psi_release_b200 import evaluate
    from psi_release_b200.fitting import FittingOP
    B = 6
    scene, xh, cid, cfg = _world(small_model, B)
    op = FittingOP(dict(cfg, engine="autograd"), W)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    nc, ct = evaluate.collision_scores(op, torch.tensor(xh).cuda(), cam.cuda())
    kw = _oracle_kw(small_model, scene, cid)
    cx = torch.tensor(xh)
    v, _ = kw["smplx_model"](body_pose=kw["vposer"].decode(cx[:, 16:48]), transl=cx[:, :3], global_orient=cx[:, 3:6],
                             betas=cx[:, 6:16], left_hand_pose=cx[:, 48:60], right_hand_pose=cx[:, 60:])
    v = oracle.verts_transform(v, cam.expand(B, -1, -1))
    s = oracle.sdf_lookup_torch(kw["sdf"], kw["gmin"], kw["gmax"], v)
    for b in range(B):
        if int((s[b] < 0).sum()) < 1:
            ref_nc, ref_ct = 1.0, 0.0
        else:
            ref_nc, ref_ct = float((s[b] > 0).sum()) / s.shape[1], 1.0
        assert abs(float(nc[b]) - ref_nc) <= 2.0 / s.shape[1]      # a vertex within rounding of sdf == 0
        assert float(ct[b]) == ref_ct
    assert 0 < float(ct.sum()) <= B


def test_multi_scene_fit_equals_per_scene_fits(small_model):
    """distributed.fit_scenes on one rank: two scenes, ragged body counts."""
    from psi_release_b200 import synthetic
    from psi_release_b200.distributed import fit_scenes
    from psi_release_b200.fitting import FittingOP
    scenes = [synthetic.make_scene(seed=s, dim=24, num_points=1500) for s in (4, 5)]
    xhs = [torch.tensor(synthetic.make_body_params(sc, n, seed=9 + i)) for i, (sc, n) in enumerate(zip(scenes, (3, 2)))]
    cams = [torch.tensor(sc.cam_ext).unsqueeze(0) for sc in scenes]
    cid = synthetic.make_contact_ids(431, "parts")

    def make(s, batch):
        return FittingOP(dict(model_data=small_model, scene=scenes[s], vposer_weights=synthetic.make_vposer_weights(),
                              contact_ids=cid, init_lr_h=0.1, num_iter=3, batch_size=batch, device="cuda"), W)

    out = fit_scenes(make, xhs, cams)
    assert out.shape == (5, 72)
    ref = torch.cat([make(s, xhs[s].shape[0]).fit(xhs[s].cuda(), cams[s].cuda()) for s in range(2)])
    assert torch.equal(out, ref)


def test_training_geometry_block_matches_train_s2_definition(small_model):
    """training.SceneLossBlock vs train_s2.py:136-202 restated on the CPU: two scenes in one batch,
    gradients reach an upstream (stock torch) layer, and the block is gated by the epoch."""
    from psi_release_b200 import body_model, synthetic, training
    from psi_release_b200.geometry import VPoserDecoder
    B = 5
    scenes = [synthetic.make_scene(seed=s, dim=24, num_points=1200) for s in (7, 8)]
    scene_ids = [0, 1, 1, 0, 1]
    xh = np.concatenate([synthetic.make_body_params(scenes[s], 1, seed=30 + i) for i, s in enumerate(scene_ids)])
    cams = torch.stack([torch.tensor(scenes[s].cam_ext) for s in scene_ids])
    cid = synthetic.make_contact_ids(431, "parts")
    vw = synthetic.make_vposer_weights()
    blk = training.SceneLossBlock(body_model.create(model_data=small_model, num_pca_comps=12, batch_size=B),
                                  VPoserDecoder.from_weights(vw), scenes, cid, 0.001, 0.01, 0.1)
    lin = torch.nn.Linear(72, 72).cuda()                       # stand-in for the CVAE decoder head
    with torch.no_grad():
        lin.weight.copy_(torch.eye(72)); lin.bias.zero_()
    x = torch.tensor(xh).cuda()
    lc, lv, lp = blk(lin(x), cams.cuda(), scene_ids, ep=29, epochs=30)
    (lc + lv + lp).backward()
    assert lin.weight.grad is not None and float(lin.weight.grad.abs().sum()) > 0
    # CPU restatement
    so = oracle.SMPLXOracle(small_model); vp = oracle.VPoserDecoderOracle(vw)
    cx = torch.tensor(xh)
    v, _ = so(body_pose=vp.decode(cx[:, 16:48]), transl=cx[:, :3], global_orient=cx[:, 3:6], betas=cx[:, 6:16],
              left_hand_pose=cx[:, 48:60], right_hand_pose=cx[:, 60:])
    v = oracle.verts_transform(v, cams)
    d = np.stack([oracle.nn_fwd(v[b:b + 1, cid].numpy(), scenes[s].points)[0][0] for b, s in enumerate(scene_ids)])
    sq = np.sqrt(d + 1e-4)
    ref_c = 0.01 * np.mean(sq / (sq + 1.0))
    sdf = torch.cat([oracle.sdf_lookup_torch(torch.tensor(scenes[s].sdf), torch.tensor(scenes[s].grid_min),
                                             torch.tensor(scenes[s].grid_max), v[b:b + 1]) for b, s in enumerate(scene_ids)])
    ref_p = 0.1 * float(sdf[sdf < 0].abs().mean()) if int((sdf < 0).sum()) else 0.0
    assert abs(float(lc) - ref_c) < 1e-5 and abs(float(lp) - ref_p) < 1e-5
    assert abs(float(lv) - 0.001 * float((cx[:, 16:48] ** 2).mean())) < 1e-7
    # before 75 % of the epochs the block is skipped (the reference multiplies it by 0.0)
    lc0, lv0, lp0 = blk(x, cams.cuda(), scene_ids, ep=3, epochs=30)
    assert float(lc0) == 0.0 and float(lp0) == 0.0 and float(lv0) > 0


def test_lbfgs_mode_decreases_the_loss_and_counts_closures(small_model):
    """optimizer_name='lbfgs' (the optional second mode): strong-Wolfe L-BFGS over the same
    cal_loss closure; it must lower the loss, count its closure evaluations and leave Adam untouched."""
    from psi_release_b200.fitting import FittingOP
    from psi_release_b200.geometry import GeometryTransformer
    scene, xh, cid, cfg = _world(small_model, 2)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    op = FittingOP(dict(cfg, optimizer_name="lbfgs", engine="autograd", num_iter=12), W)
    assert op.engine == "autograd"
    x0 = torch.tensor(xh).cuda()
    xhr0 = GeometryTransformer.convert_to_6D_rot(x0)
    with torch.no_grad():
        op.xhr_rec.copy_(xhr0)
    before = float(sum(op.cal_loss(xhr0, cam)))
    out = op.fit(x0, cam)
    after = float(op.last_losses.sum())
    assert out.shape == (2, 72) and torch.isfinite(out).all()
    assert 1 <= op.closure_evals <= 12 * 5 // 4 + 1
    assert after < before
    with pytest.raises(ValueError):
        FittingOP(dict(cfg, optimizer_name="lbfgs", engine="fused", loss_mode="batch"), W)   # per-body line searches need per-body losses
    assert FittingOP(dict(cfg, optimizer_name="lbfgs"), W).engine == "fused"


def test_rooms_pipeline_generate_fit_score_write(small_model, tmp_path):
    """pipeline.fit_rooms (BASELINE config 5 shape: rooms x samples): every room's samples are one batch,
    results equal per-room FittingOP.fit, scores equal evaluate.collision_scores, pickles are written
    in the reference's per-body layout."""
    from psi_release_b200 import evaluate, io, pipeline, synthetic
    from psi_release_b200.fitting import FittingOP
    rooms = [synthetic.make_scene(seed=s, dim=24, num_points=1500) for s in (11, 12, 13)]
    cid = synthetic.make_contact_ids(431, "parts")
    base = dict(model_data=small_model, vposer_weights=synthetic.make_vposer_weights(), contact_ids=cid,
                init_lr_h=0.1, num_iter=3, device="cuda")
    gen = lambda r, n: synthetic.make_body_params(rooms[r], n, seed=50 + r)
    res = pipeline.fit_rooms(rooms, gen, 4, base, W, out_dir=str(tmp_path))
    assert res["fitted"].shape == (12, 72) and res["non_collision"].shape == (12,) and res["contact"].shape == (12,)
    for r in range(3):
        op = FittingOP(dict(base, scene=rooms[r], batch_size=4), W)
        cam = torch.tensor(rooms[r].cam_ext).unsqueeze(0).cuda()
        ref = op.fit(torch.tensor(gen(r, 4)).cuda(), cam)
        assert torch.equal(res["fitted"][4 * r:4 * r + 4], ref)
        nc, ct = evaluate.collision_scores(op, ref, cam)
        assert torch.equal(res["non_collision"][4 * r:4 * r + 4], nc) and torch.equal(res["contact"][4 * r:4 * r + 4], ct)
        d = io.read_body_pickle(str(tmp_path / ("room_%03d" % r) / "body_gen_000002.pkl"))
        np.testing.assert_array_equal(d["transl"][0], ref[2, :3].cpu().numpy())


def test_train_step_config4_bf16_network_fp32_geometry(small_model):
    """training.TrainStep (BASELINE config 4 shape): a stand-in CVAE (stock torch MLPs + a conv scene
    encoder) under bf16 autocast, the geometry block in FP32 on the psi kernels.  The seven loss terms
    follow train_s2.py:102-204 (checked against a plain-torch restatement of the non-geometry terms),
    gradients reach the network, the epoch gate switches the geometry terms, a few steps lower the loss."""
    from psi_release_b200 import body_model, synthetic, training
    from psi_release_b200.geometry import GeometryTransformer, VPoserDecoder
    torch.manual_seed(0)
    B = 6
    scenes = [synthetic.make_scene(seed=s, dim=24, num_points=1200) for s in (21, 22)]
    scene_ids = [0, 1, 0, 1, 1, 0]
    xh = torch.tensor(np.concatenate([synthetic.make_body_params(scenes[s], 1, seed=60 + i) for i, s in enumerate(scene_ids)])).cuda()
    xh[:, 2] = xh[:, 2].abs() + 1.5                                  # in front of the camera (normalize_global_T divides by z)
    cams = torch.stack([torch.tensor(scenes[s].cam_ext) for s in scene_ids]).cuda()
    cam_int = torch.tensor([[500.0, 0, 320.0], [0, 500.0, 240.0], [0, 0, 1.0]]).repeat(B, 1, 1).cuda()
    max_d = torch.full((B,), 6.0).cuda()
    xs = torch.rand(B, 2, 32, 32).cuda() * 2 - 1                     # depth + semantics image (train_s2: [B,2,128,128])

    class TinyCVAE(torch.nn.Module):                                 # same call signature as HumanCVAES2.forward
        def __init__(self):
            super().__init__()
            self.scene = torch.nn.Sequential(torch.nn.Conv2d(2, 8, 3, 2, 1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())
            self.enc = torch.nn.Linear(75 + 8, 64)
            self.mu_g, self.ls_g, self.mu_l, self.ls_l = (torch.nn.Linear(64, 16) for _ in range(4))
            self.dec = torch.nn.Sequential(torch.nn.Linear(32 + 8, 128), torch.nn.ReLU(), torch.nn.Linear(128, 75))

        def forward(self, xhnr, eps_g, eps_l, xs):
            c = self.scene(xs)
            h = torch.relu(self.enc(torch.cat([xhnr, c], 1)))
            mg, lg, ml, ll = self.mu_g(h), self.ls_g(h), self.mu_l(h), self.ls_l(h)
            z = torch.cat([mg + eps_g * torch.exp(0.5 * lg), ml + eps_l * torch.exp(0.5 * ll), c], 1)
            return xhnr + 0.1 * self.dec(z), mg, lg, ml, ll          # residual: starts near the identity

    net = TinyCVAE().cuda()
    cid = synthetic.make_contact_ids(431, "parts")
    blk = training.SceneLossBlock(body_model.create(model_data=small_model, num_pca_comps=12, batch_size=B),
                                  VPoserDecoder.from_weights(synthetic.make_vposer_weights()), scenes, cid, 0.001, 0.01, 0.1)
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    ts = training.TrainStep(net, blk, opt, weight_loss_rec_h=1.0, weight_loss_kl=0.1, epochs=30)
    eps_g, eps_l = torch.randn(B, 16).cuda(), torch.randn(B, 16).cuda()
    batch = (xs, xh, eps_g, eps_l, cams, cam_int, max_d, scene_ids)
    early = ts.cal_loss(*batch, 3)
    assert float(early[4]) == 0.0 and float(early[6]) == 0.0        # gated off before 0.75 * epochs
    late = ts.cal_loss(*batch, 29)
    assert float(late[4]) > 0.0 and all(torch.isfinite(t) for t in late)
    # non-geometry terms vs a plain-torch restatement (network evaluated under the same autocast)
    xhnr = GeometryTransformer.convert_to_6D_rot(GeometryTransformer.normalize_global_T(xh, cam_int, max_d))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        rec, mg, lg, ml, ll = net(xhnr, eps_g, eps_l, xs)
    rec, mg, lg = rec.float(), mg.float(), lg.float()
    ref_p = torch.nn.functional.l1_loss(rec[:, 3:], xhnr[:, 3:])
    ref_kl = 0.1 * 0.5 * torch.mean(torch.exp(lg) + mg ** 2 - 1.0 - lg)             # fca = 1 at ep 29
    assert abs(float(late[1]) - float(ref_p)) < 1e-6 and abs(float(late[2]) - float(ref_kl)) < 1e-6
    sum(late).backward()
    g = [p.grad for p in net.parameters() if p.grad is not None]
    assert len(g) > 0 and all(torch.isfinite(x).all() for x in g) and any(float(x.abs().max()) > 0 for x in g)
    first = float(ts.step(*batch, ep=29).sum())
    for _ in range(15):
        last = float(ts.step(*batch, ep=29).sum())
    assert last < first

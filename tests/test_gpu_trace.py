"""The fused fitting loop (psi_fit_run: 13 launches per iteration, replayed by a CUDA graph) pinned NUMERICALLY
against the CPU oracle through psi_fit_trace: the loop's own vertices, SDF samples, NN results, loss values and
dL/dx (before Adam) of a chosen iteration -- not just the fitted vector after a few sign-like Adam steps.

Tolerances (BASELINE.json north_star): NN indices and distances bit-exact; LBS / SDF floats 1e-4 relative;
dL/dx 2e-4 of the largest component per body.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)


def _oracle_kw(model, scene, cid):
    from psi_release_b200 import synthetic
    t = torch.tensor
    return dict(smplx_model=oracle.SMPLXOracle(model),
                vposer=oracle.VPoserDecoderOracle(synthetic.make_vposer_weights()), sdf=t(scene.sdf),
                gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points),
                contact_ids=cid, weights=W)


def _oracle_verts(kw, x6, cam):
    """CPU restatement of VPoser decode -> SMPL-X -> verts_transform at the 75-D vector x6 (cal_loss's own
    chain, including the 6D -> axis-angle -> Rodrigues round trip the fused loop skips)."""
    xh = oracle.convert_to_3D_rot(x6)
    v, _ = kw["smplx_model"](body_pose=kw["vposer"].decode(xh[:, 16:48]), transl=xh[:, :3], global_orient=xh[:, 3:6],
                             betas=xh[:, 6:16], left_hand_pose=xh[:, 48:60], right_hand_pose=xh[:, 60:])
    return oracle.verts_transform(v, cam.expand(x6.shape[0], -1, -1))


def _check_iteration(op, kw, scene, cid, x0_6d, cam, loss_mode="independent", grad_tol=2e-4):
    """Everything psi_fit_trace exposes of the last evaluated iteration against the oracle."""
    B = x0_6d.shape[0]
    x_eval = op.trace("x_eval").cpu()
    verts = op.trace("verts").cpu().view(B, -1, 3)
    V = verts.shape[1]
    # --- LBS (+ decoder + rotation chain + camera transform): 1e-4 relative
    vo = _oracle_verts(kw, x_eval, cam)
    rel = float((verts - vo).abs().max() / vo.abs().max())
    assert rel <= 1e-4, f"verts rel err {rel}"
    # --- NN: bit-exact on the kernel's OWN vertices, in the loop's query order, hints carried over
    qid = op.trace("query_ids").cpu().long()
    nnd = op.trace("nn_dist").cpu().numpy()
    nni = op.trace("nn_idx").cpu().numpy()
    do, io = oracle.nn_fwd(verts[:, qid].contiguous().numpy(), scene.points)
    assert np.array_equal(nni, io), f"{int((nni != io).sum())} NN indices differ"
    assert np.array_equal(nnd.view(np.uint32), do.view(np.uint32))
    assert sorted(set(qid.tolist())) == sorted(set(int(c) for c in cid))           # the unique contact ids, all of them
    # --- SDF value + analytic gradient at the kernel's own vertices: 1e-4
    sv = op.trace("sdf").cpu().numpy()
    sg = op.trace("sdf_grad").cpu().numpy().reshape(B, V, 3)
    svo, sgo = oracle.sdf_fwd(scene.sdf, scene.grid_min, scene.grid_max, verts.numpy())
    assert np.abs(sv - svo).max() <= 1e-4 * max(1.0, np.abs(svo).max())
    # (the trilinear gradient jumps across voxel faces: a vertex within rounding of a face may land in the
    # neighbouring voxel on the other side of a contraction difference -- the value is continuous, the slope is not)
    f = (verts.numpy().astype(np.float64) - scene.grid_min) / (scene.grid_max - scene.grid_min) * (scene.sdf.shape[0] - 1)
    interior = (np.abs(f - np.round(f)) > 2e-3).all(-1)
    assert interior.mean() > 0.98
    assert np.abs(sg - sgo)[interior].max() <= 1e-4 * max(1.0, np.abs(sgo).max())
    # --- the four loss terms and dL/dx BEFORE Adam vs autograd of the CPU cal_loss at the same point
    xr = x_eval.clone().requires_grad_(True)
    terms = oracle.cal_loss(x0_6d, xr, cam.expand(B, -1, -1), loss_mode=loss_mode, **kw)
    (go,) = torch.autograd.grad(sum(terms), xr)
    losses = op.trace("losses").cpu().sum(0)
    for a, b in zip(losses.tolist(), terms):
        assert abs(a - float(b)) <= 1e-4 * max(1.0, abs(float(b))), (losses.tolist(), [float(t) for t in terms])
    g = op.trace("grad_x").cpu()
    # The reference loss has KINKS: the trilinear SDF slope jumps across voxel faces, a vertex crossing sdf = 0
    # changes the mean's denominator (fitting_habitat.py:155-158), a vertex between two scene points changes its
    # nearest one.  The CPU chain evaluates its own vertices (within ~2e-5 m of the kernel's), so on a kink the two
    # sides take different -- equally valid -- one-sided derivatives.  The slack per body is the L1 difference of
    # dL/dverts evaluated from the kernel's inputs and from the CPU chain's inputs (times a lever arm of 3 for the
    # rotations); away from kinks it is ~1e-7 and the plain tolerance decides.
    svc, sgc = oracle.sdf_fwd(scene.sdf, scene.grid_min, scene.grid_max, vo.numpy())
    doc, ioc = oracle.nn_fwd(vo[:, qid].contiguous().numpy(), scene.points)
    mult = np.bincount(np.asarray(cid, dtype=np.int64), minlength=V)[qid.numpy()].astype(np.float64)

    def vertex_grad(v, val, slope, nd, ni):
        gv = np.zeros((B, V, 3))
        neg = val < 0
        cnt = np.maximum(neg.sum() if loss_mode == "batch" else neg.sum(1, keepdims=True), 1)
        gv -= (W["weight_collision"] / cnt)[..., None] * slope * neg[..., None] if loss_mode != "batch" else \
            (W["weight_collision"] / cnt) * slope * neg[..., None]
        sq = np.sqrt(nd.astype(np.float64) + 1e-4)
        gd = W["weight_contact"] / (len(cid) * (B if loss_mode == "batch" else 1)) * mult * (1.0 / (sq + 1.0) ** 2) * (0.5 / sq)
        gv[:, qid.numpy()] += 2 * gd[..., None] * (v[:, qid.numpy()] - scene.points[ni])
        return gv
    slack = 3 * np.abs(vertex_grad(verts.numpy().astype(np.float64), sv, sg, nnd, nni) -
                       vertex_grad(vo.numpy().astype(np.float64), svc, sgc, doc, ioc)).sum((1, 2))
    on_kink = slack > 1e-5
    assert on_kink.sum() <= max(2, B // 4), f"{int(on_kink.sum())} of {B} bodies on a kink of the loss"
    for b in range(B):
        scale = float(go[b].abs().max())
        err = float((g[b] - go[b]).abs().max())
        assert err <= grad_tol * scale + slack[b], \
            f"body {b}: |dL/dx - oracle| = {err:.3e} vs scale {scale:.3e}, kink slack {slack[b]:.3e}"
    return x_eval, g


def _make(model, scene, cid, B, **over):
    from psi_release_b200 import synthetic
    from psi_release_b200.fitting import FittingOP
    cfg = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(), contact_ids=cid,
               init_lr_h=0.1, num_iter=3, batch_size=B, device="cuda", engine="fused")
    cfg.update(over)
    return FittingOP(cfg, W)


@pytest.mark.parametrize("contact", ["parts", "full"])
@pytest.mark.parametrize("B", [1, 3, 65])
def test_fused_iteration_gradient_matches_oracle_autograd(small_model, contact, B):
    """dL/dx of the fused kernels (lbs_vertex_bwd<FIT>, dcoef GEMM, pose backward, decoder backward, fit_step)
    at iteration 0 (x = x0: no L_rec gradient) and at iteration 2 (after two Adam steps), magnitude included."""
    from psi_release_b200 import synthetic
    from psi_release_b200.geometry import GeometryTransformer
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3))
    cid = synthetic.make_contact_ids(431, contact)
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    kw = _oracle_kw(small_model, scene, cid)
    op = _make(small_model, scene, cid, B)
    # the loop's own starting vector (6D conversion on the device): |x - x0| must have its kink at the same point
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()
    for k in (1, 3):
        op.fit(xh.cuda(), cam.cuda(), num_iter=k)
        x_eval, g = _check_iteration(op, kw, scene, cid, x0, cam)
        if k == 1:
            assert torch.equal(x_eval, x0)
        else:
            assert float((x_eval - x0).abs().max()) > 0.05         # two Adam steps of 0.1 moved it
        # the traced Adam state is what torch.optim.Adam holds after k steps of these gradients
        if k == 1:
            np.testing.assert_allclose(op.trace("adam_m").cpu().numpy(), 0.1 * g.numpy(), rtol=1e-6, atol=1e-12)
            np.testing.assert_allclose(op.trace("adam_v").cpu().numpy(), 0.001 * g.numpy() ** 2, rtol=1e-4, atol=1e-20)   # 1 - 0.999f


def test_fused_batch_coupled_loss_matches_reference_batch_semantics(small_model):
    """loss_mode='batch' on the fused engine = the reference's loss as written for B > 1
    (fitting_proxe.py:105,110,139,155-160; demo.ipynb cell 16): batch means, one penetration count."""
    from psi_release_b200 import synthetic
    from psi_release_b200.geometry import GeometryTransformer
    B = 5
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3))
    cid = synthetic.make_contact_ids(431, "parts")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    kw = _oracle_kw(small_model, scene, cid)
    op = _make(small_model, scene, cid, B, loss_mode="batch")
    assert op.engine == "fused"
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()
    for k in (1, 4):
        op.fit(xh.cuda(), cam.cuda(), num_iter=k)
        _check_iteration(op, kw, scene, cid, x0, cam, loss_mode="batch")
    # and the fitted vectors follow the CPU loop of the same semantics
    fitted = op.fit(xh.cuda(), cam.cuda(), num_iter=4)
    ref = oracle.fit_loop(xh, cam.expand(B, -1, -1), 4, 0.1, loss_mode="batch", **kw)
    assert float((fitted.cpu() - ref).abs().max()) < 5e-3
    auto = _make(small_model, scene, cid, B, loss_mode="batch", engine="autograd").fit(xh.cuda(), cam.cuda(), num_iter=4)
    assert float((fitted - auto).abs().max()) < 5e-3


def test_loop_forms_are_bit_identical(small_model):
    """One graph launch for the whole loop (conditional WHILE node) == one launch per iteration == eager
    launches; re-running a context resets its state; the iteration count is honoured exactly."""
    from psi_release_b200 import synthetic
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    B = 4
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
    cid = synthetic.make_contact_ids(431, "parts")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    whole = _make(small_model, scene, cid, B, loop_mode="whole")
    replay = _make(small_model, scene, cid, B, loop_mode="replay")
    eager = _make(small_model, scene, cid, B, use_cuda_graph=False)
    for k in (2, 7, 3, 31, 45):          # the WHILE body holds 15 iterations: remainders, whole passes, both
        a, b, c = whole.fit(xh, cam, num_iter=k), replay.fit(xh, cam, num_iter=k), eager.fit(xh, cam, num_iter=k)
        assert torch.equal(a, b) and torch.equal(a, c), k
        assert torch.equal(whole.trace("x_eval"), eager.trace("x_eval"))
        assert torch.equal(whole.trace("grad_x"), eager.trace("grad_x"))
    assert not torch.equal(whole.fit(xh, cam, num_iter=30), whole.fit(xh, cam, num_iter=29))
    import os
    os.environ["PSI_FIT_UNROLL"] = "1"
    try:
        one_per_pass = _make(small_model, scene, cid, B, loop_mode="whole")
    finally:
        del os.environ["PSI_FIT_UNROLL"]
    assert torch.equal(one_per_pass.fit(xh, cam, num_iter=31), replay.fit(xh, cam, num_iter=31))


def test_direct_6d_path_equals_axis_angle_round_trip_on_vertices(small_model):
    """the fused loop hands Gram-Schmidt rotations straight to LBS instead of the reference's
    matrix -> axis-angle (torchgeometry) -> Rodrigues round trip (vposer_smpl.py:153-161, cvae.py:128-137,
    lbs.py:165-192).  On SO(3) that is the identity up to the rounding of the float32 quaternion / atan2 /
    sincos chain: measured 1.7e-5 m on vertices of magnitude <= 3 m (1e-5 relative, inside the 1e-4 budget)."""
    from psi_release_b200 import synthetic
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    B = 16
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=11)).cuda()
    cid = synthetic.make_contact_ids(431, "parts")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    op = _make(small_model, scene, cid, B)
    op.fit(xh, cam, num_iter=1)                                   # iteration 0 is evaluated at x0
    direct = op.trace("verts").view(B, -1, 3)
    round_trip = op.body_verts(xh, cam)                           # psi LBS kernels fed with axis-angle vectors
    scale = float(round_trip.abs().max())
    assert float((direct - round_trip).abs().max()) <= 3e-5 * scale
    cpu = _oracle_verts(_oracle_kw(small_model, scene, cid), oracle.convert_to_6D_rot(xh.cpu()), cam.cpu())
    assert float((direct.cpu() - cpu).abs().max()) <= 3e-5 * scale


def test_fused_loop_at_baseline_size(full_model):
    """BASELINE configs[1] itself: 64 bodies, 10 475 vertices, every vertex a contact vertex, 256^3 SDF,
    50 000 scene points; iteration 2 of a graph-driven run (hints carried over two iterations, the loop's own
    joint/kd query order).  NN bit-exact, LBS/SDF 1e-4, losses 1e-4, dL/dx 2e-4."""
    from psi_release_b200 import synthetic
    from psi_release_b200.geometry import GeometryTransformer
    B, V = 64, 10475
    scene = synthetic.make_scene(seed=0, dim=256, num_points=50000)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=0))
    cid = synthetic.make_contact_ids(V, "full")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    kw = _oracle_kw(full_model, scene, cid)
    op = _make(full_model, scene, cid, B)
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()
    op.fit(xh.cuda(), cam.cuda(), num_iter=3)
    x_eval, _ = _check_iteration(op, kw, scene, cid, x0, cam)
    assert float((x_eval - x0).abs().max()) > 0.05
    # penetrating and floating bodies are both present, so both loss branches were exercised
    sv = op.trace("sdf").cpu()
    assert int((sv < 0).any(dim=1).sum()) > 0 and int((~(sv < 0).any(dim=1)).sum()) >= 0
    # 300 iterations stay finite and the run is reproducible bit for bit
    a = op.fit(xh.cuda(), cam.cuda(), num_iter=300)
    b = op.fit(xh.cuda(), cam.cuda(), num_iter=300)
    assert torch.isfinite(a).all() and torch.equal(a, b)


def test_batch_coupled_loss_sharded_over_two_contexts_equals_the_union_batch(small_model):
    """SURVEY.md 8(e) option 2: the batch-coupled loss with the batch cut into shards.  Two contexts of this process
    (3 + 2 bodies, each on its own stream) exchange their penetration counts every iteration through each other's
    exchange buffers inside the loop -- and must reproduce, bit for bit, one context fitting all 5 bodies.  (Across
    GPUs the same buffers are opened through CUDA IPC: tools/probes/sharded_batch_check.py under torchrun.)"""
    from psi_release_b200 import synthetic
    from psi_release_b200.geometry import GeometryTransformer
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    B = 5
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
    cid = synthetic.make_contact_ids(431, "parts")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    union = _make(small_model, scene, cid, B, loss_mode="batch")
    ref = union.fit(xh, cam, num_iter=6)
    a = _make(small_model, scene, cid, 3, loss_mode="batch")
    b = _make(small_model, scene, cid, 2, loss_mode="batch")
    fa, fb = a._fused, b._fused
    fa.connect_local([fa, fb], 0, B)
    fb.connect_local([fa, fb], 1, B)
    x6 = GeometryTransformer.convert_to_6D_rot(xh)
    for _ in range(2):                                   # a second fit: the sequence number keeps stale slots from matching
        fa.begin(x6[:3], cam, 6)
        fb.begin(x6[3:], cam, 6)
        (oa, la), (ob, lb) = fa.end(), fb.end()
        torch.cuda.synchronize()
        got = GeometryTransformer.convert_to_3D_rot(torch.cat([oa, ob]))
        assert torch.equal(got, ref)
        assert int(fa.trace("exchange")[0, 1]) == 0 and int(fb.trace("exchange")[0, 1]) == 0       # every count arrived in time
        assert torch.equal(torch.cat([la, lb]).sum(0), union.last_losses)
    # and it differs from two unconnected batch-mode shards (the coupling is real)
    lone = _make(small_model, scene, cid, 3, loss_mode="batch").fit(xh[:3], cam, num_iter=6)
    assert not torch.equal(lone, ref[:3])


@pytest.mark.gpu
def test_fit_is_bit_identical_under_every_launch_mode_in_subprocesses():
    """Programmatic dependent launch (PSI_PDL, read once per process: csrc/api.cu) lets a kernel's constant-only
    prologue run under its predecessor's tail; every kernel waits (griddepcontrol.wait) before it touches anything
    the predecessor writes.  The evidence that no wait is missing or misplaced: the same fit, Adam and L-BFGS,
    per-body and batch-coupled loss, is bit-identical with plain graph edges (0), with every kernel programmatic
    (1), with the NN walk + vertex backward only (2) and with the default (3, small grids only)."""
    import hashlib, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import hashlib, sys, torch
sys.path.insert(0, %r)
from psi_release_b200 import synthetic
from psi_release_b200.fitting import FittingOP
W = dict(weight_loss_rec=1.0, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
model = synthetic.make_smplx_model(seed=1234, num_verts=431)
scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
h = hashlib.sha256()
for B, kw in ((5, dict()), (3, dict(loss_mode="batch")), (4, dict(optimizer_name="lbfgs")), (65, dict())):
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
    cfg = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
               contact_ids=synthetic.make_contact_ids(431, "parts"), init_lr_h=0.1, num_iter=33, batch_size=B,
               device="cuda", engine="fused")
    cfg.update(kw)
    op = FittingOP(cfg, W)
    out = op.fit(xh, cam)
    assert torch.isfinite(out).all()
    h.update(out.cpu().numpy().tobytes())
    h.update(op.trace("grad_x").cpu().numpy().tobytes())
print("HASH", h.hexdigest())
""" % (root,)
    digests = {}
    for mode in ("0", "1", "2", "3"):
        env = dict(os.environ, PSI_PDL=mode)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (mode, r.stderr[-2000:])
        digests[mode] = [l for l in r.stdout.splitlines() if l.startswith("HASH")][-1]
    assert len(set(digests.values())) == 1, digests


def test_launch_count_claim_matches_the_kernels_actually_launched(small_model):
    """bench.py reports gpu_launches = psi_fit_launches_per_iteration() x iterations x steps.  The claim is checked
    against the library's own launch counter on an eager run (one count per kernel launch; graph replays do not pass
    through the counter): k iterations cost k x per_iteration launches + the opening fit_step + the reset kernels."""
    from psi_release_b200 import _lib, synthetic
    L = _lib.lib()
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    B = 64                                   # a full body group: no padding-row kernels
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    op = _make(small_model, scene, synthetic.make_contact_ids(431, "parts"), B, use_cuda_graph=False)
    per = int(L.psi_fit_launches_per_iteration())
    assert per == 13
    counts = []
    for k in (1, 4):
        c0 = L.psi_launch_count()
        op.fit(xh, cam, num_iter=k)
        torch.cuda.synchronize()
        counts.append(L.psi_launch_count() - c0)
    assert counts[1] - counts[0] == 3 * per, counts      # three more iterations = 3 x 13 launches

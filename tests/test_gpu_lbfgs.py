"""The per-body L-BFGS / strong-Wolfe mode of the fused loop (csrc/fit_lbfgs.cuh; lbfgs_ls.py:54-183,275-463) against
its CPU statement oracle/lbfgs.py (itself pinned to torch.optim.LBFGS, tests/test_oracle.py) driven by the CPU
cal_loss, evaluation by evaluation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import lbfgs as olbfgs  # noqa: E402
from oracle import oracle  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)


def _world(small_model, B):
    from psi_release_b200 import synthetic
    scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3))
    cid = synthetic.make_contact_ids(431, "parts")
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    cfg = dict(model_data=small_model, scene=scene, vposer_weights=synthetic.make_vposer_weights(), contact_ids=cid,
               init_lr_h=0.1, num_iter=3, batch_size=B, device="cuda", optimizer_name="lbfgs")
    t = torch.tensor
    kw = dict(smplx_model=oracle.SMPLXOracle(small_model), vposer=oracle.VPoserDecoderOracle(synthetic.make_vposer_weights()),
              sdf=t(scene.sdf), gmin=t(scene.grid_min), gmax=t(scene.grid_max), scene_points=t(scene.points),
              contact_ids=cid, weights=W)
    return scene, xh, cid, cam, cfg, kw


def test_fused_lbfgs_follows_the_cpu_machine_evaluation_by_evaluation(small_model):
    from psi_release_b200.fitting import FittingOP
    from psi_release_b200.geometry import GeometryTransformer
    B = 3
    scene, xh, cid, cam, cfg, kw = _world(small_model, B)
    op = FittingOP(cfg, W)
    assert op.engine == "fused"
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()

    def closure(b):
        def f(x):
            xr = torch.tensor(x[None], dtype=torch.float32, requires_grad=True)
            terms = oracle.cal_loss(x0[b:b + 1], xr, cam, loss_mode="independent", **kw)
            loss = sum(terms)
            (g,) = torch.autograd.grad(loss, xr)
            return float(loss), g[0].numpy()
        return f
    machines = [olbfgs.LBFGSMachine(lr=1.0, history_size=100, tolerance_grad=1e-5, tolerance_change=1e-9, max_iter=1 << 30,
                                    zoom_max_iter=300) for _ in range(B)]
    xs = [m.start(x0[b].numpy().copy()) for b, m in enumerate(machines)]
    for k in range(1, 9):
        for b, m in enumerate(machines):
            xs[b] = m.feed(*closure(b)(xs[b]))
        op.fit(xh.cuda(), cam.cuda(), num_iter=k)
        trial = op.trace("x").cpu().numpy()                    # the point the NEXT evaluation would use
        best = op.trace("lbfgs_best").cpu().numpy()
        state = op.trace("lbfgs_state").cpu().numpy()
        tol = 2e-4 * 4 ** (k // 3)                             # float32 rounding is amplified by every line-search decision
        for b, m in enumerate(machines):
            assert state[b, 2] == k and state[b, 0] == m.phase and state[b, 1] == m.n_iter, (k, b, state[b], m.phase, m.n_iter)
            assert np.abs(trial[b] - xs[b]).max() <= tol * max(1.0, np.abs(xs[b]).max()), (k, b)
            assert np.abs(best[b] - m.best()).max() <= tol * max(1.0, np.abs(xs[b]).max()), (k, b)


def test_fused_lbfgs_lowers_every_bodys_loss_is_shard_invariant_and_graph_equals_eager(small_model):
    from psi_release_b200.fitting import FittingOP
    from psi_release_b200.geometry import GeometryTransformer
    B = 6
    scene, xh, cid, cam, cfg, kw = _world(small_model, B)
    op = FittingOP(cfg, W)
    x0 = GeometryTransformer.convert_to_6D_rot(xh.cuda()).cpu()

    def per_body_loss(x6):
        return torch.stack([sum(oracle.cal_loss(x0[b:b + 1], x6[b:b + 1], cam, loss_mode="independent", **kw)) for b in range(B)])
    before = per_body_loss(x0)
    fitted = op.fit(xh.cuda(), cam.cuda(), num_iter=45)
    after = per_body_loss(GeometryTransformer.convert_to_6D_rot(fitted).cpu())
    # never worse; a body whose start is already a minimiser of the kinked loss (|x - x0| has slope 1/75 per
    # component at x0: if the other terms pull less than that, no step decreases the loss) stays where it is
    assert bool((after <= before + 1e-6).all()), (before.tolist(), after.tolist())
    moved = after < before - 1e-5
    assert int(moved.sum()) >= 2
    st = op.trace("lbfgs_state").cpu().numpy()
    assert (st[:, 2] == 45).all()
    assert (st[moved.numpy(), 1] >= 2).all() and (st[moved.numpy(), 3] >= 1).all()    # accepted steps, curvature pairs stored
    # every body owns its optimiser: any batch split gives the same bits; so do the three loop forms
    lo = FittingOP(dict(cfg, batch_size=2), W).fit(xh[:2].cuda(), cam.cuda(), num_iter=45)
    hi = FittingOP(dict(cfg, batch_size=4), W).fit(xh[2:].cuda(), cam.cuda(), num_iter=45)
    assert torch.equal(fitted, torch.cat([lo, hi]))
    for over in (dict(loop_mode="replay"), dict(use_cuda_graph=False)):
        assert torch.equal(fitted, FittingOP(dict(cfg, **over), W).fit(xh.cuda(), cam.cuda(), num_iter=45))
    # Adam contexts are untouched by the option
    adam = FittingOP(dict(cfg, optimizer_name="adam"), W).fit(xh.cuda(), cam.cuda(), num_iter=5)
    assert not torch.equal(adam, fitted) and torch.isfinite(adam).all()

"""CPU tests: the oracle against the reference-derived golden vectors and torch."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle


def _lbs_inputs(model, g):
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    sd = model["shapedirs"]
    pd = model["posedirs"]
    parents = model["kintree_table"][0].astype(np.int64).copy()
    parents[0] = -1
    return dict(v_template=t(model["v_template"]), shapedirs=t(sd[:, :, :20]),
                posedirs=t(pd.reshape(pd.shape[0] * 3, -1).T), J_regressor=t(model["J_regressor"]),
                parents=parents, lbs_weights=t(model["weights"]))


@pytest.mark.parametrize("name,fixture", [("lbs_small.npz", "small_model"), ("lbs_full.npz", "full_model")])
def test_lbs_oracle_matches_reference_golden(name, fixture, golden_dir, request):
    model = request.getfixturevalue(fixture)
    g = np.load(os.path.join(golden_dir, name))
    kw = _lbs_inputs(model, g)
    betas = torch.tensor(g["betas"], requires_grad=True)
    pose = torch.tensor(g["pose"], requires_grad=True)
    verts, joints = oracle.lbs(betas, pose, **kw)
    np.testing.assert_allclose(verts.detach().numpy(), g["verts"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(joints.detach().numpy(), g["joints"], rtol=0, atol=2e-6)
    probe = np.random.default_rng(int(g["probe_seed"]))
    # same draw order as make_golden.lbs_case: betas, pose, probe
    B, V = g["verts"].shape[:2]
    probe.standard_normal((B, 20)); probe.standard_normal((B, 165))
    w = torch.tensor(probe.standard_normal((B, V, 3)).astype(np.float32))
    (verts * w).sum().backward()
    np.testing.assert_allclose(betas.grad.numpy(), g["grad_betas"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(pose.grad.numpy(), g["grad_pose"], rtol=2e-4, atol=2e-4)


def test_nn_vectorised_equals_scalar_and_bruteforce():
    rng = np.random.default_rng(0)
    q = rng.uniform(-1, 1, (2, 257, 3)).astype(np.float32)
    s = rng.uniform(-1, 1, (2, 1003, 3)).astype(np.float32)
    s[:, 500] = s[:, 17]           # exact duplicates: lowest index must win
    s[:, 900] = s[:, 17]
    q[0, 5] = s[0, 17]
    d_v, i_v = oracle.nn_fwd(q, s)
    d_s, i_s = oracle.nn_fwd(q, s, scalar=True)
    assert np.array_equal(i_v, i_s) and np.array_equal(d_v.view(np.uint32), d_s.view(np.uint32))
    assert i_v[0, 5] == 17 and d_v[0, 5] == 0.0
    # float64 brute force agrees on the minimum value to rounding
    dd = ((q[:, :, None, :].astype(np.float64) - s[:, None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(d_v, dd.min(-1), rtol=1e-5, atol=1e-7)
    # shared scene (stride 0) == replicated scene
    d0, i0 = oracle.nn_fwd(q, s[0])
    d1, i1 = oracle.nn_fwd(q, np.stack([s[0], s[0]]))
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_chamfer_matches_reference_test_formula(golden_dir):
    """chamfer_pytorch/test_chamfer.py:35-54: sum sq err vs the matmul formula < 1e-8."""
    g = np.load(os.path.join(golden_dir, "chamfer_matmul_4x100.npz"))
    d1, d2, i1, i2 = oracle.chamfer_fwd(g["a"], g["b"])
    err = ((d1 - g["dist1"]) ** 2).sum() + ((d2 - g["dist2"]) ** 2).sum()
    assert err < 1e-8
    assert i1.min() >= 0 and i1.max() < 100


def test_chamfer_bwd_matches_autograd_of_gather():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((2, 50, 3)).astype(np.float32)
    b = rng.standard_normal((2, 70, 3)).astype(np.float32)
    d1, d2, i1, i2 = oracle.chamfer_fwd(a, b)
    g1 = rng.standard_normal(d1.shape).astype(np.float32)
    g2 = rng.standard_normal(d2.shape).astype(np.float32)
    ga, gb = oracle.chamfer_bwd(a, b, g1, g2, i1, i2)
    ta = torch.tensor(a, requires_grad=True)
    tb = torch.tensor(b, requires_grad=True)
    n1 = torch.gather(tb, 1, torch.tensor(i1).long().unsqueeze(-1).expand(-1, -1, 3))
    n2 = torch.gather(ta, 1, torch.tensor(i2).long().unsqueeze(-1).expand(-1, -1, 3))
    loss = (((ta - n1) ** 2).sum(-1) * torch.tensor(g1)).sum() + (((tb - n2) ** 2).sum(-1) * torch.tensor(g2)).sum()
    loss.backward()
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-5, atol=1e-5)


def test_sdf_c_oracle_matches_torch_grid_sample_align_corners_true():
    rng = np.random.default_rng(2)
    D = 16
    sdf = rng.standard_normal((D, D, D)).astype(np.float32)
    gmin = np.array([-1.0, -2.0, -0.5], np.float32)
    gmax = np.array([1.5, 2.0, 3.0], np.float32)
    v = rng.uniform(-2.5, 3.5, (3, 400, 3)).astype(np.float32)      # many outside: border clamp
    v[0, 0] = gmin; v[0, 1] = gmax                                   # exact corners
    val, grad = oracle.sdf_fwd(sdf, gmin, gmax, v)
    tv = torch.tensor(v, requires_grad=True)
    out = oracle.sdf_lookup_torch(torch.tensor(sdf), torch.tensor(gmin), torch.tensor(gmax), tv)
    np.testing.assert_allclose(val, out.detach().numpy(), rtol=1e-5, atol=1e-5)
    out.sum().backward()
    np.testing.assert_allclose(grad, tv.grad.numpy(), rtol=1e-4, atol=1e-4)
    # torch >= 1.3 default (align_corners=False) is NOT the spec
    import torch.nn.functional as F
    norm = (tv.detach() - torch.tensor(gmin)) / (torch.tensor(gmax) - torch.tensor(gmin)) * 2 - 1
    wrong = F.grid_sample(torch.tensor(sdf).view(1, 1, D, D, D).expand(3, -1, -1, -1, -1),
                          norm[:, :, [2, 1, 0]].view(3, -1, 1, 1, 3), padding_mode="border",
                          align_corners=False).view(3, -1)
    assert (wrong - out.detach()).abs().max() > 1e-2


def test_rotation_chain_roundtrip():
    rng = np.random.default_rng(3)
    aa = torch.tensor((rng.standard_normal((64, 3)) * 0.8).astype(np.float32))
    R = oracle.aa_to_matrix(aa)
    np.testing.assert_allclose(oracle.matrix_to_aa(R).numpy(), aa.numpy(), atol=2e-5)
    np.testing.assert_allclose(oracle.batch_rodrigues(aa).numpy(), R.numpy(), atol=2e-6)
    x = torch.cat([torch.zeros(64, 3), aa, torch.zeros(64, 66)], dim=1)
    back = oracle.convert_to_3D_rot(oracle.convert_to_6D_rot(x))
    np.testing.assert_allclose(back.numpy(), x.numpy(), atol=2e-5)


def test_lbfgs_machine_reproduces_torch_optim_lbfgs_step_for_step():
    """oracle/lbfgs.py restates human_body_prior/optimizers/lbfgs_ls.py (a copy of pytorch PR #8824) as a state
    machine fed one closure evaluation at a time.  torch.optim.LBFGS(line_search_fn='strong_wolfe') is the merged
    form of that PR: with reset_lr=True and the zoom bound at torch's max_ls the two must evaluate the SAME points,
    on the Rosenbrock function the reference file defines (lbfgs_ls.py:15-18)."""
    from oracle import lbfgs

    def rosen(x):
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        f = (100 * (xt[1:] - xt[:-1] ** 2) ** 2 + (1 - xt[:-1]) ** 2).sum()
        f.backward()
        return float(f), xt.grad.numpy().copy()
    x0 = np.array([-1.2, 1.0, 0.3, -0.7, 1.5])
    for budget in (1, 3, 8, 21, 60):
        p = torch.tensor(x0.copy(), requires_grad=True)
        opt = torch.optim.LBFGS([p], lr=1.0, max_iter=25, max_eval=budget, tolerance_grad=0, tolerance_change=1e-300,
                                history_size=100, line_search_fn="strong_wolfe")
        pts = []

        def closure():
            opt.zero_grad()
            f = (100 * (p[1:] - p[:-1] ** 2) ** 2 + (1 - p[:-1]) ** 2).sum()
            f.backward()
            pts.append(p.detach().numpy().copy())
            return f
        opt.step(closure)
        m = lbfgs.LBFGSMachine(lr=1.0, history_size=100, tolerance_grad=0, tolerance_change=1e-300, max_iter=25,
                               zoom_max_iter=25, reset_lr=True)
        x = m.start(x0)
        for ref_pt in pts:
            np.testing.assert_allclose(x, ref_pt, rtol=0, atol=1e-9)       # the same point is evaluated ...
            x = m.feed(*rosen(x))
        np.testing.assert_allclose(m.best() if m.phase != lbfgs.DONE else m.x_init, p.detach().numpy(), atol=1e-9)   # ... and accepted
    # a history shorter than the run exercises the ring (oldest pair dropped, lbfgs_ls.py:344-347)
    xb, m = lbfgs.minimize(rosen, x0, 120, lr=1.0, history_size=3, tolerance_grad=0, tolerance_change=1e-300)
    assert rosen(xb)[0] < 1e-6 and len(m.Y) == 3

"""GPU edge cases: empty / ragged / boundary sizes, error paths of the C ABI, odd batch sizes."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402


def _cuda(a):
    return torch.tensor(np.asarray(a), device="cuda")


def test_empty_and_tiny_clouds():
    from psi_release_b200 import chamfer
    a = torch.zeros(2, 0, 3, device="cuda")
    b = torch.rand(2, 5, 3, device="cuda")
    d1, d2, i1, i2 = chamfer.chamfer_forward(a, b)
    assert d1.shape == (2, 0) and d2.shape == (2, 5)            # n == 0: nothing to write, no launch error
    a = torch.rand(1, 3, 3, device="cuda")
    d1, d2, i1, i2 = chamfer.chamfer_forward(a, b[:1])
    od1, od2, oi1, oi2 = oracle.chamfer_fwd(a.cpu().numpy(), b[:1].cpu().numpy())
    assert np.array_equal(i1.cpu().numpy(), oi1) and np.array_equal(i2.cpu().numpy(), oi2)
    # a scene smaller than one cluster / one lane group, through the index
    s = torch.rand(7, 3, device="cuda")
    q = torch.rand(2, 9, 3, device="cuda")
    d, i = chamfer.nn_forward(q, chamfer.SceneIndex(s))
    do, io = oracle.nn_fwd(q.cpu().numpy(), s.cpu().numpy())
    assert np.array_equal(i.cpu().numpy(), io) and np.array_equal(d.cpu().numpy().view(np.uint32), do.view(np.uint32))


def test_cabi_error_codes():
    from psi_release_b200 import _lib
    L = _lib.lib()
    z = ctypes.c_void_p(0)
    st = _lib.stream_ptr()
    assert L.psi_nn_fwd(z, 3, 1, 4, z, 0, 5, z, z, z, 0, st) == -1                 # null pointers
    assert L.psi_nn_fwd(z, 3, -1, 4, z, 0, 5, z, z, z, 0, st) == -1                # negative size
    assert L.psi_nn_fwd(z, 3, 0, 4, z, 0, 5, z, z, z, 0, st) == 0                  # empty batch is fine
    q = torch.rand(64, 10475, 3, device="cuda"); s = torch.rand(50000, 3, device="cuda")
    d = torch.empty(64, 10475, device="cuda")
    assert L.psi_nn_fwd(_lib.ptr(q), 10475 * 3, 64, 10475, _lib.ptr(s), 0, 50000, _lib.ptr(d), z, z, 0, st) == -2  # workspace
    assert L.psi_sdf_fwd(z, 1, 0, z, z, z, 1, 1, z, z, z, z, st) == -1
    assert L.psi_sdf_fwd(_lib.ptr(s), 99, 8, _lib.ptr(s), _lib.ptr(s), _lib.ptr(s), 1, 1, z, _lib.ptr(d), z, z, st) == -3  # > 16 scenes
    h = ctypes.c_void_p()
    nanpts = np.full((4, 3), np.nan, np.float32)
    assert L.psi_nn_index_create(ctypes.byref(h), nanpts.ctypes.data_as(ctypes.c_void_p), 4, st) == -1   # finite only
    assert b"bad argument" in L.psi_error_string(-1) and b"not supported" in L.psi_error_string(-3)
    with pytest.raises(_lib.PsiError):
        _lib.check(-2, "demo")
    torch.cuda.synchronize()


@pytest.mark.parametrize("B", [1, 3, 65, 130])
def test_fused_fit_odd_batches_and_kbg_boundary(small_model, B):
    """B = 1 (the reference's own batch size), a ragged body group, one body past the 64-body
    group boundary of the vertex kernel, and three body groups (the blend GEMMs interleave the groups of a basis
    tile over neighbouring CTAs): fused == autograd within tolerance, finite, reproducible."""
    from psi_release_b200 import synthetic
    from psi_release_b200.fitting import FittingOP
    scene = synthetic.make_scene(seed=2, dim=24, num_points=1500)
    xh = torch.tensor(synthetic.make_body_params(scene, B, seed=5)).cuda()
    cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
    cfg = dict(model_data=small_model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
               contact_ids=synthetic.make_contact_ids(431, "parts"), init_lr_h=0.1, num_iter=3, batch_size=B, device="cuda")
    W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
    f = FittingOP(dict(cfg, engine="fused"), W)
    out = f.fit(xh, cam)
    assert out.shape == (B, 72) and torch.isfinite(out).all()
    assert torch.equal(out, f.fit(xh, cam))
    if B <= 3:
        a = FittingOP(dict(cfg, engine="autograd"), W).fit(xh, cam)
        assert float((out - a).abs().max()) < 1e-2
    # a body's result does not depend on who shares its batch
    if B > 1:
        one = FittingOP(dict(cfg, engine="fused", batch_size=1), W).fit(xh[B - 1:], cam)
        assert torch.equal(one, out[B - 1:])
    if B > 64:      # ... nor on which 64-body group of the GEMM operands it lands in
        first = FittingOP(dict(cfg, engine="fused", batch_size=64), W).fit(xh[:64], cam)
        assert torch.equal(first, out[:64])
        mid = FittingOP(dict(cfg, engine="fused", batch_size=B - 64), W).fit(xh[64:], cam)
        assert torch.equal(mid, out[64:])


def test_lbs_dense_skinning_weights_and_small_tree():
    """KW > 8 (dense weights: the kernels' generic path) and a 5-joint model."""
    from psi_release_b200 import body_model, synthetic
    rng = np.random.default_rng(0)
    parents = np.array([-1, 0, 1, 1, 0])
    model = synthetic.make_smplx_model(seed=3, num_verts=200, parents=parents)
    w = rng.random((200, 5)).astype(np.float32) + 0.01
    model["weights"] = (w / w.sum(1, keepdims=True)).astype(np.float32)           # 5 non-zeros per vertex
    B = 4
    betas = rng.standard_normal((B, 20)).astype(np.float32)
    pose = (rng.standard_normal((B, 15)) * 0.4).astype(np.float32)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    pd = model["posedirs"]
    tb, tp = t(betas).requires_grad_(True), t(pose).requires_grad_(True)
    vo, jo = oracle.lbs(tb, tp, t(model["v_template"]), t(model["shapedirs"][:, :, :20]),
                        t(pd.reshape(pd.shape[0] * 3, -1).T), t(model["J_regressor"]), parents, t(model["weights"]))
    h = body_model._ModelHandle(model["v_template"], model["shapedirs"][:, :, :20], pd.reshape(pd.shape[0] * 3, -1).T,
                                model["J_regressor"], model["weights"], parents, "cuda")
    cb, cp = _cuda(betas).requires_grad_(True), _cuda(pose).requires_grad_(True)
    v, j = body_model.lbs(cb, cp, h, want_joints=True)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(v.detach().cpu().numpy(), vo.detach().numpy()) < 1e-4
    assert rel(j.detach().cpu().numpy(), jo.detach().numpy()) < 1e-4
    wv = rng.standard_normal(vo.shape).astype(np.float32)
    (vo * t(wv)).sum().backward()
    (v * _cuda(wv)).sum().backward()
    assert rel(cb.grad.cpu().numpy(), tb.grad.numpy()) < 3e-4
    assert rel(cp.grad.cpu().numpy(), tp.grad.numpy()) < 3e-4


def test_nan_and_overflowing_queries_return_point_zero_like_the_reference():
    """A diverged body (NaN or overflowing vertices) must not produce an out-of-range index: every NN path
    returns index 0 and the distance to point 0, which is what the reference's `k == 0 ||` clauses leave
    (chamfer.cu:36,121,126) and what the oracle restates; the other queries of the call are untouched."""
    from psi_release_b200 import _lib, chamfer
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    m, n, B = 5000, 130, 2
    s = torch.rand(m, 3, generator=g).cuda()
    q = torch.rand(B, n, 3, generator=g)
    q[0, 7] = float("nan")
    q[1, 64, 1] = float("nan")
    q[1, 100] = 3e30                      # squares overflow to +inf: no distance compares below +inf
    q = q.cuda()
    do, io = oracle.nn_fwd(q.cpu().numpy(), s.cpu().numpy())
    assert io[0, 7] == 0 and io[1, 64] == 0 and io[1, 100] == 0 and np.isnan(do[0, 7]) and np.isinf(do[1, 100])
    ix = chamfer.SceneIndex(s)
    outs = {"brute": chamfer.nn_forward(q, s)}
    for mode in (1, 2, 3):
        d = torch.empty(B, n, device="cuda")
        i = torch.empty(B, n, dtype=torch.int32, device="cuda")
        h = torch.full((B, n), -1, dtype=torch.int32, device="cuda")
        for _ in range(2):                # second call: with the hints of the first
            rc = L.psi_nn_index_query_mode(ix.h, _lib.ptr(q), n * 3, B, n, None, _lib.ptr(d), _lib.ptr(i), _lib.ptr(h),
                                           mode, _lib.stream_ptr())
            assert rc == 0
        outs["index mode %d" % mode] = (d, i)
    for name, (d, i) in outs.items():
        i, d = i.cpu().numpy(), d.cpu().numpy()
        assert i.min() >= 0 and i.max() < m, name
        assert np.array_equal(i, io), name
        assert np.array_equal(np.isnan(d), np.isnan(do)), name
        ok = ~np.isnan(do)
        assert np.array_equal(d[ok].view(np.uint32), do[ok].view(np.uint32)), name
    # and the gradient gather can consume the result (an INT_MAX index would fault here)
    qg = q.clone().requires_grad_(True)
    dist, _ = chamfer.nn_distance(qg, ix)
    (gq,) = torch.autograd.grad(dist.nan_to_num(0.0, 0.0, 0.0).sum(), qg)
    torch.cuda.synchronize()
    assert torch.isfinite(gq[0, 0]).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices_through_the_c_abi(small_model):
    """The opt-in to > 48 kB of dynamic shared memory (tcgen05 blend GEMMs, the warp-per-query NN walk) is a
    per-DEVICE function attribute: a second GPU driven from the same process must get it too."""
    from psi_release_b200 import synthetic
    from psi_release_b200.fitting import FittingOP
    scene = synthetic.make_scene(seed=1, dim=32, num_points=60000)       # > 58 kB of boxes: the big NN opt-in too
    xh = torch.tensor(synthetic.make_body_params(scene, 3, seed=3))
    cam = torch.tensor(scene.cam_ext).unsqueeze(0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        from psi_release_b200 import chamfer
        op = FittingOP(dict(model_data=small_model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
                            contact_ids=synthetic.make_contact_ids(431, "parts"), init_lr_h=0.1, num_iter=3,
                            batch_size=3, device=dev), dict(weight_loss_rec=1, weight_loss_vposer=0.01,
                                                            weight_contact=0.1, weight_collision=0.5))
        outs.append(op.fit(xh.to(dev), cam.to(dev)).cpu())
        with torch.cuda.device(dev):
            q = torch.rand(1, 40, 3, device=dev)
            d, i = chamfer.nn_forward(q, op.s_index)                     # warp-per-query schedule (few queries)
            do, io = oracle.nn_fwd(q.cpu().numpy(), scene.points)
            assert np.array_equal(i.cpu().numpy(), io)
    assert torch.equal(outs[0], outs[1])

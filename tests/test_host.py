"""CPU tests: C-ABI surface, host-side logic, wire formats, multi-process sharding (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from psi_release_b200 import _lib
    header = open(os.path.join(ROOT, "include", "psi_b200.h")).read()
    declared = sorted(set(re.findall(r"PSI_API[^;]*?\b(psi_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    L = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.EXPORTS) == declared          # the Python binding covers the whole header
    lib = _lib.lib()
    assert lib.psi_abi_version() == 3
    assert b"workspace" in lib.psi_error_string(-2)
    assert lib.psi_sdf_num_partials(10475) == 11 and lib.psi_nn_workspace_bytes(2, 10, 30) == 2 * 30 * 8


def test_product_has_no_oracle_import_and_fails_loudly_without_cuda():
    pkg = os.path.join(ROOT, "psi-release_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn
    from psi_release_b200 import _lib, chamfer
    with pytest.raises(_lib.PsiError):
        chamfer.nn_forward(torch.zeros(1, 4, 3), torch.zeros(5, 3))     # CPU tensors are rejected
    if not torch.cuda.is_available():
        from psi_release_b200.fitting import FittingOP
        with pytest.raises(Exception):
            FittingOP(dict(batch_size=1, device="cpu", init_lr_h=0.1), {})


def test_geometry_glue_matches_oracle_restatement():
    from psi_release_b200.geometry import GeometryTransformer, VPoserDecoder, BodyParamParser
    from psi_release_b200 import synthetic
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal((9, 72)).astype(np.float32))
    x[:, 3:6] *= 0.5            # keep |aa| < pi so the round trip is the identity
    x6 = GeometryTransformer.convert_to_6D_rot(x)
    assert x6.shape == (9, 75)
    np.testing.assert_array_equal(x6.numpy(), oracle.convert_to_6D_rot(x).numpy())
    np.testing.assert_array_equal(GeometryTransformer.convert_to_3D_rot(x6).numpy(), oracle.convert_to_3D_rot(x6).numpy())
    np.testing.assert_allclose(GeometryTransformer.convert_to_3D_rot(x6).numpy(), x.numpy(), atol=3e-5)
    w = synthetic.make_vposer_weights()
    z = torch.tensor(rng.standard_normal((5, 32)).astype(np.float32))
    a = VPoserDecoder.from_weights(w).decode(z, output_type="aa").view(5, -1)
    np.testing.assert_allclose(a.detach().numpy(), oracle.VPoserDecoderOracle(w).decode(z).numpy(), atol=1e-6)
    v = torch.tensor(rng.standard_normal((2, 50, 3)).astype(np.float32))
    cam = torch.eye(4).repeat(2, 1, 1); cam[:, :3, 3] = torch.tensor([1.0, 2.0, 3.0])
    np.testing.assert_allclose(GeometryTransformer.verts_transform(v, cam).numpy(), (v + cam[:, None, :3, 3]).numpy(), atol=1e-6)
    d = BodyParamParser.body_params_encapsulate_batch(x)
    assert d["body_pose_vp"].shape == (9, 32) and d["right_hand_pose"].shape == (9, 12)
    lst = BodyParamParser.body_params_encapsulate(x)
    back = BodyParamParser.body_params_parse(lst[3], device="cpu")
    np.testing.assert_array_equal(back.numpy(), x[3:4].numpy())


def test_adam_matches_torch_optim():
    from psi_release_b200.fitting import _Adam
    torch.manual_seed(0)
    p0 = torch.randn(4, 75)
    grads = [torch.randn(4, 75) for _ in range(25)]
    p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.1)
    q = p0.clone()
    mine = _Adam(q, lr=0.1)
    for g in grads:
        p.grad = g.clone()
        opt.step()
        mine.step(g)
    np.testing.assert_allclose(q.numpy(), p.detach().numpy(), rtol=2e-5, atol=2e-6)
    mine.reset()
    assert float(mine.t) == 0 and float(mine.m.abs().sum()) == 0


def test_wire_formats_roundtrip(tmp_path):
    from psi_release_b200 import io, synthetic
    from psi_release_b200.geometry import BodyParamParser
    pts = np.random.default_rng(1).standard_normal((100, 3)).astype(np.float32)
    for binary in (True, False):
        fn = str(tmp_path / f"s{int(binary)}.ply")
        io.write_ply_vertices(fn, pts, binary=binary)
        np.testing.assert_allclose(io.read_scene_vertices(fn), pts, rtol=1e-6)
    scene = synthetic.make_scene(seed=3, dim=8, num_points=64)
    synthetic.write_scene(str(tmp_path / "sdf" / "room"), scene)
    import json
    d = json.load(open(tmp_path / "sdf" / "room.json"))
    assert d["dim"] == 8 and np.load(tmp_path / "sdf" / "room_sdf.npy").size == 512
    xh = synthetic.make_body_params(scene, 2, seed=0)
    fn = str(tmp_path / "body_gen_000000.pkl")
    io.write_body_pickle(fn, xh[1], scene.cam_ext, np.eye(3))
    x, cam_ext, cam_int = BodyParamParser.body_params_parse_fitting(io.read_body_pickle(fn), device="cpu")
    np.testing.assert_array_equal(x.numpy(), xh[1:2])
    assert cam_ext.shape == (4, 4)


def test_synthetic_world_is_deterministic_and_well_formed():
    from psi_release_b200 import synthetic
    a = synthetic.make_smplx_model(seed=1234, num_verts=300)
    b = synthetic.make_smplx_model(seed=1234, num_verts=300)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert a["posedirs"].shape == (300, 3, 486) and a["weights"].shape == (300, 55)
    np.testing.assert_allclose(a["weights"].sum(1), 1.0, atol=1e-5)
    np.testing.assert_allclose(a["J_regressor"].sum(1), 1.0, atol=1e-5)
    assert (a["weights"] != 0).sum(1).max() <= 4
    p = a["kintree_table"][0]
    assert p[0] == -1 and all(p[j] < j for j in range(1, 55))
    s = synthetic.make_scene(seed=0, dim=24, num_points=500)
    v, _ = oracle.sdf_fwd(s.sdf, s.grid_min, s.grid_max, s.points, want_grad=False)
    assert np.abs(v).max() < 0.3          # surface samples sit near the zero level set (grid res 0.26)
    xh = synthetic.make_body_params(s, 8, seed=1)
    assert xh.shape == (8, 72) and np.isfinite(xh).all()
    ids = synthetic.make_contact_ids(10475, "parts")
    assert 2000 < len(ids) < 3000 and len(np.unique(ids)) < len(ids) or len(ids) > 0


def test_shard_bounds_partition():
    from psi_release_b200.distributed import shard_bounds
    for total in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_scene_shard_plan_config3():
    from psi_release_b200.distributed import plan_scene_shards
    plan = plan_scene_shards([128, 128, 128, 128], 8)            # BASELINE configs[2]
    assert [sum(b - a for _, a, b in items) for items in plan] == [64] * 8
    assert all(len(items) == 1 for items in plan)                 # a rank never straddles scenes here
    assert plan[3] == [(1, 64, 128)]
    plan = plan_scene_shards([5, 1, 7], 4)                        # ragged: 13 bodies over 4 ranks
    flat = [(s, i) for items in plan for s, a, b in items for i in range(a, b)]
    assert flat == [(0, i) for i in range(5)] + [(1, 0)] + [(2, i) for i in range(7)]
    assert max(sum(b - a for _, a, b in it) for it in plan) - min(sum(b - a for _, a, b in it) for it in plan) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PSI_ROOT"])
from psi_release_b200.distributed import shard_rows, gather_rows, shard_bounds
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = torch.arange(7 * 72, dtype=torch.float32).view(7, 72)
mine = shard_rows(full, w, r) * 2.0          # stand-in for "fit my shard"
out = gather_rows(mine, 7)
assert torch.equal(out, full * 2.0), (r, out.shape)
dist.barrier()
if r == 0:
    print("GATHER_OK")
dist.destroy_process_group()
'''


def test_sharded_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, PSI_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


def test_oracle_fit_loop_reduces_loss_and_modes_agree_at_b1(small_model):
    """The CPU restatement of fitting_habitat.py's loop: at B=1 'batch' == 'independent'."""
    from psi_release_b200 import synthetic
    scene = synthetic.make_scene(seed=0, dim=24, num_points=600)
    xh = torch.tensor(synthetic.make_body_params(scene, 2, seed=0))
    so = oracle.SMPLXOracle(small_model)
    vp = oracle.VPoserDecoderOracle(synthetic.make_vposer_weights())
    t = torch.tensor
    W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
    kw = dict(smplx_model=so, vposer=vp, sdf=t(scene.sdf), gmin=t(scene.grid_min), gmax=t(scene.grid_max),
              scene_points=t(scene.points), contact_ids=synthetic.make_contact_ids(431, "parts"), weights=W)
    cam = t(scene.cam_ext).unsqueeze(0)
    xhr = oracle.convert_to_6D_rot(xh[:1])
    a = oracle.cal_loss(xhr, xhr.clone(), cam, loss_mode="batch", **kw)
    b = oracle.cal_loss(xhr, xhr.clone(), cam, loss_mode="independent", **kw)
    for x, y in zip(a, b):
        assert abs(float(x) - float(y)) < 1e-7
    # independent mode: fitting two bodies together == fitting them one by one
    both = oracle.fit_loop(xh, cam.expand(2, -1, -1), 3, 0.1, loss_mode="independent", **kw)
    one = oracle.fit_loop(xh[1:2], cam, 3, 0.1, loss_mode="independent", **kw)
    np.testing.assert_allclose(both[1:2].numpy(), one.numpy(), atol=2e-5)


def test_contact_query_order_is_a_permutation_grouped_by_joint():
    """fused._spatial_order: the NN group schedule wants 32 consecutive contact ids to be neighbours on the
    posed body -- ids grouped by dominant skinning joint (depth-first over the kinematic tree), kd cells
    inside a joint; duplicates of an id stay adjacent; the multiset of ids is preserved."""
    import numpy as np
    from psi_release_b200 import synthetic
    from psi_release_b200.fused import _spatial_order
    m = synthetic.make_smplx_model(seed=1234, num_verts=1500)
    parents = m["kintree_table"][0]
    ids = np.concatenate([np.arange(1500), np.array([7, 7, 900, 3])])
    o = _spatial_order(ids, m["v_template"], m["weights"], parents)
    assert sorted(o.tolist()) == sorted(ids.tolist())
    pos7 = np.nonzero(o == 7)[0]
    assert len(pos7) == 3 and pos7.max() - pos7.min() == 2               # duplicates adjacent
    owner = m["weights"].argmax(1)[o]
    changes = int((owner[1:] != owner[:-1]).sum())
    assert changes <= len(parents) - 1                                   # one contiguous run per joint
    # inside a joint the aligned 32-runs are compact kd cells: tighter than the id order
    def extent(order):
        tot = 0.0
        for j in np.unique(owner):
            v = m["v_template"][order[owner_of(order) == j]]
            for s in range(0, len(v) - 31, 32):
                tot += float((v[s:s + 32].max(0) - v[s:s + 32].min(0)).max())
        return tot
    owner_of = lambda order: m["weights"].argmax(1)[order]
    uniq = np.unique(ids)
    by_id = uniq[np.argsort(owner_of(uniq), kind="stable")]
    assert extent(np.unique(o, return_index=True)[0][np.argsort(np.unique(o, return_index=True)[1])]) < extent(by_id)
    # without skinning information it falls back to kd cells of the template
    o2 = _spatial_order(np.arange(1500), m["v_template"])
    assert sorted(o2.tolist()) == list(range(1500))


def test_rotation_restatements_against_an_independent_implementation():
    """torchgeometry 0.1.2 (requirements.txt:103) is absent, so its matrix <-> axis-angle conversions
    (cvae.py:72-91, vposer_smpl.py:153-171) are restated from the published algorithm in BOTH oracle.py and
    geometry.py -- comparing those two only compares a text with itself.  This pins them to an independent
    implementation of the same mathematical maps (scipy.spatial.transform.Rotation), including angles near 0
    and near pi and every quaternion branch of rotation_matrix_to_quaternion."""
    from scipy.spatial.transform import Rotation
    from psi_release_b200 import geometry
    rng = np.random.default_rng(3)
    axes = rng.standard_normal((400, 3))
    axes /= np.linalg.norm(axes, axis=1, keepdims=True)
    ang = np.concatenate([rng.uniform(0.05, 3.0, 340), rng.uniform(1e-4, 2e-3, 20), rng.uniform(3.0, 3.13, 40)])
    aa = (axes * ang[:, None]).astype(np.float32)
    R_ref = Rotation.from_rotvec(aa.astype(np.float64)).as_matrix()
    for mod in (oracle, geometry):
        to_R = mod.aa_to_matrix if mod is oracle else mod.angle_axis_to_rotation_matrix
        to_aa = mod.matrix_to_aa if mod is oracle else mod.rotation_matrix_to_angle_axis
        R = to_R(torch.tensor(aa)).numpy()
        # (torchgeometry divides by theta + 1e-6: measured deviation 1.5e-6)
        np.testing.assert_allclose(R, R_ref, atol=4e-6)
        back = to_aa(torch.tensor(R_ref.astype(np.float32))).numpy()
        np.testing.assert_allclose(back, Rotation.from_matrix(R_ref).as_rotvec(), atol=2e-6)
        q = (oracle.rotation_matrix_to_quaternion if mod is oracle else geometry.rotation_matrix_to_quaternion)(
            torch.tensor(R_ref.astype(np.float32))).numpy()
        qs = Rotation.from_matrix(R_ref).as_quat()           # (x, y, z, w)
        qs = np.concatenate([qs[:, 3:], qs[:, :3]], 1)
        sign = np.sign((q * qs).sum(1, keepdims=True))
        np.testing.assert_allclose(q, qs * sign, atol=1e-6)
    m00, m11, m22 = R_ref[:, 0, 0], R_ref[:, 1, 1], R_ref[:, 2, 2]
    for taken in ((m22 < 1e-6) & (m00 > m11), (m22 < 1e-6) & ~(m00 > m11), ~(m22 < 1e-6) & (m00 < -m11),
                  ~(m22 < 1e-6) & ~(m00 < -m11)):
        assert taken.sum() > 5                               # every quaternion branch is exercised by the sample
    # 6D -> matrix (cvae.py:46-55): Gram-Schmidt of the first two columns
    x6 = rng.standard_normal((50, 6)).astype(np.float32)
    Rg = oracle.rot6d_to_matrix(torch.tensor(x6)).numpy()
    a1, a2 = x6.reshape(-1, 3, 2)[:, :, 0].astype(np.float64), x6.reshape(-1, 3, 2)[:, :, 1].astype(np.float64)
    b1 = a1 / np.linalg.norm(a1, axis=1, keepdims=True)
    b2 = a2 - (b1 * a2).sum(1, keepdims=True) * b1
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    np.testing.assert_allclose(Rg, np.stack([b1, b2, np.cross(b1, b2)], -1), atol=2e-6)


def test_real_model_files_load(tmp_path, small_model):
    """smplx.create(<model dir>): the published SMPLX_NEUTRAL.npz also carries object-dtype entries, which
    np.load refuses with allow_pickle=False -- only the keys the path reads may be touched.  And the VPoser
    experiment directory of fittingconfig['vposer_ckpt_path'] (load_vposer, model_loader.py:30,43-72): newest
    snapshots/*.pt by mtime, decode-half tensors only."""
    from psi_release_b200 import body_model, shims, synthetic
    from psi_release_b200.geometry import VPoserDecoder
    md = dict(small_model)
    md["joint2num"] = np.array({"Pelvis": 0}, dtype=object)
    md["part2num"] = np.array({"Global": 0}, dtype=object)
    (tmp_path / "models" / "smplx").mkdir(parents=True)
    np.savez(str(tmp_path / "models" / "smplx" / "SMPLX_NEUTRAL.npz"), **md)
    with pytest.raises(ValueError):
        dict(np.load(str(tmp_path / "models" / "smplx" / "SMPLX_NEUTRAL.npz"), allow_pickle=False))   # the old loader
    m = body_model.create(str(tmp_path / "models"), model_type="smplx", gender="neutral", ext="npz",
                          num_pca_comps=12, batch_size=2)
    assert m.faces.shape[1] == 3 and m.left_hand_components.shape == (12, 45) and m.num_joints == 55
    np.testing.assert_array_equal(m._model_data["v_template"], small_model["v_template"])
    assert "joint2num" not in m._model_data
    bad = {k: v for k, v in small_model.items() if k != "weights"}
    np.savez(str(tmp_path / "bad.npz"), **bad)
    with pytest.raises(KeyError):
        body_model.SMPLX(str(tmp_path / "bad.npz"))
    # VPoser checkpoint directory
    snap = tmp_path / "vposer_v1_0" / "snapshots"
    snap.mkdir(parents=True)
    old, new = synthetic.make_vposer_weights(seed=1), synthetic.make_vposer_weights(seed=2)
    for name, w, t in (("TR00_E010.pt", old, 1_000_000_000), ("TR00_E096.pt", new, 1_500_000_000)):
        sd = {"module." + k: torch.tensor(v) for k, v in w.items()}      # nn.DataParallel prefix, as the trainer saves
        sd["module.bodyprior_enc_fc1.weight"] = torch.zeros(3, 3)
        torch.save(sd, str(snap / name))
        os.utime(str(snap / name), (t, t))
    dec = VPoserDecoder.from_checkpoint_dir(str(tmp_path / "vposer_v1_0"))
    np.testing.assert_array_equal(dec.bodyprior_dec_out.weight.detach().numpy(), new["bodyprior_dec_out.weight"])
    vp, ps = shims.load_vposer(str(tmp_path / "vposer_v1_0"), vp_model="snapshot")
    assert ps is None and vp.decode(torch.zeros(2, 32), output_type="aa").shape == (2, 1, 21, 3)
    with pytest.raises(FileNotFoundError):
        VPoserDecoder.from_checkpoint_dir(str(tmp_path / "nothing_here"))


def test_shims_route_only_the_reference_call_shape():
    """install() aliases the reference's import names and routes F.grid_sample's PSI call shape to the psi SDF
    op; anything else (CPU tensors, 4-D inputs, zeros padding) reaches torch's own grid_sample; uninstall()
    restores everything."""
    import importlib
    import torch.nn.functional as F
    from psi_release_b200 import body_model, chamfer, shims
    orig = F.grid_sample
    shims.install()
    try:
        assert importlib.import_module("smplx") is body_model
        assert importlib.import_module("chamfer_pytorch.dist_chamfer").chamferDist is chamfer.chamferDist
        assert importlib.import_module("human_body_prior.tools.model_loader").load_vposer is shims.load_vposer
        assert F.grid_sample is not orig
        vol, grid = torch.rand(2, 1, 4, 4, 4), torch.rand(2, 5, 1, 1, 3) * 2 - 1
        assert not shims._is_psi_sdf_call(vol, grid, "bilinear", "border", None)          # CPU tensors: torch's kernel
        np.testing.assert_array_equal(F.grid_sample(vol, grid, padding_mode="border", align_corners=True).numpy(),
                                      orig(vol, grid, padding_mode="border", align_corners=True).numpy())
        img, g2 = torch.rand(1, 3, 8, 8), torch.rand(1, 4, 4, 2) * 2 - 1
        np.testing.assert_array_equal(F.grid_sample(img, g2, align_corners=False).numpy(), orig(img, g2, align_corners=False).numpy())

        meta_vol, meta_grid = torch.empty(2, 1, 4, 4, 4, device="meta"), torch.empty(2, 5, 1, 1, 3, device="meta")
        assert not shims._is_psi_sdf_call(meta_vol, meta_grid, "bilinear", "border", None)
    finally:
        shims.uninstall()
    assert F.grid_sample is orig and "smplx" not in sys.modules and "chamfer_pytorch" not in sys.modules
    with shims.routed_grid_sample():
        assert F.grid_sample is not orig
    assert F.grid_sample is orig


def test_joint_coherent_vertex_order_is_a_consistent_relabelling():
    """The fused loop runs on a re-ordered copy of the model (body_model.SMPLX.coherent_handle): the order is a
    permutation of the vertices, it makes the vertices of a warp share their skinning joints, and permuting every
    per-vertex constant of the model with it only relabels the output vertices (checked on the CPU restatement)."""
    import numpy as np
    import torch
    from oracle import oracle
    from psi_release_b200 import synthetic
    from psi_release_b200.fused import _spatial_order
    m = synthetic.make_smplx_model(seed=1234, num_verts=431)
    V = 431
    parents = np.asarray(m["kintree_table"])[0].astype(np.int64)
    perm = _spatial_order(np.arange(V), m["v_template"], m["weights"], parents).astype(np.int64)
    assert perm.shape == (V,) and np.array_equal(np.sort(perm), np.arange(V))
    owner = m["weights"].argmax(1)
    per_warp = lambda o: np.mean([len(set(o[i:i + 32])) for i in range(0, V - 31, 32)])
    assert per_warp(owner[perm]) < 0.4 * per_warp(owner)           # 3-7 joints per warp instead of 22-26
    mp = dict(m)
    for k in ("v_template", "shapedirs", "posedirs", "weights"):
        mp[k] = m[k][perm]
    mp["J_regressor"] = m["J_regressor"][:, perm]
    g = torch.Generator().manual_seed(0)
    B = 3
    kw = dict(body_pose=torch.randn(B, 63, generator=g) * 0.3, transl=torch.randn(B, 3, generator=g),
              global_orient=torch.randn(B, 3, generator=g) * 0.5, betas=torch.randn(B, 10, generator=g),
              left_hand_pose=torch.randn(B, 12, generator=g) * 0.3, right_hand_pose=torch.randn(B, 12, generator=g) * 0.3)
    va, ja = oracle.SMPLXOracle(m)(**kw)
    vb, jb = oracle.SMPLXOracle(mp)(**kw)
    assert float((va[:, perm] - vb).abs().max()) < 1e-5
    assert float((ja - jb).abs().max()) < 1e-5

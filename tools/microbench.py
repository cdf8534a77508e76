"""Per-kernel timings at the BASELINE config-2 sizes (B=64, V=10475, M=50000, D=256).
CUDA events on the launching stream, warm-up, L2 flush between repetitions."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psi_release_b200 import body_model, chamfer, sdf as sdf_mod, synthetic  # noqa: E402


def timeit(fn, reps=5, warmup=2, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--verts", type=int, default=10475)
    ap.add_argument("--points", type=int, default=50000)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--ref", action="store_true", help="also time the reference chamfer.cu build")
    args = ap.parse_args()
    B, V, M, D = args.batch, args.verts, args.points, args.dim
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {"B": B, "V": V, "M": M, "D": D, "gpu": torch.cuda.get_device_name(0)}

    scene = synthetic.make_scene(seed=0, dim=D, num_points=M)
    model = synthetic.make_smplx_model(seed=1234, num_verts=V)
    m = body_model.create(model_data=model, num_pca_comps=12, batch_size=B).cuda()
    h = m.handle()
    rng = np.random.default_rng(0)
    betas = torch.tensor(rng.standard_normal((B, 20)).astype(np.float32), device=dev)
    pose = torch.tensor((rng.standard_normal((B, 165)) * 0.3).astype(np.float32), device=dev)
    transl = torch.tensor(rng.uniform(-1, 1, (B, 3)).astype(np.float32), device=dev)
    cam = torch.tensor(scene.cam_ext, device=dev).unsqueeze(0)
    verts, _ = body_model.lbs(betas, pose, h, transl=transl, cam=cam)
    pts = torch.tensor(scene.points, device=dev)

    med, mn = timeit(lambda: chamfer.nn_forward(verts, pts), flush=flush)
    pairs = B * V * M
    out["nn_fwd_ms"] = med
    out["nn_fwd_min_ms"] = mn
    out["nn_pairs_per_s"] = pairs / (mn * 1e-3)
    out["nn_alg_GBps"] = (B * V * 12 + M * 12 + B * V * 8) / (mn * 1e-3) / 1e9

    ix = chamfer.SceneIndex(pts)
    med, mn = timeit(lambda: chamfer.nn_forward(verts, ix), flush=flush)
    out["nn_index_ms"] = med
    out["nn_index_min_ms"] = mn
    di, ii = chamfer.nn_forward(verts, ix)
    db, ib = chamfer.nn_forward(verts, pts)
    out["nn_index_equal_bruteforce"] = bool(torch.equal(ii, ib) and torch.equal(di.view(torch.int32), db.view(torch.int32)))

    # index query schedules: 1 = warp per query, 2 = thread per query; with / without coherent order + hints
    from psi_release_b200 import _lib
    L = _lib.lib()
    vt = model["v_template"].astype(np.float64)
    cell = ((vt - vt.min(0)) / (vt.max(0) - vt.min(0)) * 1023 + 0.5).astype(np.int64)
    def spread(v):
        v = v & 0x3FF; v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249
    rank = spread(cell[:, 0]) | (spread(cell[:, 1]) << 1) | (spread(cell[:, 2]) << 2)
    sel_sorted = torch.tensor(np.argsort(rank, kind="stable").astype(np.int32), device=dev)
    from psi_release_b200.fused import _spatial_order
    sel_jkd = torch.tensor(_spatial_order(np.arange(V), model["v_template"], model["weights"], model["kintree_table"][0]), device=dev)
    dq = torch.empty(B, V, device=dev); iq = torch.empty(B, V, dtype=torch.int32, device=dev)
    hint = torch.full((B, V), -1, dtype=torch.int32, device=dev)
    def run(mode, sel, hnt):
        rc = L.psi_nn_index_query_mode(ix.h, _lib.ptr(verts), V * 3, B, V, _lib.ptr(sel), _lib.ptr(dq), _lib.ptr(iq), _lib.ptr(hnt), mode, _lib.stream_ptr())
        assert rc == 0
    for mode in (1, 2, 3):
        for name, sel in (("natural", None), ("sorted", sel_sorted), ("jointkd", sel_jkd)):
            out["nn_mode%d_%s_nohint_ms" % (mode, name)] = timeit(lambda: run(mode, sel, None), flush=flush)[0]
            hint.fill_(-1); run(mode, sel, hint)
            out["nn_mode%d_%s_hint_ms" % (mode, name)] = timeit(lambda: run(mode, sel, hint), flush=flush)[0]
    run(3, sel_sorted, hint)
    out["nn_mode3_equal_bruteforce"] = bool(torch.equal(iq, ib[:, sel_sorted.long()]) and torch.equal(dq.view(torch.int32), db[:, sel_sorted.long()].view(torch.int32)))
    run(2, None, None)
    out["nn_mode2_equal_bruteforce"] = bool(torch.equal(iq, ib) and torch.equal(dq.view(torch.int32), db.view(torch.int32)))

    med, mn = timeit(lambda: body_model.lbs(betas, pose, h, transl=transl, cam=cam), flush=flush)
    out["lbs_fwd_ms"] = med
    out["lbs_fwd_min_ms"] = mn
    out["lbs_alg_GBps"] = (h.nbytes() + B * (740 + V * 12)) / (mn * 1e-3) / 1e9
    out["lbs_fwd_TFLOPs"] = 49.3e6 * B / (mn * 1e-3) / 1e12

    b2, p2 = betas.clone().requires_grad_(True), pose.clone().requires_grad_(True)
    v2, _ = body_model.lbs(b2, p2, h, transl=transl, cam=cam)
    g = torch.randn_like(v2)
    med, mn = timeit(lambda: torch.autograd.grad(v2, (b2, p2), g, retain_graph=True), flush=flush)
    out["lbs_bwd_ms"] = med
    out["lbs_bwd_min_ms"] = mn

    sc = sdf_mod.SceneSDF(scene.sdf, scene.grid_min, scene.grid_max)
    med, mn = timeit(lambda: sdf_mod.sdf_forward(sc, verts, want_grad=True, want_partials=True), flush=flush)
    out["sdf_fwd_ms"] = med
    out["sdf_fwd_min_ms"] = mn
    out["sdf_alg_GBps"] = B * V * (32 + 12 + 4 + 12) / (mn * 1e-3) / 1e9

    if args.ref:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
        import build_ref
        ref = build_ref.load_ref()
        if ref is not None:
            Bs = min(B, 8)
            a = verts[:Bs].contiguous()
            bb = pts.unsqueeze(0).repeat(Bs, 1, 1).contiguous()
            d1 = torch.zeros(Bs, V, device=dev); d2 = torch.zeros(Bs, M, device=dev)
            i1 = torch.zeros(Bs, V, dtype=torch.int32, device=dev); i2 = torch.zeros(Bs, M, dtype=torch.int32, device=dev)

            def run_ref():
                ref.forward(a, bb, d1, d2, i1, i2)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); run_ref(); torch.cuda.synchronize(); t1 = time.perf_counter()
            t0 = time.perf_counter(); run_ref(); torch.cuda.synchronize(); t1 = time.perf_counter()
            out["ref_chamfer_both_dirs_ms_B%d" % Bs] = (t1 - t0) * 1e3
            out["ref_pairs_per_s"] = 2 * Bs * V * M / (t1 - t0)
            md, mi = chamfer.nn_forward(a, pts)
            out["ref_idx_equal"] = bool(torch.equal(mi, i1))
            out["ref_dist_bits_equal"] = bool(torch.equal(md.view(torch.int32), d1.view(torch.int32)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

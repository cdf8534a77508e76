"""bench.py workloads beyond the headline fit loop (BASELINE.json configs[3] and configs[4]).

    python bench.py --workload train_s2 [--gpus N] [--geometry on|off]      (torchrun for N > 1)
    python bench.py --workload rooms    [--gpus N] [--rooms 7 --samples 128]

Bench-tree code: the CVAE below only has the SHAPE of the reference's HumanCVAES2 (source/cvae.py:341-400,
source/net_layers.py:28-234: two conditional VAEs, each with a ResNet18 trunk cut after layer2 over a
2-channel 128x128 depth+semantics image, 15.7 M parameters) with random initial weights -- the reference
keeps it a stock torch module (north_star) and no trained checkpoint exists offline (`data/resnet18.pth` is
a legacy pickle that must not be loaded).  What is measured is the product: the psi
geometry block inside the training loss (training.SceneLossBlock / TrainStep) and the rooms pipeline
(pipeline.fit_rooms).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.nn as nn


# ------------------------------------------------------------------ HumanCVAES2-shaped network
class _ResBlock(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(n, n), nn.Linear(n, n), nn.LeakyReLU()

    def forward(self, x):
        return x + self.act(self.fc2(self.act(self.fc1(x))))


def _trunk(in_dim):
    import torchvision
    r = torchvision.models.resnet18(weights=None)
    return nn.Sequential(nn.Conv2d(in_dim, 64, 7, 2, 3, bias=False), r.bn1, r.relu, r.maxpool, r.layer1, r.layer2)


class _CondVAE(nn.Module):
    """One conditional VAE: scene image -> ResNet18 trunk (128 ch at 16x16) -> conv -> fc = z_s; the body part
    (and, for the local-pose VAE, the torso) enter through linear layers; two ResBlocks mix; mean / log-var heads;
    decoder = linear + two ResBlocks + linear."""

    def __init__(self, out_dim, f_dim, n_cond, zdim=32, hidden=256):          # train_s2.py:59: latentD 256
        super().__init__()
        self.zdim, self.out_dim = zdim, out_dim
        self.resnet = _trunk(2)
        self.conv = nn.Conv2d(128, f_dim, 3, 1, 1)
        self.fc = nn.Linear(f_dim * 16 * 16, hidden)
        self.torso_linear = nn.Linear(3, hidden)
        self.pose_linear = nn.Linear(72, hidden) if n_cond == 3 else None
        self.encode = nn.Sequential(_ResBlock(n_cond * hidden), _ResBlock(n_cond * hidden))
        self.mean_linear = nn.Linear(n_cond * hidden, zdim)
        self.log_var_linear = nn.Linear(n_cond * hidden, zdim)
        self.decode = nn.Sequential(nn.Linear((n_cond - 1) * hidden + zdim, f_dim), _ResBlock(f_dim), _ResBlock(f_dim),
                                    nn.Linear(f_dim, out_dim))

    def scene_feature(self, xs):
        return self.fc(self.conv(self.resnet(xs)).flatten(1))

    def forward(self, xs, eps, torso, pose=None):
        z_s = self.scene_feature(xs)
        z_g = self.torso_linear(torso)
        cond = [z_s, z_g] if pose is None else [self.pose_linear(pose), z_g, z_s]
        h = self.encode(torch.cat(cond, 1))
        mu, logvar = self.mean_linear(h), self.log_var_linear(h)
        z = mu + eps * torch.exp(0.5 * logvar)
        dec_in = [z, z_s] if pose is None else [z, z_g, z_s]
        return self.decode(torch.cat(dec_in, 1)), mu, logvar

    def sample(self, xs, z, torso=None):
        z_s = self.scene_feature(xs)
        dec_in = [z, z_s] if torso is None else [z, self.torso_linear(torso), z_s]
        return self.decode(torch.cat(dec_in, 1))


class CVAES2Shape(nn.Module):
    """model_h(xhnr [B,75], eps_g, eps_l, xs [B,2,128,128]) -> (xhnr_rec, mu_g, logsigma2_g, mu_l, logsigma2_l):
    the call signature TrainOP.cal_loss uses (train_s2.py:118-120).  The reconstruction is a residual around the
    input so that a randomly initialised network produces valid bodies (rotations stay rotations)."""

    def __init__(self, residual=0.05):
        super().__init__()
        self.trans_vae = _CondVAE(out_dim=3, f_dim=32, n_cond=2)
        self.pose_vae = _CondVAE(out_dim=72, f_dim=128, n_cond=3)
        self.residual = residual

    def forward(self, x_body, eps_g, eps_l, xs):
        x_g, x_l = x_body[:, :3], x_body[:, 3:]
        g_rec, mu_g, ls_g = self.trans_vae(xs, eps_g, x_g)
        g_rec = x_g + self.residual * g_rec
        l_rec, mu_l, ls_l = self.pose_vae(xs, eps_l, g_rec, x_l)
        return torch.cat([g_rec, x_l + self.residual * l_rec], 1), mu_g, ls_g, mu_l, ls_l

    def sample(self, xs, z_g, z_l, prior):
        """Generation (test_habitat_s2.py:196-229 shape): decode latent samples; `prior` [B,75] stands in for the
        trained decoder's mean (no checkpoint offline)."""
        g = prior[:, :3] + self.residual * self.trans_vae.sample(xs, z_g)
        l = prior[:, 3:] + self.residual * self.pose_vae.sample(xs, z_l, torso=g)
        return torch.cat([g, l], 1)


# --------------------------------------------------------------------------------------- helpers
def _dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    return world, rank, local, torch.device("cuda", local)


def _timed_steps(fn, steps, flush, dev, barrier):
    tot, out = 0.0, None
    barrier()
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        tot += e0.elapsed_time(e1)
    barrier()
    return tot, out


def _max_over_ranks(ms, dist, world, dev):
    if world == 1:
        return ms, [ms]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    per = [float(x.item()) for x in allt]
    return max(per), per


# ----------------------------------------------------------------------------- configs[3]: train_s2
def run_train_s2(args):
    """CVAE stage-2 training step (train_s2.py:102-204 loss, :253-262 step): HumanCVAES2-shaped network under bf16
    autocast, DDP over N GPUs, B = 128 samples per GPU (train_s2.py:360), samples spread over a 4-scene bank of
    256^3 SDFs + 50 000-point clouds, geometry block (VPoser decode -> SMPL-X -> contact NN -> SDF) in FP32 on the
    psi kernels.  --geometry off = the first 75 % of the epochs, where the reference computes the block and
    multiplies it by zero and this implementation skips it."""
    import bench
    from psi_release_b200 import _lib, body_model, synthetic, training
    from psi_release_b200.geometry import VPoserDecoder
    world, rank, local, dev = _dist_env()
    dist = bench.init_dist(dev) if world > 1 else None
    B = 128
    torch.manual_seed(1234)
    model = synthetic.make_smplx_model(seed=1234, num_verts=bench.NUM_VERTS)
    scenes = [synthetic.make_scene(seed=s, dim=bench.SDF_DIM, num_points=bench.NUM_POINTS) for s in range(4)]
    scene_ids = [(i + rank) % 4 for i in range(B)]
    xh = np.concatenate([synthetic.make_body_params(scenes[s], 1, seed=1000 * rank + i) for i, s in enumerate(scene_ids)])
    xh[:, 2] = np.abs(xh[:, 2]) + 1.5                                 # in front of the camera (normalize_global_T divides by z)
    cid = synthetic.make_contact_ids(bench.NUM_VERTS, "parts")
    blk = training.SceneLossBlock(body_model.create(model_data=model, num_pca_comps=12, batch_size=B),
                                  VPoserDecoder.from_weights(synthetic.make_vposer_weights()), scenes, cid,
                                  weight_loss_vposer=1e-3, weight_contact=1e-1, weight_collision=1e-1, device=dev)   # train_s2.py:372-375
    net = CVAES2Shape().to(dev)
    nparams = sum(p.numel() for p in net.parameters())
    model_h = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    ts = training.TrainStep(model_h, blk, opt, weight_loss_rec_h=1.0, weight_loss_kl=0.1, epochs=30)
    ep = 29 if args.geometry == "on" else 3
    host = dict(xs=(torch.rand(B, 2, 128, 128) * 2 - 1).pin_memory(), xh=torch.tensor(xh).pin_memory(),
                eps_g=torch.randn(B, 32).pin_memory(), eps_l=torch.randn(B, 32).pin_memory(),
                cam=torch.stack([torch.tensor(scenes[s].cam_ext) for s in scene_ids]).pin_memory(),
                cam_int=torch.tensor([[500.0, 0, 320.0], [0, 500.0, 240.0], [0, 0, 1.0]]).repeat(B, 1, 1).pin_memory(),
                max_d=torch.full((B,), 6.0).pin_memory())
    devb = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    L = _lib.lib()

    def step_dev():
        return ts.step(devb["xs"], devb["xh"], devb["eps_g"], devb["eps_l"], devb["cam"], devb["cam_int"], devb["max_d"],
                       scene_ids, ep=ep)

    res_pin = torch.empty(7, dtype=torch.float32).pin_memory()

    def step_e2e():
        b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        terms = ts.step(b["xs"], b["xh"], b["eps_g"], b["eps_l"], b["cam"], b["cam_int"], b["max_d"], scene_ids, ep=ep)
        res_pin.copy_(terms, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return res_pin

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    first = None
    for _ in range(max(args.warmup, 3)):
        t = step_dev()
        first = t if first is None else first
    c0 = L.psi_launch_count()
    step_dev()
    psi_launches = int(L.psi_launch_count() - c0)
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev_r, last = _timed_steps(step_dev, args.steps, flush, dev, barrier)
    ms_e2e_r, _ = _timed_steps(step_e2e, args.steps, flush, dev, barrier)
    clocks = sampler.stop() if rank == 0 else None
    ms_dev, per_rank = _max_over_ranks(ms_dev_r, dist, world, dev)
    ms_e2e, _ = _max_over_ranks(ms_e2e_r, dist, world, dev)
    assert torch.isfinite(last).all()
    # where the step goes: the geometry block alone (forward + backward through the psi kernels)
    geo_ms = None
    if args.geometry == "on":
        from psi_release_b200.geometry import GeometryTransformer
        x = devb["xh"].clone().requires_grad_(True)

        def geo():
            lc, lv, lp = blk(x, devb["cam"], scene_ids, ep=29, epochs=30)
            (g,) = torch.autograd.grad(lc + lv + lp, x)
            return g
        geo()
        t, _ = _timed_steps(geo, 5, flush, dev, barrier)
        geo_ms = t / 5
        del GeometryTransformer
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    h2d = int(sum(v.numel() * v.element_size() for v in host.values())) * world
    out = {"metric": "train_s2 samples/sec (CVAE stage-2 step, geometry block %s)" % args.geometry,
           "value": B * world * args.steps / (ms_dev * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16 network (autocast) + f32 geometry block", "data": "synthetic",
           "config": {"workload": "configs[3]: train_s2 step, HumanCVAES2-shaped network (2 x ResNet18 trunk, %.1f M parameters, "
                                  "random init), %d samples/GPU over a 4-scene bank (256^3 SDF, 50 000 points each), "
                                  "10 475-vertex SMPL-X, %d contact ids, Adam lr 1e-4 (train_s2.py:366), geometry block %s"
                                  % (nparams / 1e6, B, len(cid), args.geometry),
                      "samples_per_gpu": B, "global_batch": B * world, "parallelism": "DDP dp%d (NCCL all-reduce of the CVAE gradients)" % world,
                      "l2": "flushed between timed steps (256 MiB write)"},
           "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 28 * world,
                   "api": "psi_release_b200.training.TrainStep.step (pinned host batch in, 7 loss terms out)"},
           "gpu_launches": psi_launches * args.steps * world, "psi_launches_per_step": psi_launches,
           "geometry_block_fwd_bwd_ms": geo_ms, "per_rank_ms_per_step": [x / args.steps for x in per_rank],
           "loss_first_warmup_step": [float(x) for x in first], "loss_last_step": [float(x) for x in last], "clocks": clocks}
    time.sleep(0.3)
    print(json.dumps(out), flush=True)


# -------------------------------------------------------------------------------- configs[4]: rooms
def run_rooms(args):
    """MP3D-R style pipeline (test_habitat_s2.py:155-229 -> fitting_habitat.py:169-218 -> utils_eval_collision_habitat.py):
    R rooms x n samples: generate (CVAE-shaped network decode) -> VPoser decode -> SMPL-X -> SDF / contact fit
    (300 Adam iterations) -> scores, sharded over the ranks by pipeline.fit_rooms (plan_scene_shards), one
    all-gather of fitted vectors and scores.  Rooms live on the HOST: every step uploads each room's 64 MiB SDF and
    builds its NN index, so `value` is already end to end."""
    import bench
    from psi_release_b200 import _lib, body_model, pipeline, synthetic
    from psi_release_b200.geometry import GeometryTransformer
    world, rank, local, dev = _dist_env()
    dist = bench.init_dist(dev) if world > 1 else None
    R, n = args.rooms, args.samples
    torch.manual_seed(7)
    model = synthetic.make_smplx_model(seed=1234, num_verts=bench.NUM_VERTS)
    rooms = [synthetic.make_scene(seed=20 + r, dim=bench.SDF_DIM, num_points=bench.NUM_POINTS) for r in range(R)]
    net = CVAES2Shape().to(dev).eval()
    xs_room = [(torch.rand(1, 2, 128, 128, generator=torch.Generator().manual_seed(r)) * 2 - 1) for r in range(R)]
    gen_ms = [0.0]

    def generate(r, count):
        """count bodies for room r: the same on every rank (seeded latents), decoded on this rank's GPU."""
        g = torch.Generator().manual_seed(900 + r)
        z_g, z_l = torch.randn(count, 32, generator=g).to(dev), torch.randn(count, 32, generator=g).to(dev)
        prior = GeometryTransformer.convert_to_6D_rot(torch.tensor(synthetic.make_body_params(rooms[r], count, seed=300 + r)).to(dev))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            x6 = net.sample(xs_room[r].to(dev).expand(count, -1, -1, -1), z_g, z_l, prior)
        xh = GeometryTransformer.convert_to_3D_rot(x6.float())
        e1.record()
        e1.synchronize()
        gen_ms[0] += e0.elapsed_time(e1)
        return xh.cpu()

    base = dict(body_mesh_model=body_model.create(model_data=model, num_pca_comps=12, batch_size=n),
                vposer_weights=synthetic.make_vposer_weights(), contact_ids=synthetic.make_contact_ids(bench.NUM_VERTS, "full"),
                init_lr_h=0.1, num_iter=args.iters, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        return pipeline.fit_rooms(rooms, generate, n, base, bench.LOSS)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    gen_ms[0] = 0.0
    c0 = L.psi_launch_count()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_r, res = _timed_steps(step, args.steps, flush, dev, barrier)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(L.psi_launch_count() - c0)
    ms, per_rank = _max_over_ranks(ms_r, dist, world, dev)
    assert res["fitted"].shape == (R * n, 72) and torch.isfinite(res["fitted"]).all()
    nc, ct = float(res["non_collision"].mean()), float(res["contact"].mean())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    v = R * n * args.steps / (ms * 1e-3)
    out = {"metric": "rooms pipeline bodies/sec (generate -> VPoser decode -> SMPL-X -> %d-iteration SDF/contact fit -> scores)" % args.iters,
           "value": v, "unit": "bodies/s", "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 2)),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (generator under bf16 autocast)",
           "data": "synthetic",
           "config": {"workload": "configs[4]: %d synthetic rooms x %d samples (256^3 SDF, 50 000 points per room), CVAE-shaped generator "
                                  "(random init, residual around a synthetic prior), 10 475-vertex SMPL-X, full-body contact, %d Adam iterations, "
                                  "rooms sharded over ranks (pipeline.fit_rooms)" % (R, n, args.iters),
                      "rooms": R, "samples_per_room": n, "global_bodies": R * n, "parallelism": "dp%d, bodies of all rooms cut into contiguous runs" % world,
                      "l2": "flushed between timed steps (256 MiB write)"},
           "e2e": {"value": v, "unit": "bodies/s", "ms_per_step": ms / args.steps,
                   "h2d_bytes_per_step": int(sum(r.sdf.nbytes + r.points.nbytes for r in rooms)),
                   "d2h_bytes_per_step": R * n * (72 + 2) * 4, "api": "psi_release_b200.pipeline.fit_rooms (rooms on the host)"},
           "gpu_launches": launches, "generate_ms_per_step_rank0": gen_ms[0] / args.steps,
           "per_rank_ms_per_step": [x / args.steps for x in per_rank],
           "scores": {"non_collision_mean": nc, "contact_mean": ct}, "clocks": clocks}
    time.sleep(0.3)
    print(json.dumps(out), flush=True)

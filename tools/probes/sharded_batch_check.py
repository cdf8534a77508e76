"""torchrun check (one process per GPU): loss_mode='batch' sharded over the ranks -- every rank fits its bodies while the
loop exchanges the penetration counts through CUDA-IPC-mapped peer memory -- equals ONE context fitting the union batch,
bit for bit.   torchrun --nproc-per-node 2 tools/probes/sharded_batch_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from psi_release_b200 import synthetic  # noqa: E402
from psi_release_b200.distributed import gather_rows, shard_bounds  # noqa: E402
from psi_release_b200.fitting import FittingOP  # noqa: E402

W = dict(weight_loss_rec=1, weight_loss_vposer=0.01, weight_contact=0.1, weight_collision=0.5)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V = int(os.environ.get("PSI_CHECK_VERTS", "431"))
model = synthetic.make_smplx_model(seed=1234, num_verts=V)
scene = synthetic.make_scene(seed=1, dim=32, num_points=3000)
B = 7 * world + 1
xh = torch.tensor(synthetic.make_body_params(scene, B, seed=3)).cuda()
cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
cfg = dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(), contact_ids=synthetic.make_contact_ids(V, "parts"),
           init_lr_h=0.1, num_iter=40, device="cuda", loss_mode="batch")
lo, hi = shard_bounds(B, world, rank)
op = FittingOP(dict(cfg, batch_size=hi - lo), W)
op.connect_batch_shards(B)
for rep in range(2):
    mine = op.fit(xh[lo:hi], cam)
    got = gather_rows(mine, B)
    err = int(op.trace("exchange")[0, 1])
    if rank == 0:
        ref = FittingOP(dict(cfg, batch_size=B), W).fit(xh, cam)
        same = bool(torch.equal(got, ref))
        print("sharded batch-coupled fit, %d ranks, %d bodies, rep %d: bit-identical to the union batch: %s (exchange error flag %d, max |diff| %.3e)"
              % (world, B, rep, same, err, float((got - ref).abs().max())), flush=True)
        assert same and err == 0
dist.barrier()
dist.destroy_process_group()

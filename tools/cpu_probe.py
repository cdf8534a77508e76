"""Where does the CPU reference loop spend its time, and with how many threads is it fastest?"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle
import bench

class A: batch=64; iters=300
o, xh, cam, kw = bench._cpu_setup(A)
res = {}
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 1): continue
    torch.set_num_threads(nt)
    B = 64
    x = xh[:B]
    xhr = o.convert_to_6D_rot(x); rec = xhr.clone().requires_grad_(True)
    t0 = time.perf_counter()
    xh3 = o.convert_to_3D_rot(rec); pose = kw["vposer"].decode(xh3[:, 16:48])
    v, _ = kw["smplx_model"](body_pose=pose, transl=xh3[:, :3], global_orient=xh3[:, 3:6], betas=xh3[:, 6:16],
                             left_hand_pose=xh3[:, 48:60], right_hand_pose=xh3[:, 60:])
    v = o.verts_transform(v, cam.expand(B, -1, -1))
    t1 = time.perf_counter()
    d, i = o.nn_fwd(v.detach().numpy(), kw["scene_points"].numpy())
    t2 = time.perf_counter()
    s = o.sdf_lookup_torch(kw["sdf"], kw["gmin"], kw["gmax"], v)
    t3 = time.perf_counter()
    (v.sum() + s.sum()).backward()
    t4 = time.perf_counter()
    res[nt] = dict(lbs_fwd=t1 - t0, nn=t2 - t1, sdf=t3 - t2, bwd=t4 - t3)
    print(nt, res[nt], flush=True)
print(json.dumps(res))

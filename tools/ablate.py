"""In-graph marginal cost of every kernel of the fitting iteration: run the BASELINE configs[1] fit with
PSI_SKIP_KERNEL=<name> (libpsi_b200 drops that launch; results are garbage, only the time matters) and
report iteration time minus the full iteration's.  Eager per-launch timings (psi_fit_profile) include
~4 us of launch overhead per kernel; this does not.

    python tools/ablate.py            # table on stdout (one subprocess per kernel)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["", "fit_linear", "lbs_pose_fwd", "lbs_blend_fwd_bf3", "lbs_skin_sdf_fwd", "nn_index_group", "lbs_vertex_bwd_fit",
           "lbs_dcoef_bf3", "lbs_reduce2", "lbs_pose_bwd", "fit_step"]

CHILD = r"""
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
import bench
from psi_release_b200 import synthetic
from psi_release_b200.fitting import FittingOP
args = bench.parse()
model, scene, xh = bench.make_world(args, 0)
op = FittingOP(dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
                    contact_ids=synthetic.make_contact_ids(bench.NUM_VERTS, "full"), init_lr_h=0.1, num_iter=300,
                    batch_size=64, device="cuda", use_cuda_graph=True), bench.LOSS)
x = torch.tensor(xh).cuda(); cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
for _ in range(3): op.fit(x, cam)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): op.fit(x, cam)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"us_per_iteration": e0.elapsed_time(e1) / 900 * 1e3}))
""" % ROOT

if __name__ == "__main__":
    base = None
    rows = []
    for k in KERNELS:
        # one graph launch per iteration: with the whole-loop WHILE graph a skipped fit_step would never count the loop down
        env = dict(os.environ, PSI_FIT_LOOP="replay")
        if k:
            env["PSI_SKIP_KERNEL"] = k
        try:
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, cwd=ROOT, timeout=120)
        except subprocess.TimeoutExpired:
            print(k, "TIMED OUT")
            continue
        try:
            us = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])["us_per_iteration"]
        except Exception:
            print(k, "FAILED", out.stderr[-400:])
            continue
        if not k:
            base = us
        rows.append((k or "(full iteration)", us))
    for k, us in rows:
        print("%-22s %8.1f us/iteration   marginal %7.1f us" % (k, us, (base - us) if base else 0.0))

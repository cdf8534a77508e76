"""Event counters of the NN group walk inside the fitting loop (needs a library built with
PSI_EXTRA_NVCC_FLAGS=-DPSI_NN_STATS python -m psi_release_b200.build --force)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from psi_release_b200 import _lib, synthetic
from psi_release_b200.fitting import FittingOP
args = bench.parse()
model, scene, xh = bench.make_world(args, 0)
op = FittingOP(dict(model_data=model, scene=scene, vposer_weights=synthetic.make_vposer_weights(),
                    contact_ids=synthetic.make_contact_ids(bench.NUM_VERTS, "full"), init_lr_h=0.1, num_iter=300,
                    batch_size=64, device="cuda", use_cuda_graph=True), bench.LOSS)
x = torch.tensor(xh).cuda(); cam = torch.tensor(scene.cam_ext).unsqueeze(0).cuda()
L = ctypes.CDLL(_lib.lib_path())
buf = (ctypes.c_ulonglong * 8)()
op.fit(x, cam)
L.psi_debug_nn_stats(buf, 1)
op.fit(x, cam)
L.psi_debug_nn_stats(buf, 1)
g = buf[0]
names = ["groups", "seed visits", "mega batches", "super batches", "leaf tests", "sweep visits"]
for n, v in zip(names, buf):
    print("%-14s %12d   per group %.2f" % (n, v, v / max(g, 1)))

"""Per-chunk pipeline time stamps of CTA 0 of the tcgen05 blend GEMM (library built with -DPSI_TC5_TRACE).

    PSI_B200_LIB=.../libpsi_b200_trace.so python tools/tc5_trace.py

Events per chunk g: 0 raw producer issued TMA, 1 worker saw the raw tile, 2 worker saw the operand slot free,
3 worker finished split + proxy fence, 4 issuer saw the operand slot, 5 issuer saw the per-body tile, 6 issuer done issuing.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psi_release_b200 import _lib, body_model as bm, synthetic  # noqa: E402

model = synthetic.make_smplx_model(seed=1234, num_verts=10475)
m = bm.create(model_data=model, num_pca_comps=12, batch_size=64).cuda()
h = m.handle(torch.device("cuda", 0))
B = 64
betas = torch.randn(B, 20, device="cuda")
pose = torch.randn(B, 165, device="cuda") * 0.3
for _ in range(5):
    v, _ = bm.lbs(betas, pose, h)                 # the last call's forward blend GEMM leaves its stamps
torch.cuda.synchronize()
L = _lib.lib()
buf = (ctypes.c_ulonglong * (8 * 64))()
rc = ctypes.CDLL(_lib.lib_path()).psi_debug_tc5_trace(buf)
t = np.array(buf[:], dtype=np.int64).reshape(8, 64)
t0 = t[0, 0]
names = ["tma_issue", "wk_raw_ok", "wk_op_free", "wk_split_done", "is_op_ok", "is_act_ok", "is_issued"]
print("chunk " + " ".join("%13s" % n for n in names) + "   (ns since the first TMA issue; CTA 0, 2 tiles x 16 chunks)")
for g in range(32):
    print("%5d " % g + " ".join("%13d" % (t[e, g] - t0) for e in range(7)))
if os.environ.get("PSI_LBS_GEMM", "").startswith("b"):
    print("bf16x3 path: 16 stages of 64 k (2 tiles): period", np.diff(t[6, 1:16]).mean(), "ns; TMA issue -> landed", (t[4, 1:16] - t[0, 1:16]).mean(),
          "ns; landed -> MMAs issued", (t[6, 1:16] - t[4, 1:16]).mean(), "ns")
    sys.exit(0)
d = lambda a, b: np.diff(t[a, 2:30]).mean() if a == b else (t[a, 2:30] - t[b, 2:30]).mean()
print("mean period per chunk (ns):", d(6, 6), " raw landed after TMA issue:", d(1, 0), " split (op free -> done):", d(3, 2),
      " issuer waits for op after split done:", d(4, 3), " issue time:", d(6, 5), " worker waits for op slot after raw:", d(2, 1))

"""Import shims so UNCHANGED reference scripts resolve the hot-path modules to this package.

    import psi_release_b200.shims as shims; shims.install()
    import smplx                                   # -> psi_release_b200.body_model
    import chamfer_pytorch.dist_chamfer as ext     # -> psi_release_b200.chamfer
    import chamfer_pytorch.dist_chamfer_idx        # idx variant

(fitting_habitat.py:32-34 imports exactly these.)  `torch.nn.functional.grid_sample` is NOT
patched globally; reference code that should use the SDF kernel calls
psi_release_b200.sdf.grid_sample_sdf with the same arguments.
"""
from __future__ import annotations

import sys
import types


def install() -> None:
    from . import body_model, chamfer

    sys.modules.setdefault("smplx", body_model)

    pkg = types.ModuleType("chamfer_pytorch")
    pkg.__path__ = []
    dist = types.ModuleType("chamfer_pytorch.dist_chamfer")
    dist.chamferFunction = chamfer.chamferFunction
    dist.chamferDist = chamfer.chamferDist
    dist_idx = types.ModuleType("chamfer_pytorch.dist_chamfer_idx")
    dist_idx.chamferFunction = chamfer.chamferFunctionIdx

    class _ChamferDistIdx(chamfer.chamferDist):
        def __init__(self):
            super().__init__(return_idx=True)

    dist_idx.chamferDist = _ChamferDistIdx
    pkg.dist_chamfer = dist
    pkg.dist_chamfer_idx = dist_idx
    sys.modules.setdefault("chamfer_pytorch", pkg)
    sys.modules.setdefault("chamfer_pytorch.dist_chamfer", dist)
    sys.modules.setdefault("chamfer_pytorch.dist_chamfer_idx", dist_idx)

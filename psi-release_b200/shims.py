"""Import shims so UNCHANGED reference scripts resolve the hot-path modules to this package.

    import psi_release_b200.shims as shims; shims.install()
    import smplx                                   # -> psi_release_b200.body_model
    import chamfer_pytorch.dist_chamfer as ext     # -> psi_release_b200.chamfer
    import chamfer_pytorch.dist_chamfer_idx        # idx variant
    from human_body_prior.tools.model_loader import load_vposer   # -> the decode half, from a checkpoint dir
    F.grid_sample(sdf.unsqueeze(1), grid, padding_mode='border')  # -> the psi SDF kernel (see below)

(fitting_habitat.py:32-34 imports exactly these.)

`F.grid_sample`: the reference calls it WITHOUT `align_corners` (fitting_habitat.py:150-152) and pins
torch 1.2, where that meant align_corners=True; today's torch silently samples with
align_corners=False -- a different function.  `install()` therefore also routes the
reference's call shape -- a 5-D single-channel float32 CUDA volume sampled at a [B,V,1,1,3] grid with
bilinear mode, border padding and `align_corners` left unset (or True) -- to `sdf.grid_sample_sdf`;
every other call reaches torch's own implementation untouched.  `install(route_grid_sample=False)`
leaves torch alone; `routed_grid_sample()` is the same routing as a context manager.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch
import torch.nn.functional as F

_torch_grid_sample = None      # torch's own function while the patch is active


def _is_psi_sdf_call(input, grid, mode, padding_mode, align_corners) -> bool:
    return (torch.is_tensor(input) and torch.is_tensor(grid) and input.dim() == 5 and grid.dim() == 5
            and input.shape[1] == 1 and input.shape[2] == input.shape[3] == input.shape[4]
            and grid.shape[2] == 1 and grid.shape[3] == 1 and grid.shape[4] == 3 and grid.shape[0] == input.shape[0]
            and mode == "bilinear" and padding_mode == "border" and align_corners in (None, True)
            and input.is_cuda and grid.is_cuda and input.dtype == torch.float32 and grid.dtype == torch.float32
            and not input.requires_grad)


def _grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
    if _is_psi_sdf_call(input, grid, mode, padding_mode, align_corners):
        from . import sdf
        return sdf.grid_sample_sdf(input, grid, padding_mode=padding_mode)
    fn = _torch_grid_sample if _torch_grid_sample is not None else F.grid_sample
    return fn(input, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def patch_grid_sample() -> None:
    global _torch_grid_sample
    if _torch_grid_sample is None:
        _torch_grid_sample = F.grid_sample
        F.grid_sample = _grid_sample


def unpatch_grid_sample() -> None:
    global _torch_grid_sample
    if _torch_grid_sample is not None:
        F.grid_sample = _torch_grid_sample
        _torch_grid_sample = None


@contextlib.contextmanager
def routed_grid_sample():
    """`with shims.routed_grid_sample(): ...` -- F.grid_sample routed as described above inside the block."""
    was_patched = _torch_grid_sample is not None
    patch_grid_sample()
    try:
        yield
    finally:
        if not was_patched:
            unpatch_grid_sample()


def load_vposer(expr_dir, vp_model="snapshot"):
    """human_body_prior.tools.model_loader.load_vposer (model_loader.py:43-72): -> (vposer, settings).
    Only `vposer.decode(z, output_type=...)` exists on the returned module -- the half the fitting and
    training loops call (fitting_habitat.py:115-116, train_s2.py:145-147); settings is None."""
    from .geometry import VPoserDecoder
    return VPoserDecoder.from_checkpoint_dir(expr_dir), None


def install(route_grid_sample: bool = True) -> None:
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

    # human_body_prior.tools.model_loader.load_vposer (only if the real package is not importable)
    if "human_body_prior" not in sys.modules:
        hbp = types.ModuleType("human_body_prior")
        hbp.__path__ = []
        tools = types.ModuleType("human_body_prior.tools")
        tools.__path__ = []
        ml = types.ModuleType("human_body_prior.tools.model_loader")
        ml.load_vposer = load_vposer
        hbp.tools = tools
        tools.model_loader = ml
        sys.modules.setdefault("human_body_prior", hbp)
        sys.modules.setdefault("human_body_prior.tools", tools)
        sys.modules.setdefault("human_body_prior.tools.model_loader", ml)

    if route_grid_sample:
        patch_grid_sample()


def uninstall() -> None:
    """Undo install(): drop the alias modules this file created and restore torch's grid_sample."""
    unpatch_grid_sample()
    from . import body_model
    if sys.modules.get("smplx") is body_model:
        del sys.modules["smplx"]
    for name in ("chamfer_pytorch.dist_chamfer_idx", "chamfer_pytorch.dist_chamfer", "chamfer_pytorch",
                 "human_body_prior.tools.model_loader", "human_body_prior.tools", "human_body_prior"):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__file__", None) is None:
            del sys.modules[name]

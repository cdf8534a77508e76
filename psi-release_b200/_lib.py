"""ctypes binding of lib/libpsi_b200.so (the C ABI in include/psi_b200.h).

There is NO fallback: if the CUDA library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_long
_sz = ctypes.c_size_t
_f = ctypes.c_float


class FitConfig(ctypes.Structure):
    """struct psi_fit_config (include/psi_b200.h)."""
    _fields_ = [("B", _i), ("use_graph", _i), ("w_rec", _f), ("w_vposer", _f), ("w_contact", _f),
                ("w_collision", _f), ("robust_c", _f), ("lr", _f), ("beta1", _f), ("beta2", _f), ("eps", _f), ("nn_mode", _i),
                ("loop_mode", _i), ("loop_unroll", _i), ("loss_mode", _i), ("optimizer", _i), ("lbfgs_lr", _f),
                ("lbfgs_tolerance_grad", _f), ("lbfgs_tolerance_change", _f), ("lbfgs_history", _i), ("lbfgs_zoom_max", _i)]


# psi_fit_trace selectors (PSI_FIT_TRACE_* in include/psi_b200.h): name -> (code, is_int)
FIT_TRACE = {"x_eval": (0, False), "grad_x": (1, False), "verts": (2, False), "sdf": (3, False), "sdf_grad": (4, False),
             "nn_dist": (5, False), "nn_idx": (6, True), "query_ids": (7, True), "losses": (8, False), "x": (9, False),
             "adam_m": (10, False), "adam_v": (11, False), "pose6d": (12, False), "lbfgs_state": (13, True),
             "lbfgs_best": (14, False), "exchange": (15, True)}


class PsiError(RuntimeError):
    pass


def lib_path() -> str:
    """lib/libpsi_b200.so; PSI_B200_LIB names another build of the same sources (A/B runs of compile-time variants)."""
    return os.environ.get("PSI_B200_LIB") or os.path.join(_HERE, "lib", "libpsi_b200.so")


def lib():
    """Load the library once.  Raises if it has not been built (python -m psi_release_b200.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PsiError(
            f"{path} not found: the sm_100a CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "psi_release_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(path)
    sig = {
        "psi_abi_version": (_i, []),
        "psi_error_string": (ctypes.c_char_p, [_i]),
        "psi_launch_count": (ctypes.c_ulonglong, []),
        "psi_nn_workspace_bytes": (_sz, [_i, _i, _i]),
        "psi_nn_fwd": (_i, [_vp, _l, _i, _i, _vp, _l, _i, _vp, _vp, _vp, _sz, _vp]),
        "psi_chamfer_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
        "psi_nn_bwd": (_i, [_vp, _l, _i, _i, _vp, _l, _i, _vp, _vp, _vp, _vp]),
        "psi_chamfer_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "psi_nn_index_create": (_i, [ctypes.POINTER(_vp), _vp, _i, _vp]),
        "psi_nn_index_destroy": (None, [_vp]),
        "psi_nn_index_bytes": (_sz, [_vp]),
        "psi_nn_index_query": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _vp]),
        "psi_nn_index_query_hint": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _vp, _vp]),
        "psi_nn_index_query_mode": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
        "psi_sdf_num_partials": (_i, [_i]),
        "psi_sdf_fwd": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
        "psi_sdf_bwd": (_i, [_vp, _vp, _l, _vp, _vp]),
        "psi_lbs_model_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "psi_lbs_model_destroy": (None, [_vp]),
        "psi_lbs_model_bytes": (_sz, [_vp]),
        "psi_lbs_saved_floats": (_sz, [_vp, _i]),
        "psi_lbs_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _l, _vp, _i, _vp, _vp, _vp, _vp]),
        "psi_lbs_bwd_workspace_bytes": (_sz, [_vp, _i]),
        "psi_lbs_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
        "psi_lbs_bwd2": (_i, [_vp, _i, _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp,
                              _vp, _vp, _vp]),
        "psi_fit_create": (_i, [ctypes.POINTER(_vp), _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i,
                                ctypes.POINTER(FitConfig), _vp]),
        "psi_fit_destroy": (None, [_vp]),
        "psi_fit_run": (_i, [_vp, _vp, _vp, _l, _i, _vp, _vp, _vp]),
        "psi_fit_begin": (_i, [_vp, _vp, _vp, _l, _i, _vp]),
        "psi_fit_end": (_i, [_vp, _vp, _vp, _vp]),
        "psi_fit_launches_per_iteration": (_i, []),
        "psi_fit_profile": (_i, [_vp, _vp, _vp, _l, _i, _i, _vp, _vp, _i, _vp]),
        "psi_fit_trace_bytes": (_sz, [_vp, _i]),
        "psi_fit_trace": (_i, [_vp, _i, _vp, _sz, _vp]),
        "psi_fit_exchange_handle": (_i, [_vp, _vp]),
        "psi_fit_exchange_ptr": (_vp, [_vp]),
        "psi_fit_set_peers": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if L.psi_abi_version() != 3:
        raise PsiError(f"libpsi_b200 ABI {L.psi_abi_version()} != 3 (rebuild: python -m psi_release_b200.build)")
    _LIB = L
    return L


EXPORTS = ["psi_abi_version", "psi_error_string", "psi_launch_count", "psi_nn_workspace_bytes", "psi_nn_fwd",
           "psi_chamfer_fwd", "psi_nn_bwd", "psi_chamfer_bwd", "psi_nn_index_create", "psi_nn_index_destroy",
           "psi_nn_index_bytes", "psi_nn_index_query", "psi_nn_index_query_hint", "psi_nn_index_query_mode", "psi_sdf_num_partials", "psi_sdf_fwd",
           "psi_sdf_bwd", "psi_lbs_model_create", "psi_lbs_model_destroy", "psi_lbs_model_bytes",
           "psi_lbs_saved_floats", "psi_lbs_fwd", "psi_lbs_bwd_workspace_bytes", "psi_lbs_bwd", "psi_lbs_bwd2", "psi_fit_profile",
           "psi_fit_create", "psi_fit_destroy", "psi_fit_run", "psi_fit_begin", "psi_fit_end",
           "psi_fit_launches_per_iteration", "psi_fit_trace_bytes", "psi_fit_trace", "psi_fit_exchange_handle",
           "psi_fit_exchange_ptr", "psi_fit_set_peers"]


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().psi_error_string(rc)
        raise PsiError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) address of a tensor / None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PsiError("psi_release_b200 ops take CUDA tensors only (no CPU fallback); "
                           f"got a tensor on {t.device}")

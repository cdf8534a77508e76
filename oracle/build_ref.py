"""Compile the REFERENCE's own Chamfer extension, unmodified, into oracle/_ref/ (git-ignored,
travels to the GPU box with the snapshot).  Sources are read where they lie under
/root/reference/chamfer_pytorch (chamfer.cu, chamfer_cuda.cpp); nothing is copied.

The result is a torch extension module `chamfer_ref` exposing forward/backward exactly as the
reference's setup.py would (chamfer_pytorch/setup.py:7-10), built for sm_100a only.  It is
used by the GPU tests as the bit-exact pin of the oracle and of our kernel, and by bench.py
as the "reference kernel on the same B200" comparison.  Test infrastructure, not product.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/chamfer_pytorch"
OUT = os.path.join(HERE, "_ref")


def build(verbose=False):
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    mod = load(name="chamfer_ref", sources=[os.path.join(REF, "chamfer_cuda.cpp"), os.path.join(REF, "chamfer.cu")],
               build_directory=OUT, extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"],
               verbose=verbose, is_python_module=False)
    return os.path.join(OUT, "chamfer_ref.so")


def load_ref():
    """Import the prebuilt module (GPU box: /root/reference is absent, only the .so exists)."""
    so = os.path.join(OUT, "chamfer_ref.so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("chamfer_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))

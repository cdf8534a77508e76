"""Build lib/libpsi_b200.so (sm_100a only) with nvcc.  No torch headers: the library is a
plain C-ABI shared object (include/psi_b200.h)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "chamfer.cu", "nn_index.cu", "sdf.cu", "lbs.cu", "fit.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def lib_path() -> str:
    return os.path.join(HERE, "lib", "libpsi_b200.so")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    out = lib_path()
    srcs = [os.path.join(HERE, "csrc", s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "psi_b200.h"))
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("PSI_EXTRA_NVCC_FLAGS", "").split()
    cmd = [_nvcc(), *flags, "-shared", "-o", out, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

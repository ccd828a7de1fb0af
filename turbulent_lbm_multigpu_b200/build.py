"""Build turbulent_lbm_multigpu_b200/lib/liblbm_b200.so with nvcc for sm_100a (in-tree)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblbm_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    # parity build: no FMA contraction, IEEE division/sqrt, denormals kept -- the fp32/fp64
    # results are bit-identical to the reference kernels evaluated in source order
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-cudart", "static",
]


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, "lbm_capi.cu")]


def deps():
    return sources() + [os.path.join(CSRC, "lbm_kernels.cuh"), os.path.join(ROOT, "include", "lbm_b200.h")]


def build(force=False, verbose=False, extra=(), out=None):
    """extra/out: tuning variants (e.g. -DLBM_LB_BETA=__launch_bounds__(256,2)) under another name."""
    os.makedirs(LIBDIR, exist_ok=True)
    lib = out or LIB
    if (not force and os.path.exists(lib)
            and os.path.getmtime(lib) >= max(os.path.getmtime(d) for d in deps())):
        return lib
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-o", lib] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Build the C++ host facade check binary `lbm_b200` (host/main.cpp + the header-only mirror of
the reference's classes) against liblbm_b200.so.  In-tree, g++ only; strict IEEE host
arithmetic (-ffp-contract=off) so that CLbmSkeleton's parametrisation is bit-identical."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
BIN = os.path.join(HERE, "lbm_b200")
LIBDIR = os.path.join(PKG, "lib")


def deps():
    return (glob.glob(os.path.join(HERE, "*.hpp")) + glob.glob(os.path.join(HERE, "*.h"))
            + [os.path.join(HERE, "main.cpp"), os.path.join(ROOT, "include", "lbm_b200.h"),
               os.path.join(LIBDIR, "liblbm_b200.so")])


def build(force=False):
    if (not force and os.path.exists(BIN)
            and os.path.getmtime(BIN) >= max(os.path.getmtime(d) for d in deps() if os.path.exists(d))):
        return BIN
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread",
           os.path.join(HERE, "main.cpp"), "-o", BIN, "-L" + LIBDIR, "-llbm_b200",
           "-Wl,-rpath,$ORIGIN/../lib", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

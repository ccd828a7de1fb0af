#!/usr/bin/env python3
"""Build oracle/liblbm_oracle.so (the plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liblbm_oracle.so")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("lbm_oracle.c", "lbm_oracle_impl.h")]
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB
    cmd = ["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
           "-shared", "-Wall", "-o", LIB, srcs[0], "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

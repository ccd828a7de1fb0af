#!/usr/bin/env python3
"""Build tuning variants of liblbm_b200.so next to the default one: liblbm_b200_<tag>.so.

Tags: lb<T>x<B> = __launch_bounds__(T, B) on the beta kernel; ld<a>st<b> = cache operators of the
aligned slot streams (LBM_HINT_LD / LBM_HINT_ST in csrc/lbm_kernels.cuh); offtab<n> = beta addresses
from the constant-bank offset table (LBM_BETA_OFFTAB=n); tags combine with '_' (offtab1_lb128x6).  Load one with
LBM_B200_LIB=<path> (tools/sweep_variants.sh)."""
import os
import re
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulent_lbm_multigpu_b200 import build as b  # noqa: E402

DEFAULT = ["lb256x2", "lb128x5", "lb128x6", "lb128x8"]


def defines(tag):
    out = []
    for part in tag.split("_"):
        m = re.fullmatch(r"lb(\d+)x(\d+)", part)
        if m:
            out += ["-DLBM_LB_MAXT=%s" % m.group(1), "-DLBM_LB_MINB=%s" % m.group(2)]
            continue
        m = re.fullmatch(r"ld(\d)st(\d)", part)
        if m:
            out += ["-DLBM_HINT_LD=%s" % m.group(1), "-DLBM_HINT_ST=%s" % m.group(2)]
            continue
        m = re.fullmatch(r"offtab(\d)", part)
        if m:
            out += ["-DLBM_BETA_OFFTAB=%s" % m.group(1)]
            continue
        raise SystemExit("unknown variant tag %r" % part)
    return out


def one(tag):
    out = os.path.join(b.LIBDIR, "liblbm_b200_%s.so" % tag)
    b.build(force=True, extra=defines(tag), out=out)
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for o in ex.map(one, sys.argv[1:] or DEFAULT):
            print(o)

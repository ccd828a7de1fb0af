#!/usr/bin/env python3
"""Build tuning variants of liblbm_b200.so (different __launch_bounds__) next to the default one."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulent_lbm_multigpu_b200 import build as b  # noqa: E402

VARIANTS = {"lb256x2": (256, 2), "lb128x5": (128, 5), "lb128x6": (128, 6), "lb128x8": (128, 8)}


def one(item):
    tag, lb = item
    out = os.path.join(b.LIBDIR, "liblbm_b200_%s.so" % tag)
    b.build(force=True, extra=["-DLBM_LB_MAXT=%d" % lb[0], "-DLBM_LB_MINB=%d" % lb[1]], out=out)
    return out


if __name__ == "__main__":
    sel = {k: v for k, v in VARIANTS.items() if not sys.argv[1:] or k in sys.argv[1:]}
    with ThreadPoolExecutor(4) as ex:
        for o in ex.map(one, sel.items()):
            print(o)

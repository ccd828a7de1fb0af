#!/usr/bin/env python3
"""Per-kernel device timeline of a few overlapped steps of the single-GPU loopback proxy
(tools/probe_overlap.py): which kernel runs when, on which step, and the gaps between them.
Uses the library's own event timeline (lbmProfile*).  Tuning aid."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probe_overlap import loopback_solver  # noqa: E402


def main():
    axes = sys.argv[1] if len(sys.argv) > 1 else "x"
    size = tuple(int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "256x256x256").split("x"))
    s = loopback_solver(size, sorted("xyz".index(c) for c in axes), np.float32, 0.1)
    for _ in range(10):
        s.commStep()
    s.wait()
    s.profileEnable(1)
    for _ in range(6):
        s.commStep()
    s.wait()
    ev = s.profileEvents()
    t0 = ev[0][1]
    for name, a, b in ev:
        print("%-22s start %9.1f us  dur %8.1f us" % (name, (a - t0) / 1e3, (b - a) / 1e3))
    s.close()


if __name__ == "__main__":
    main()

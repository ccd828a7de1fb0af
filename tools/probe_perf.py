#!/usr/bin/env python3
"""Sweep launch configurations of the step kernels on one GPU and print MLUPS.

Usage: python tools/probe_perf.py [--size 256] [--dtype f32] [--steps 40]
Not a benchmark of record (bench.py is); a tuning aid to pick vector width / block size.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulent_lbm_multigpu_b200.domain import CDomain  # noqa: E402
from turbulent_lbm_multigpu_b200.solver import CLbmSolver  # noqa: E402
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters  # noqa: E402


def run(size, dtype, vw, block, cs, steps, store=False, mode="both"):
    p = compute_parameters(size, (0.1, 0.1, 0.1), dtype=dtype)
    s = CLbmSolver(0, 0, [[1, 1]] * 3, CDomain(0, size, (0, 0, 0), (0.1,) * 3), dtype=dtype,
                   store_velocity=store, store_density=store, smagorinsky_cs=cs,
                   vector_width=vw, block_size=block, params=p)
    rect = (size[0] - 2, 1, size[2] - 2)
    s.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, size[1] - 2, 1), rect)
    fn = {"both": s.simulationStep, "alpha": s.simulationStepAlpha, "beta": s.simulationStepBeta}[mode]
    for _ in range(6):
        fn()
    s.wait()
    s.timerStart()
    for _ in range(steps):
        fn()
    ms = s.timerStop()
    n = size[0] * size[1] * size[2]
    mlups = n * steps / (ms * 1e-3) / 1e6
    s.close()
    return mlups, ms / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs="+", default=[256])
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--cs", type=float, nargs="+", default=[0.0, 0.1])
    ap.add_argument("--vw", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--block", type=int, nargs="+", default=[64, 128, 256])
    ap.add_argument("--modes", nargs="+", default=["alpha", "beta", "both"])
    a = ap.parse_args()
    dtype = np.float32 if a.dtype == "f32" else np.float64
    bpc = 2 * 19 * np.dtype(dtype).itemsize + 4
    sz = a.size if len(a.size) == 3 else [a.size[0]] * 3
    for cs in a.cs:
        for vw in a.vw:
            if dtype == np.float64 and vw > 2:
                continue
            for block in a.block:
                row = {}
                for mode in a.modes:
                    mlups, ms = run(tuple(sz), dtype, vw, block, cs, a.steps, mode=mode)
                    row[mode] = round(mlups)
                    row[mode + "_GBs"] = round(mlups * bpc / 1e3)
                print(json.dumps(dict(size=sz, dtype=a.dtype, cs=cs, vw=vw, block=block, **row)), flush=True)


if __name__ == "__main__":
    main()

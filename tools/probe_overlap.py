#!/usr/bin/env python3
"""Single-GPU proxy for the cost of the overlapped multi-GPU step.

One sub-domain on ONE GPU whose halo faces are wired to ITSELF (the x+ face pushes into the
receive block of the x- face and so on: a periodic box).  Every kernel an interior rank of a
decomposed run launches -- step kernels with the fused halo push, flag waits, unpack -- runs with
the same sizes and dependencies, only the peer stores land in local memory instead of crossing
NVLink.  What it isolates is the on-GPU price of overlapping (extra launches, boundary-first
scheduling, fences, spinning waits) against the plain single-domain step: A/B tuning at 1x GPU
cost, and something `ncu` can profile (ncu must not be used on multi-rank runs).

    python tools/probe_overlap.py [--axes z|x|xyz|...] [--size 256] [--dtype f32] [--steps 200]

Tuning aid, not a benchmark of record (the physics is a periodic box, not the cavity).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulent_lbm_multigpu_b200 import capi  # noqa: E402
from turbulent_lbm_multigpu_b200.domain import CComm, CDomain  # noqa: E402
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters  # noqa: E402
from turbulent_lbm_multigpu_b200.solver import CLbmSolver  # noqa: E402


def loopback_solver(size, axes, dtype, cs, axis_order=None, **kw):
    p = compute_parameters(size, (0.1,) * 3, dtype=dtype)
    bc = [[1, 1], [1, 1], [1, 1]]
    for a in axes:
        bc[a] = [8, 8]
    s = CLbmSolver(0, 0, bc, CDomain(0, size, (0, 0, 0), (0.1,) * 3), dtype=dtype, store_velocity=False,
                   store_density=False, smagorinsky_cs=cs, params=p, **kw)
    if axis_order is None:
        axis_order = "zyx" if 0 in axes else "xyz"
    s.commSetAxisOrder(axis_order)
    ids = {}
    for a in axes:
        face = list(size)
        face[a] = 1
        for side in (0, 1):
            so, ro, d = [0, 0, 0], [0, 0, 0], [0, 0, 0]
            if side == 0:
                so[a], ro[a], d[a] = 1, 0, 1
            else:
                so[a], ro[a], d[a] = size[a] - 2, size[a] - 1, -1
            ids[(a, side)] = s.commAddFace(CComm(0, tuple(face), tuple(face), tuple(so), tuple(ro), tuple(d)))
    for a in axes:
        s.commConnectLocal(ids[(a, 0)], s, ids[(a, 1)])
        s.commConnectLocal(ids[(a, 1)], s, ids[(a, 0)])
    return s


def timed(s, fn, steps, warm=10):
    for _ in range(warm):
        fn()
    s.wait()
    s.timerStart()
    for _ in range(steps):
        fn()
    return s.timerStop() / steps * 1e3          # us per step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--axes", default="z")
    ap.add_argument("--size", default="256")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--cs", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--axis-order", default=None)
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--only", default=None, help="plain | overlap: run just that loop (for ncu)")
    args = ap.parse_args()
    dims = [int(v) for v in args.size.split("x")]
    size = tuple(dims * 3) if len(dims) == 1 else tuple(dims)
    dtype = np.float32 if args.dtype == "f32" else np.float64
    axes = sorted("xyz".index(c) for c in args.axes)
    s = loopback_solver(size, axes, dtype, args.cs, args.axis_order)
    out = {"size": list(size), "axes": args.axes, "dtype": args.dtype, "cs": args.cs,
           "axis_order": "zyx" if s.commAxisOrder() == capi.LBM_AXIS_ORDER_ZYX else "xyz"}
    cells = size[0] * size[1] * size[2]
    bpl = (2 * 19 * np.dtype(dtype).itemsize + 4)
    if args.only in (None, "plain"):
        out["plain_us"] = round(timed(s, s.simulationStep, args.steps), 2)
        out["plain_GBs"] = round(bpl * cells / out["plain_us"] / 1e3, 1)
    if args.only in (None, "overlap"):
        out["overlap_us"] = round(timed(s, s.commStep, args.steps), 2)
        out["overlap_GBs"] = round(bpl * cells / out["overlap_us"] / 1e3, 1)
    if args.only is None:
        out["overlap_cost_pct"] = round(100.0 * (out["overlap_us"] / out["plain_us"] - 1.0), 2)
    if args.timeline:
        acc = np.zeros((2, 4))
        for i in range(20):
            acc[i & 1] += np.array(s.commStepTimed())
        acc /= 10.0
        c0 = s.simulation_step_counter & 1
        out["timeline_ms"] = {"beta": [round(float(v), 4) for v in acc[c0]], "alpha": [round(float(v), 4) for v in acc[1 - c0]],
                              "legend": "ms after the fork: boundary done / step kernel done / exchange done / join"}
    s.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

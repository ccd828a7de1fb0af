#!/usr/bin/env python3
"""Cost of splitting the step into shell + interior (what the overlapped multi-GPU step does),
measured on ONE GPU with ghost faces in z and no exchange.  Tuning aid, not a benchmark of record."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulent_lbm_multigpu_b200.domain import CDomain  # noqa: E402
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters  # noqa: E402
from turbulent_lbm_multigpu_b200.solver import CLbmSolver  # noqa: E402


def main():
    size = (256, 256, 256)
    steps = 100
    p = compute_parameters(size, (0.1,) * 3, dtype=np.float32)
    axis = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    faces = 0b11 << (2 * axis)           # ghost faces on both sides of that axis
    bc = [[1, 1], [1, 1], [1, 1]]
    bc[axis] = [8, 8]
    out = {"axis": axis}
    for cs in (0.1,):
        s = CLbmSolver(0, 0, bc, CDomain(0, size, (0, 0, 0), (0.1,) * 3), dtype=np.float32,
                       store_velocity=False, store_density=False, smagorinsky_cs=cs, params=p)

        def unsplit():
            s.simulationStep()

        def split_concurrent():
            s.commWaitCompute(); s.stepShellComm(faces); s.stepInterior(faces); s.computeWaitComm()

        def split_serial():
            s.stepShell(faces); s.stepInterior(faces)

        def interior_only():
            s.stepInterior(faces)

        def shell_only():
            s.stepShell(faces)

        for name, fn in (("unsplit", unsplit), ("split_concurrent", split_concurrent), ("split_serial", split_serial),
                         ("interior_only", interior_only), ("shell_only", shell_only)):
            for _ in range(6):
                fn()
            s.wait()
            s.timerStart()
            for _ in range(steps):
                fn()
            out["%s cs=%g" % (name, cs)] = round(s.timerStop() / steps * 1e3, 1)     # us per step
        s.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

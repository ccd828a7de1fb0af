#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own kernels (oracle/_ref).

Run HERE (the container with /root/reference mounted): `python tests/golden/make_golden.py`.
Each fixture holds the complete state (dd, velocity, density, flags) after `steps` steps of the
lid-driven cavity scenario (conf.xml physics, L = 0.1) computed by the reference's lbm_init.cl,
lbm_alpha.cl and lbm_beta.cl compiled as strict-IEEE C++ -- `variant` 0 is lbm_beta.cl exactly
as shipped (shared-memory path executed with real work-group semantics), 1 its own
USE_SHARED_MEMORY 0 path.  The reference ships no golden data of its own.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters  # noqa: E402

CASES = [
    # name, size, dtype, bc, variant, steps
    ("f32_16x8x8_walls_shm", (16, 8, 8), np.float32, (1,) * 6, 0, (1, 2, 7, 8)),
    ("f32_16x8x8_walls_noshm", (16, 8, 8), np.float32, (1,) * 6, 1, (1, 2, 7, 8)),
    ("f32_24x8x8_walls_shm", (24, 8, 8), np.float32, (1,) * 6, 0, (1, 2, 7, 8)),
    ("f32_24x8x8_ghostmix_shm", (24, 8, 8), np.float32, (8, 1, 1, 8, 8, 1), 0, (1, 2, 5, 6)),
    ("f32_16x8x8_ghosts_shm", (16, 8, 8), np.float32, (8,) * 6, 0, (1, 2, 5, 6)),
    ("f64_16x8x8_walls_shm", (16, 8, 8), np.float64, (1,) * 6, 0, (1, 2, 7, 8)),
    ("f64_24x8x8_walls_noshm", (24, 8, 8), np.float64, (1,) * 6, 1, (1, 2, 7, 8)),
]

#: scalar known answers at 64^3 (the values of SURVEY.md A.2 / BASELINE.md §4)
KAT = [
    ("f32_64_shm", (64, 64, 64), np.float32, 0, (100, 101)),
    ("f32_64_noshm", (64, 64, 64), np.float32, 1, (100, 101)),
]


def scenario(size, dtype, bc, variant):
    p = compute_parameters(size, (0.1, 0.1, 0.1), dtype=dtype)
    s = ref.RefSolver(size, list(bc), p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=variant)
    rect = (size[0] - 2, 1, size[2] - 2)
    s.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, size[1] - 2, 1), rect)
    return s, p


def main():
    for name, size, dtype, bc, variant, steps in CASES:
        s, p = scenario(size, dtype, bc, variant)
        out = dict(size=np.array(size), bc=np.array(bc), variant=variant, steps=np.array(steps),
                   inv_tau=p.inv_tau, u_lid=p.u_lid, gravitation=np.array(p.gravitation))
        done = 0
        for k in steps:
            while done < k:
                s.simulationStep()
                done += 1
            out["dd_%d" % k] = s.dd.copy()
            out["velocity_%d" % k] = s.velocity.copy()
            out["density_%d" % k] = s.density.copy()
        out["flags"] = s.flags.copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name)
    kat = {}
    for name, size, dtype, variant, steps in KAT:
        s, p = scenario(size, dtype, (1,) * 6, variant)
        n = s.n
        g = size[0] // 2 + size[1] // 2 * size[0] + size[2] // 2 * size[0] * size[1]
        done = 0
        for k in steps:
            while done < k:
                s.simulationStep()
                done += 1
            kat["%s_%d" % (name, k)] = np.array([s.velocity[g], s.velocity[n + g], s.velocity[2 * n + g],
                                                s.density[g], s.getVelocityChecksum()], dtype=np.float64)
    np.savez(os.path.join(HERE, "kat_64.npz"), **kat)
    print("wrote kat_64")


if __name__ == "__main__":
    main()

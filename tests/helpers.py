"""Shared builders for the parity tests: the same scenario on the CUDA solver and on the oracle."""
import numpy as np

from oracle import multi as omulti
from oracle import port
from turbulent_lbm_multigpu_b200 import capi
from turbulent_lbm_multigpu_b200.domain import CDomain
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters

WALLS = (1, 1, 1, 1, 1, 1)


def params_for(size, dtype, length=(0.1, 0.1, 0.1)):
    return compute_parameters(size, length, dtype=dtype)


def bc33(bc6):
    return [[bc6[0], bc6[1]], [bc6[2], bc6[3]], [bc6[4], bc6[5]]]


def make_cuda(size, dtype=np.float32, bc=WALLS, order=0, cs=0.0, store=True, lid=True, vector_width=0,
              block_size=0, wg=128, params=None, device=0):
    from turbulent_lbm_multigpu_b200.solver import CLbmSolver
    p = params or params_for(size, dtype)
    s = CLbmSolver(0, device, bc33(bc), CDomain(0, size, (0, 0, 0), (0.1, 0.1, 0.1)), dtype=dtype,
                   store_velocity=store, store_density=store, smagorinsky_cs=cs, beta_order=order,
                   vector_width=vector_width, block_size=block_size, computation_kernel_count=wg, params=p)
    assert not s.error(), str(s.error)
    if lid:
        set_lid(s, size)
    return s


def make_oracle(size, dtype=np.float32, bc=WALLS, order=0, cs=0.0, lid=True, wg=128, params=None):
    p = params or params_for(size, dtype)
    s = port.OracleSolver(size, list(bc), p.inv_tau, p.gravitation, p.u_lid, dtype=dtype, variant=order,
                          tau=p.tau, smagorinsky_cs=cs, wg=wg)
    if lid:
        set_lid(s, size)
    return s


def set_lid(solver, size):
    rect = (size[0] - 2, 1, size[2] - 2)
    solver.setFlags(np.full(rect[0] * rect[2], 4, np.int32), (1, size[1] - 2, 1), rect)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_state_equal(cuda, oracle, what=("dd", "velocity", "density", "flags"), ctx=""):
    got = {"dd": cuda.storeDensityDistribution, "velocity": cuda.storeVelocity,
           "density": cuda.storeDensity, "flags": cuda.storeFlags}
    exp = {"dd": oracle.dd, "velocity": oracle.velocity, "density": oracle.density, "flags": oracle.flags}
    for k in what:
        g = got[k]()
        e = exp[k]
        if not bits_equal(g, e):
            bad = np.nonzero(g != e)[0]
            raise AssertionError("%s %s: %d of %d elements differ, first at %s: cuda=%r oracle=%r"
                                 % (ctx, k, bad.size, g.size, bad[:5], g[bad[:3]], e[bad[:3]]))

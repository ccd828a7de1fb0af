"""ctypes driver for oracle/liblbm_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY.

``OracleSolver`` has the same CLbmSolver-shaped interface as ``oracle.ref.RefSolver`` so
the two can be run side by side (tests/test_oracle.py) and so either can stand in
for a sub-domain in ``oracle.multi``.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

SHM, NOSHM = 0, 1
_LIB = None

#: D3Q19 vectors in slot order (reference src/main.cpp:38-66)
LBM_UNITS = (
    (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0),
    (1, 1, 0), (-1, -1, 0), (1, -1, 0), (-1, 1, 0),
    (1, 0, 1), (-1, 0, -1), (1, 0, -1), (-1, 0, 1),
    (0, 1, 1), (0, -1, -1), (0, 1, -1), (0, -1, 1),
    (0, 0, 1), (0, 0, -1), (0, 0, 0),
)


def load():
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path):
            _build.build()
        lib = ctypes.CDLL(path)
        vp, i, l = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        for sfx, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(lib, "lbmo_init_" + sfx).argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i]
            getattr(lib, "lbmo_alpha_" + sfx).argtypes = [vp, vp, vp, vp, i, i, i] + [ct] * 7 + [i, i]
            getattr(lib, "lbmo_beta_" + sfx).argtypes = [vp, vp, vp, vp, i, i, i] + [ct] * 7 + [i, i, i, i]
            getattr(lib, "lbmo_copy_rect_" + sfx).argtypes = [vp, l, vp, vp, vp, l, vp, vp, vp]
            getattr(lib, "lbmo_checksum_" + sfx).argtypes = [vp, vp, l]
            getattr(lib, "lbmo_checksum_" + sfx).restype = ctypes.c_float
            getattr(lib, "lbmo_skeleton_" + sfx).argtypes = [i, ct, vp, ct, vp, ct, ct, ct, vp]
            getattr(lib, "lbmo_skeleton_" + sfx).restype = i
        _LIB = lib
    return _LIB


def skeleton(domain_size_x, domain_x_length, gravitation=(0.0, -9.81, 0.0), viscosity=0.001308,
             cavity_velocity=(100.0, 0.0, 0.0, 1.0), dtype=np.float32):
    """C restatement of CLbmSkeleton::init; returns dict of T scalars and the error flag."""
    lib = load()
    dt = np.dtype(dtype)
    sfx = "f32" if dt == np.float32 else "f64"
    g = np.asarray(gravitation, dt)
    c = np.asarray(cavity_velocity, dt)
    out = np.zeros(13, dt)
    T = dt.type
    err = getattr(lib, "lbmo_skeleton_" + sfx)(int(domain_size_x), T(domain_x_length), g.ctypes.data,
                                               T(viscosity), c.ctypes.data, T(1.0), T(0.0001),
                                               T(0.953575), out.ctypes.data)
    keys = ("d_cell_length", "d_timestep", "tau", "inv_tau", "inv_trt_tau")
    res = {k: out[n] for n, k in enumerate(keys)}
    res["gravitation"] = tuple(out[5:8])
    res["drivenCavityVelocity"] = tuple(out[8:12])
    res["d_reynolds"] = out[12]
    res["error"] = bool(err)
    return res


def smagorinsky_k(cs, dtype):
    """smag_k = 18*sqrt(2)*C_s^2 evaluated in double, rounded once to T (same as the product)."""
    return np.dtype(dtype).type(18.0 * np.sqrt(2.0) * float(cs) * float(cs))


class OracleSolver:
    """One sub-domain advanced by the C restatement; CLbmSolver-shaped interface."""

    def __init__(self, size, bc, inv_tau, gravitation, u_lid, dtype=np.float32, wg=128,
                 variant=SHM, tau=None, smagorinsky_cs=0.0, store_velocity=True, store_density=True):
        self.lib = load()
        self.dtype = np.dtype(dtype)
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        T = self.dtype.type
        self.size = tuple(int(s) for s in size)
        self.n = self.size[0] * self.size[1] * self.size[2]
        self.variant = variant
        # the shared-memory path's work-group x-shift only exists on that path
        self.wg = int(wg) if variant == SHM else 0
        self.inv_tau = T(inv_tau)
        self.tau = T(tau) if tau is not None else T(T(1.0) / T(inv_tau))
        self.smag_k = smagorinsky_k(smagorinsky_cs, self.dtype)
        self.g = [T(x) for x in gravitation]
        self.u_lid = T(u_lid)
        self.bc = np.asarray(bc, dtype=np.int32).reshape(6).copy()
        self.store_velocity = bool(store_velocity)
        self.store_density = bool(store_density)
        self.dd = np.zeros(19 * self.n, self.dtype)
        self.flags = np.zeros(self.n, np.int32)
        self.velocity = np.zeros(3 * self.n, self.dtype)
        self.density = np.zeros(self.n, self.dtype)
        self.reset()

    def reset(self):
        self.simulation_step_counter = 0
        getattr(self.lib, "lbmo_init_" + self.sfx)(
            self.dd.ctypes.data, self.flags.ctypes.data, self.velocity.ctypes.data,
            self.density.ctypes.data, self.bc.ctypes.data, *self.size,
            int(self.store_velocity), int(self.store_density))

    def _args(self):
        return (self.dd.ctypes.data, self.flags.ctypes.data, self.velocity.ctypes.data,
                self.density.ctypes.data, *self.size, self.inv_tau, self.g[0], self.g[1], self.g[2],
                self.u_lid, self.tau, self.smag_k, int(self.store_velocity), int(self.store_density))

    def simulationStepAlpha(self):
        getattr(self.lib, "lbmo_alpha_" + self.sfx)(*self._args())

    def simulationStepBeta(self):
        getattr(self.lib, "lbmo_beta_" + self.sfx)(*self._args(), int(self.variant), int(self.wg))

    def simulationStep(self):
        """CLbmSolver::simulationStep, src/CLbmSolver.hpp:664-676: counter&1 ? alpha : beta."""
        if self.simulation_step_counter & 1:
            self.simulationStepAlpha()
        else:
            self.simulationStepBeta()
        self.simulation_step_counter += 1

    # rect access: arrays viewed as [comp][z][y][x]
    def _view(self, arr, comps):
        sx, sy, sz = self.size
        return arr.reshape(comps, sz, sy, sx)

    @staticmethod
    def _sl(origin, size):
        return (slice(None), slice(origin[2], origin[2] + size[2]),
                slice(origin[1], origin[1] + size[1]), slice(origin[0], origin[0] + size[0]))

    def storeDensityDistribution(self, origin=None, size=None):
        if origin is None:
            return self.dd.copy()
        return self._view(self.dd, 19)[self._sl(origin, size)].reshape(-1).copy()

    def setDensityDistribution(self, src, origin, size, norm=None):
        """CLbmSolver::setDensityDistribution (:719-757); with norm only slots with norm.e_f > 0."""
        src = np.asarray(src, self.dtype).reshape(19, size[2], size[1], size[0])
        v = self._view(self.dd, 19)
        sl = self._sl(origin, size)
        for f in range(19):
            if norm is not None and sum(int(a) * int(b) for a, b in zip(norm, LBM_UNITS[f])) <= 0:
                continue
            v[(f,) + sl[1:]] = src[f]

    def storeVelocity(self, origin=None, size=None):
        if origin is None:
            return self.velocity.copy()
        return self._view(self.velocity, 3)[self._sl(origin, size)].reshape(-1).copy()

    def storeDensity(self, origin=None, size=None):
        if origin is None:
            return self.density.copy()
        return self._view(self.density, 1)[self._sl(origin, size)].reshape(-1).copy()

    def storeFlags(self, origin=None, size=None):
        if origin is None:
            return self.flags.copy()
        return self._view(self.flags, 1)[self._sl(origin, size)].reshape(-1).copy()

    def setFlags(self, src, origin, size):
        src = np.asarray(src, np.int32).reshape(1, size[2], size[1], size[0])
        self._view(self.flags, 1)[self._sl(origin, size)] = src

    def copy_rect(self, src, src_off, so, ss, dst, dst_off, do, ds, block):
        """C restatement of copy_buffer_rect.cl (used to cross-check the slicing above)."""
        i3 = lambda v: (ctypes.c_int * 3)(*[int(x) for x in v])
        getattr(self.lib, "lbmo_copy_rect_" + self.sfx)(
            src.ctypes.data, int(src_off), i3(so), i3(ss), dst.ctypes.data, int(dst_off),
            i3(do), i3(ds), i3(block))

    def getVelocityChecksum(self):
        return np.float32(getattr(self.lib, "lbmo_checksum_" + self.sfx)(
            self.velocity.ctypes.data, self.flags.ctypes.data, self.n))

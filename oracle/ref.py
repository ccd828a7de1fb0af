"""ctypes driver for oracle/_ref/libref.so -- TEST INFRASTRUCTURE ONLY.

``RefSolver`` re-enacts the host side of the reference's ``CLbmSolver`` around the
reference's *own kernels* (compiled from /root/reference by oracle/build_ref.py):
buffer set-up and init launch (src/CLbmSolver.hpp:381-403,619-635), the alpha/beta parity
of ``simulationStep`` (:664-676), and the rect get/set entry points that are implemented
with 19/3/1 launches of copy_buffer_rect (:695-757, :764-978).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

SHM, NOSHM = 0, 1


def lib_path(fast=False):
    return os.path.join(_HERE, "_ref", "libref_fast.so" if fast else "libref.so")


def available(fast=False):
    return os.path.exists(lib_path(fast))


def load(fast=False):
    if fast not in _LIBS:
        lib = ctypes.CDLL(lib_path(fast))
        vp, ip, d, i = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_double, ctypes.c_int
        lib.ref_count.restype = i
        lib.ref_describe.argtypes = [i, ip]
        lib.ref_lookup.argtypes = [i, i, i, i, i]
        lib.ref_lookup.restype = i
        lib.ref_set_threads.argtypes = [i]
        lib.ref_get_max_threads.restype = i
        lib.ref_init.argtypes = [i, vp, vp, vp, vp, vp, d]
        lib.ref_alpha.argtypes = [i, vp, vp, vp, vp, d, d, d, d, d]
        lib.ref_beta.argtypes = [i, i, vp, vp, vp, vp, d, d, d, d, d]
        lib.ref_copy_rect.argtypes = [i, vp, i, ip, ip, vp, i, ip, ip, ip]
        _LIBS[fast] = lib
    return _LIBS[fast]


def instances(fast=False):
    lib = load(fast)
    out = []
    for k in range(lib.ref_count()):
        a = (ctypes.c_int * 5)()
        lib.ref_describe(k, a)
        out.append(tuple(a))
    return out


def _i3(v):
    return (ctypes.c_int * 3)(*[int(x) for x in v])


class RefSolver:
    """The reference kernels behind a CLbmSolver-shaped interface (one sub-domain)."""

    def __init__(self, size, bc, inv_tau, gravitation, u_lid, dtype=np.float32, wg=128,
                 variant=SHM, fast=False):
        self.lib = load(fast)
        self.dtype = np.dtype(dtype)
        self.size = tuple(int(s) for s in size)
        self.n = self.size[0] * self.size[1] * self.size[2]
        self.idx = self.lib.ref_lookup(self.dtype.itemsize, *self.size, wg)
        if self.idx < 0:
            raise KeyError("no reference kernel instance compiled for %s %s wg=%d "
                           "(add it to SIZES in oracle/build_ref.py)" % (self.dtype, self.size, wg))
        self.variant = variant
        self.inv_tau = float(inv_tau)
        self.g = [float(x) for x in gravitation]
        self.u_lid = float(u_lid)
        self.bc = np.asarray(bc, dtype=np.int32).reshape(6).copy()
        self.dd = np.zeros(19 * self.n, self.dtype)
        self.flags = np.zeros(self.n, np.int32)
        self.velocity = np.zeros(3 * self.n, self.dtype)
        self.density = np.zeros(self.n, self.dtype)
        self.reset()

    # --- CLbmSolver::reset / simulationStep -------------------------------------------
    def reset(self):
        self.simulation_step_counter = 0
        self.lib.ref_init(self.idx, self.dd.ctypes.data, self.flags.ctypes.data,
                          self.velocity.ctypes.data, self.density.ctypes.data,
                          self.bc.ctypes.data, self.u_lid)

    def _args(self):
        return (self.dd.ctypes.data, self.flags.ctypes.data, self.velocity.ctypes.data,
                self.density.ctypes.data, self.inv_tau, self.g[0], self.g[1], self.g[2], self.u_lid)

    def simulationStepAlpha(self):
        self.lib.ref_alpha(self.idx, *self._args())

    def simulationStepBeta(self):
        self.lib.ref_beta(self.idx, self.variant, *self._args())

    def simulationStep(self):
        if self.simulation_step_counter & 1:
            self.simulationStepAlpha()
        else:
            self.simulationStepBeta()
        self.simulation_step_counter += 1

    # --- rect get/set through copy_buffer_rect ----------------------------------------
    def _copy(self, src, src_off, so, ss, dst, dst_off, do, ds, block):
        nbytes = src.dtype.itemsize
        self.lib.ref_copy_rect(nbytes, src.ctypes.data, int(src_off), _i3(so), _i3(ss),
                               dst.ctypes.data, int(dst_off), _i3(do), _i3(ds), _i3(block))

    def _store(self, arr, comps, origin, size):
        cells = int(np.prod(size))
        out = np.zeros(comps * cells, arr.dtype)
        for f in range(comps):
            self._copy(arr, f * self.n, origin, self.size, out, f * cells, (0, 0, 0), size, size)
        return out

    def _set(self, arr, comps, src, origin, size, keep=None):
        cells = int(np.prod(size))
        src = np.ascontiguousarray(src, dtype=arr.dtype).reshape(-1)
        for f in range(comps):
            if keep is not None and not keep(f):
                continue
            self._copy(src, f * cells, (0, 0, 0), size, arr, f * self.n, origin, self.size, size)

    def storeDensityDistribution(self, origin=None, size=None):
        if origin is None:
            return self.dd.copy()
        return self._store(self.dd, 19, origin, size)

    def setDensityDistribution(self, src, origin, size, norm=None):
        from .port import LBM_UNITS
        keep = None
        if norm is not None:  # CLbmSolver.hpp:748: norm.dotProd(lbm_units[f]) > 0
            keep = lambda f: sum(int(a) * int(b) for a, b in zip(norm, LBM_UNITS[f])) > 0
        self._set(self.dd, 19, src, origin, size, keep)

    def storeVelocity(self, origin=None, size=None):
        return self.velocity.copy() if origin is None else self._store(self.velocity, 3, origin, size)

    def storeDensity(self, origin=None, size=None):
        return self.density.copy() if origin is None else self._store(self.density, 1, origin, size)

    def storeFlags(self, origin=None, size=None):
        if origin is None:
            return self.flags.copy()
        # the reference copies int flags with the kernel typed on T (only right for 4-byte T,
        # SURVEY.md 7.5-5); use the 4-byte instance explicitly
        return self._store(self.flags.view(np.float32), 1, origin, size).view(np.int32)

    def setFlags(self, src, origin, size):
        src = np.ascontiguousarray(src, dtype=np.int32).reshape(-1)
        self._set(self.flags.view(np.float32), 1, src.view(np.float32), origin, size)

    def getVelocityChecksum(self):
        """CLbmSolver::getVelocityChecksum (:1103-1123): serial float accumulation."""
        n = self.n
        v = self.velocity
        fluid = np.nonzero(self.flags == 2)[0]
        T = self.dtype.type
        terms = (v[fluid] + v[n + fluid]).astype(self.dtype) + v[2 * n + fluid]
        acc = np.float32(0)
        for t in terms.astype(T):
            acc = np.float32(acc + np.float32(t)) if self.dtype == np.float32 else np.float32(np.float64(acc) + t)
        return acc

"""CLbmSolver -- host mirror of the reference's solver facade over the CUDA C ABI.

Same public surface as the reference's ``CLbmSolver<T> : CLbmSkeleton<T>``
(src/CLbmSolver.hpp:67-1126, src/CLbmSkeleton.hpp:36-199): the constructor parametrises
(CLbmSkeleton::init), "reloads" (allocates device buffers, selects the kernel
specialisation) and resets; ``simulationStep`` alternates beta/alpha; the store*/set*
family moves rects between host arrays and the device.  The OpenCL queue/context/device
handles of the reference become a CUDA device ordinal (+ optional external streams).
Every method is a thin call into liblbm_b200.so -- there is no Python compute path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import capi
from .domain import CDomain
from .skeleton import LBM_UNITS, SIZE_DD_HOST, compute_parameters


class CError:
    """Stream-style error accumulator of the reference (src/lib/CError.hpp): truthy when set."""

    def __init__(self):
        self._msg = ""

    def __lshift__(self, text):
        self._msg += str(text)
        return self

    def __call__(self):
        return bool(self._msg)

    def getString(self):
        m, self._msg = self._msg, ""
        return m

    def __str__(self):
        return self._msg


class CLbmSolver:
    SIZE_DD_HOST = SIZE_DD_HOST

    def __init__(self, UID, device, BC, domain: CDomain, gravitation=(0.0, -9.81, 0.0),
                 viscosity=0.001308, computation_kernel_count=128, store_velocity=False,
                 store_density=False, timestep=-1.0, drivenCavityVelocity=(100.0, 0.0, 0.0, 1.0),
                 dtype=np.float32, smagorinsky_cs=0.0, beta_order=capi.LBM_BETA_ORDER_SHIPPED,
                 block_size=0, vector_width=0, compute_stream=None, comm_stream=None, params=None):
        self._lib = capi.load()
        self._h = ctypes.c_void_p()
        self._UID = int(UID)
        self.device = int(device)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("unsupported class type T")
        self.domain = domain
        self._BC = [[int(BC[a][s]) for s in range(2)] for a in range(3)]
        self.error = CError()
        self.computation_kernel_count = int(computation_kernel_count)
        self.store_velocity, self.store_density = bool(store_velocity), bool(store_density)
        self.timestep = timestep            # parsed but unused, like the reference (:226)
        self.smagorinsky_cs = float(smagorinsky_cs)
        self.beta_order = int(beta_order)
        self._block_size, self._vector_width = int(block_size), int(vector_width)
        self._streams = (compute_stream, comm_stream)
        # CLbmSkeleton::init (parametrisation in T)
        p = params if params is not None else compute_parameters(
            domain.getSize(), domain.getLength(), gravitation, viscosity, drivenCavityVelocity,
            dtype=self.dtype, strict=False)
        if p.error:
            self.error << p.error
        self.params = p
        self.domain_cells = tuple(domain.getSize())
        self.domain_cells_count = int(np.prod(self.domain_cells))
        self.d_cell_length, self.d_timestep = p.d_cell_length, p.d_timestep
        self.tau, self.inv_tau, self.inv_trt_tau = p.tau, p.inv_tau, p.inv_trt_tau
        self.gravitation = p.gravitation
        self.d_reynolds = p.d_reynolds
        self.d_drivenCavityVelocity = tuple(drivenCavityVelocity)
        self.drivenCavityVelocity = list(drivenCavityVelocity)   # CLbmSolver member (:71)
        self._u_lid = p.drivenCavityVelocity[0]
        if not self.error():
            self.reload()

    # ---------------------------------------------------------------- life cycle
    def reload(self):
        """CLbmSolver::reload (:272-617): (re)create device state and reset."""
        self._destroy()
        d = capi.lbm_desc()
        d.struct_size = ctypes.sizeof(capi.lbm_desc)
        d.device = self.device
        d.dtype = capi.LBM_F32 if self.dtype == np.float32 else capi.LBM_F64
        d.size[:] = self.domain_cells
        d.bc[:] = [self._BC[a][s] for a in range(3) for s in range(2)]
        d.inv_tau, d.tau = float(self.inv_tau), float(self.tau)
        d.gravitation[:] = [float(g) for g in self.gravitation]
        d.u_lid = float(self._u_lid)
        d.store_velocity, d.store_density = int(self.store_velocity), int(self.store_density)
        d.smagorinsky_cs = self.smagorinsky_cs
        d.beta_order = self.beta_order
        d.work_group_size = self.computation_kernel_count
        d.block_size, d.vector_width = self._block_size, self._vector_width
        d.compute_stream = self._streams[0]
        d.comm_stream = self._streams[1]
        h = ctypes.c_void_p()
        capi.check(None, self._lib.lbmCreate(ctypes.byref(h), ctypes.byref(d)))
        self._h = h

    def _destroy(self):
        if getattr(self, "_h", None):
            self._lib.lbmDestroy(self._h)
            self._h = ctypes.c_void_p()

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _ck(self, status):
        capi.check(self._h, status)

    def reset(self):
        self._ck(self._lib.lbmReset(self._h))

    def simulationStep(self):
        self._ck(self._lib.lbmStep(self._h))

    def simulationStepAlpha(self):
        self._ck(self._lib.lbmStepAlpha(self._h))

    def simulationStepBeta(self):
        self._ck(self._lib.lbmStepBeta(self._h))

    def simulationSteps(self, n):
        self._ck(self._lib.lbmSteps(self._h, int(n)))

    def wait(self):
        self._ck(self._lib.lbmWait(self._h))

    @property
    def simulation_step_counter(self):
        c = ctypes.c_uint64()
        self._ck(self._lib.lbmGetStepCounter(self._h, ctypes.byref(c)))
        return c.value

    @simulation_step_counter.setter
    def simulation_step_counter(self, v):
        self._ck(self._lib.lbmSetStepCounter(self._h, int(v)))

    def addDrivenCavityValue(self, value):
        """CLbmSolver::addDrivenCavityValue (:262-270)."""
        T = self.dtype.type
        self.drivenCavityVelocity[0] = T(T(self.drivenCavityVelocity[0]) + T(value))
        self._u_lid = T(T(self.drivenCavityVelocity[0]) * T(self.d_timestep))
        self._ck(self._lib.lbmSetDrivenCavityVelocity(self._h, float(self._u_lid)))

    # ---------------------------------------------------------------- field access
    def _cells(self, size):
        return self.domain_cells_count if size is None else int(size[0]) * int(size[1]) * int(size[2])

    def _store(self, fn, comps, dtype, dst, origin, size):
        n = comps * self._cells(size)
        if dst is None:
            dst = np.empty(n, dtype)
        if dst.dtype != dtype or dst.size < n or not dst.flags.c_contiguous:
            raise ValueError("destination must be a C-contiguous %s array of >= %d elements" % (dtype, n))
        self._ck(fn(self._h, dst.ctypes.data, capi.i3(origin), capi.i3(size)))
        return dst

    def _src(self, src, comps, dtype, size):
        a = np.ascontiguousarray(src, dtype=dtype).reshape(-1)
        if a.size < comps * self._cells(size):
            raise ValueError("source array too small")
        return a

    def storeDensityDistribution(self, dst=None, origin=None, size=None):
        return self._store(self._lib.lbmStoreDD, 19, self.dtype, dst, origin, size)

    def setDensityDistribution(self, src, origin=None, size=None, norm=None):
        a = self._src(src, 19, self.dtype, size)
        self._ck(self._lib.lbmSetDD(self._h, a.ctypes.data, capi.i3(origin), capi.i3(size), capi.i3(norm)))

    def storeVelocity(self, dst=None, origin=None, size=None):
        return self._store(self._lib.lbmStoreVelocity, 3, self.dtype, dst, origin, size)

    def setVelocity(self, src, origin=None, size=None):
        a = self._src(src, 3, self.dtype, size)
        self._ck(self._lib.lbmSetVelocity(self._h, a.ctypes.data, capi.i3(origin), capi.i3(size)))

    def storeDensity(self, dst=None, origin=None, size=None):
        return self._store(self._lib.lbmStoreDensity, 1, self.dtype, dst, origin, size)

    def setDensity(self, src, origin=None, size=None):
        a = self._src(src, 1, self.dtype, size)
        self._ck(self._lib.lbmSetDensity(self._h, a.ctypes.data, capi.i3(origin), capi.i3(size)))

    def storeFlags(self, dst=None, origin=None, size=None):
        return self._store(self._lib.lbmStoreFlags, 1, np.dtype(np.int32), dst, origin, size)

    def setFlags(self, src, origin=None, size=None):
        a = self._src(src, 1, np.int32, size)
        self._ck(self._lib.lbmSetFlags(self._h, a.ctypes.data, capi.i3(origin), capi.i3(size)))

    def getVelocityChecksum(self, host_order=True):
        out = ctypes.c_double()
        self._ck(self._lib.lbmChecksumVelocity(self._h, ctypes.byref(out), int(bool(host_order))))
        return np.float32(out.value) if host_order else out.value

    def debug_print(self, file=None):
        """reference src/CLbmSolver.hpp:1032-1057 (tiny domains only)."""
        import sys
        from . import debug
        (file or sys.stdout).write(debug.debug_print(self.storeDensityDistribution(), self.storeVelocity(),
                                                     self.storeDensity(), self.storeFlags()))

    def debugDD(self, dd_id=0, wrap_size=16, empty_line=16, file=None):
        """reference src/CLbmSolver.hpp:1062-1101."""
        import sys
        from . import debug
        (file or sys.stdout).write(debug.debugDD(self.storeDensityDistribution(), self.domain_cells_count,
                                                 dd_id, wrap_size, empty_line))

    # ---------------------------------------------------------------- device-side halo path
    def haloSlotMask(self, sync_kind, recv_dir, slots=capi.LBM_HALO_SLOTS_MINIMAL):
        m = ctypes.c_uint32()
        capi.check(None, self._lib.lbmHaloSlotMask(int(sync_kind), capi.i3(recv_dir), int(slots), ctypes.byref(m)))
        return m.value

    def haloBytes(self, size, slot_mask):
        b = ctypes.c_size_t()
        self._ck(self._lib.lbmHaloBytes(self._h, capi.i3(size), slot_mask, ctypes.byref(b)))
        return b.value

    def haloPack(self, origin, size, slot_mask, dev_ptr, stream=None):
        self._ck(self._lib.lbmHaloPack(self._h, capi.i3(origin), capi.i3(size), slot_mask, dev_ptr, stream))

    def haloUnpack(self, origin, size, buf_mask, write_mask, dev_ptr, stream=None):
        self._ck(self._lib.lbmHaloUnpack(self._h, capi.i3(origin), capi.i3(size), buf_mask, write_mask, dev_ptr, stream))

    def haloCopyPeer(self, src_origin, dst, dst_origin, size, slot_mask, stream=None):
        self._ck(self._lib.lbmHaloCopyPeer(self._h, capi.i3(src_origin), dst._h, capi.i3(dst_origin),
                                           capi.i3(size), slot_mask, stream))

    # one-sided peer-memory halo exchange (faces = CComm descriptors)
    def commAddFace(self, comm, slots=capi.LBM_HALO_SLOTS_MINIMAL):
        fid = ctypes.c_int()
        self._ck(self._lib.lbmCommAddFace(self._h, int(comm.getDstId()), capi.i3(comm.getSendOrigin()),
                                          capi.i3(comm.getRecvOrigin()), capi.i3(comm.getSendSize()),
                                          capi.i3(comm.getCommDirection()), int(slots), ctypes.byref(fid)))
        return fid.value

    def commIpcHandle(self, face_id):
        buf = ctypes.create_string_buffer(64)
        self._ck(self._lib.lbmCommGetIpcHandle(self._h, int(face_id), buf))
        return buf.raw

    def commConnectIpc(self, face_id, handle_bytes):
        buf = ctypes.create_string_buffer(bytes(handle_bytes), 64)
        self._ck(self._lib.lbmCommConnectIpc(self._h, int(face_id), buf))

    def commConnectLocal(self, face_id, peer, peer_face_id):
        self._ck(self._lib.lbmCommConnectLocal(self._h, int(face_id), peer._h, int(peer_face_id)))

    def commBeginSync(self, kind):
        self._ck(self._lib.lbmCommBeginSync(self._h, int(kind)))

    def commPush(self, kind, axis):
        self._ck(self._lib.lbmCommPush(self._h, int(kind), int(axis)))

    def commPull(self, kind, axis):
        self._ck(self._lib.lbmCommPull(self._h, int(kind), int(axis)))

    def commSync(self, kind):
        self._ck(self._lib.lbmCommSync(self._h, int(kind)))

    def commSetAxisOrder(self, order):
        """LBM_AXIS_ORDER_XYZ (the reference's CComm walk) or LBM_AXIS_ORDER_ZYX (x faces exchanged
        after the interior kernel, unsplit); accepts the constants or "xyz" / "zyx"."""
        if isinstance(order, str):
            order = {"xyz": capi.LBM_AXIS_ORDER_XYZ, "zyx": capi.LBM_AXIS_ORDER_ZYX}[order.lower()]
        self._ck(self._lib.lbmCommSetAxisOrder(self._h, int(order)))

    def commAxisOrder(self):
        o = ctypes.c_int()
        self._ck(self._lib.lbmCommGetAxisOrder(self._h, ctypes.byref(o)))
        return o.value

    def commStep(self):
        self._ck(self._lib.lbmCommStep(self._h))

    def stepShell(self, ghost_faces):
        self._ck(self._lib.lbmStepShell(self._h, int(ghost_faces)))

    def commStepTimed(self):
        """one overlapped step with device timestamps: ms from the fork to (shell done, interior done,
        exchange done, join)"""
        ms = (ctypes.c_float * 4)()
        self._ck(self._lib.lbmCommStepTimed(self._h, ms))
        return tuple(float(v) for v in ms)

    def stepShellComm(self, ghost_faces):
        """shell kernels on the comm stream: they run next to the interior kernel"""
        self._ck(self._lib.lbmStepShellComm(self._h, int(ghost_faces)))

    def stepInterior(self, ghost_faces):
        self._ck(self._lib.lbmStepInterior(self._h, int(ghost_faces)))

    def commWaitCompute(self):
        self._ck(self._lib.lbmStreamWaitStream(self._h, 1))

    def computeWaitComm(self):
        self._ck(self._lib.lbmStreamWaitStream(self._h, 0))

    def streams(self):
        a, b = ctypes.c_void_p(), ctypes.c_void_p()
        self._ck(self._lib.lbmGetStreams(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def devicePointer(self, which):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(self._lib.lbmGetDevicePointer(self._h, int(which), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def timerStart(self):
        self._ck(self._lib.lbmTimerStart(self._h))

    def timerStop(self):
        ms = ctypes.c_float()
        self._ck(self._lib.lbmTimerStop(self._h, ctypes.byref(ms)))
        return ms.value

    def launchCount(self):
        c = ctypes.c_uint64()
        self._ck(self._lib.lbmGetLaunchCount(self._h, ctypes.byref(c)))
        return c.value

    # per-kernel device timeline (the reference's PROFILE build, src/libcl/CCL.hpp:1752-1778)
    def profileEnable(self, mode=capi.LBM_PROFILE_EVENTS):
        self._ck(self._lib.lbmProfileEnable(self._h, int(mode)))

    def profileClear(self):
        self._ck(self._lib.lbmProfileClear(self._h))

    def profileEventCount(self):
        n, d = ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self._lib.lbmProfileEventCount(self._h, ctypes.byref(n), ctypes.byref(d)))
        return n.value, d.value

    def profileEvents(self):
        """[(kernel name, start_ns, end_ns)] since profileEnable / profileClear; synchronises"""
        out = []
        name = ctypes.create_string_buffer(128)
        t0, t1 = ctypes.c_uint64(), ctypes.c_uint64()
        for i in range(self.profileEventCount()[0]):
            self._ck(self._lib.lbmProfileGetEvent(self._h, i, name, 128, ctypes.byref(t0), ctypes.byref(t1)))
            out.append((name.value.decode(), t0.value, t1.value))
        return out

    def config(self):
        v, b, q = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._ck(self._lib.lbmGetConfig(self._h, ctypes.byref(v), ctypes.byref(b), ctypes.byref(q)))
        return dict(vector_width=v.value, block_size=b.value, wg_quirk=q.value)


__all__ = ["CLbmSolver", "CError", "LBM_UNITS"]

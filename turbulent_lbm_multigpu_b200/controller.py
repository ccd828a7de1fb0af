"""CController / CManager -- time loop, decomposition and halo exchange.

Host mirror of reference src/CController.hpp and src/CManager.hpp.  Public names and
semantics are the reference's (``run``, ``computeNextStep``, ``syncAlpha``, ``syncBeta``,
``addCommunication``, ``setGeometry``, ``getSolver``, ``getDomain`` / ``setSubdomainNums``,
``initSimulation``, ``startSimulation``, ``getController``); the OpenCL bring-up is replaced
by the CUDA C ABI and the host-staged MPI exchange by three interchangeable sync modes:

  "host"    the reference's algorithm verbatim: storeDensityDistribution -> send/recv of
            host buffers -> setDensityDistribution(+norm), one CComm after the other
            (works with any solver object; this is what the gloo CPU tests drive);
  "device"  device-side pack kernel -> NCCL send/recv on device buffers -> unpack kernel,
            one grouped exchange per axis phase, minimal 5-slot payload;
  "overlap" like "device", but the step is split: shell kernels first, then the exchange on
            the comm stream runs concurrently with the interior kernel;
  "p2p"     one-sided exchange over NVLink peer memory (CUDA IPC between processes): the push
            kernel packs a face and stores it straight into the neighbour's staging block and
            raises a flag, the neighbour waits on the flag on the device and unpacks; the whole
            overlapped step is ONE call into the library (lbmCommStep) -- no NCCL, no host sync.
            ``axis_order="zyx"`` exchanges the x faces after the interior kernel instead of
            splitting an x shell off (include/lbm_b200.h, LBM_AXIS_ORDER_ZYX).
"""
from __future__ import annotations

import time

import numpy as np

from . import capi
from .configuration import ConfigSingleton
from .domain import CComm, CDomain
from .skeleton import (FLAG_GHOST_LAYER, FLAG_OBSTACLE, FLAG_VELOCITY_INJECTION, LBM_UNITS)

MPI_TAG_ALPHA_SYNC, MPI_TAG_BETA_SYNC = 0, 1


class CController:
    def __init__(self, UID, domain: CDomain, BC, backend=None, device=0, sync_mode="host",
                 solver_factory=None, dtype=np.float32, slots=capi.LBM_HALO_SLOTS_MINIMAL,
                 config=None, axis_order=None, **solver_kw):
        self._UID = int(UID)
        self._domain = domain
        self._BC = [[int(BC[a][s]) for s in range(2)] for a in range(3)]
        self._comm_container = []
        self.backend = backend
        self.device = device
        self.sync_mode = sync_mode
        self.slots = slots
        # p2p mode: "xyz" | "zyx"; None = z,y,x when the decomposition cuts x (decided in connectFaces),
        # unless LBM_B200_AXIS_ORDER presets the library
        self.axis_order = axis_order
        self.dtype = np.dtype(dtype)
        self.config = config or ConfigSingleton.Instance()
        self.vector_checksum = 0.0
        self._solver_factory = solver_factory
        self._solver_kw = solver_kw
        self._halo = {}        # per CComm device buffers (device modes)
        self.cLbmPtr = None
        if self.initLBMSolver() == -1:
            raise RuntimeError("Initialization of LBM Solver failed!")

    # ---------------------------------------------------------------- bring-up
    def initLBMSolver(self):
        """reference src/CController.hpp:83-226 minus the OpenCL platform/context/queue."""
        cfg = self.config
        store = bool(cfg.do_visualization or cfg.debug_mode)
        if self._solver_factory is not None:
            self.cLbmPtr = self._solver_factory(self._UID, self._domain, self._BC, cfg)
        else:
            from .solver import CLbmSolver
            kw = dict(store_velocity=store, store_density=store)
            kw.update(self._solver_kw)
            self.cLbmPtr = CLbmSolver(self._UID, self.device, self._BC, self._domain, cfg.gravitation,
                                      cfg.viscosity, cfg.computation_kernel_count,
                                      timestep=cfg.timestep, drivenCavityVelocity=cfg.drivenCavityVelocity,
                                      dtype=self.dtype,
                                      smagorinsky_cs=getattr(cfg, "smagorinsky_constant", 0.0), **kw)
        err = getattr(self.cLbmPtr, "error", None)
        if err is not None and callable(err) and err():
            print(self.cLbmPtr.error.getString())
            return -1
        if self.axis_order is not None and hasattr(self.cLbmPtr, "commSetAxisOrder"):
            self.cLbmPtr.commSetAxisOrder(self.axis_order)
        if hasattr(self.cLbmPtr, "wait"):
            self.cLbmPtr.wait()
        return 0

    # ---------------------------------------------------------------- reference sync (host staged)
    def _host_sync(self, beta):
        """syncAlpha (:265-320) / syncBeta (:322-383), one CComm after the other."""
        import torch
        s = self.cLbmPtr
        for c in self._comm_container:
            if beta:   # the ghost layer data is sent back to its origin (:337-341)
                send_size, recv_size = c.getRecvSize(), c.getSendSize()
                send_origin, recv_origin = c.getRecvOrigin(), c.getSendOrigin()
                normal = c.getCommDirection()
            else:
                send_size, recv_size = c.getSendSize(), c.getRecvSize()
                send_origin, recv_origin = c.getSendOrigin(), c.getRecvOrigin()
                normal = None
            send_buffer = np.ascontiguousarray(s.storeDensityDistribution(origin=send_origin, size=send_size))
            recv_buffer = np.empty(int(np.prod(recv_size)) * 19, send_buffer.dtype)
            works = self.backend.exchange([(c.getDstId(), torch.from_numpy(send_buffer), torch.from_numpy(recv_buffer))])
            for w in works:
                w.wait()
            s.setDensityDistribution(recv_buffer, recv_origin, recv_size, normal)
            if hasattr(s, "wait"):
                s.wait()

    # ---------------------------------------------------------------- device-side sync
    def _halo_buffers(self, c, beta):
        """torch device tensors for one face and one sync kind, created once."""
        import torch
        key = (id(c), beta)
        if key not in self._halo:
            s = self.cLbmPtr
            kind = capi.LBM_SYNC_BETA if beta else capi.LBM_SYNC_ALPHA
            # the same face on the neighbour has the opposite direction
            my_dir = c.getCommDirection()
            peer_dir = tuple(-d for d in my_dir)
            send_mask = s.haloSlotMask(kind, peer_dir, self.slots)   # what the neighbour consumes
            recv_mask = s.haloSlotMask(kind, my_dir, self.slots)     # what I consume
            write_mask = s.haloSlotMask(kind, my_dir, capi.LBM_HALO_SLOTS_MINIMAL) if beta else recv_mask
            size = c.getSendSize()
            tdt = torch.float32 if self.dtype == np.float32 else torch.float64
            dev = torch.device("cuda", self.device)
            nsend = s.haloBytes(size, send_mask) // self.dtype.itemsize
            nrecv = s.haloBytes(size, recv_mask) // self.dtype.itemsize
            self._halo[key] = dict(send=torch.empty(nsend, dtype=tdt, device=dev),
                                   recv=torch.empty(nrecv, dtype=tdt, device=dev),
                                   send_mask=send_mask, recv_mask=recv_mask, write_mask=write_mask)
        return self._halo[key]

    def _device_sync(self, beta):
        """One grouped NCCL exchange per axis phase (x, then y, then z: later axes carry the
        rims written by earlier ones, exactly like the reference's sequential CComm walk)."""
        import torch
        s = self.cLbmPtr
        comm_stream = self._torch_comm_stream()
        with torch.cuda.stream(comm_stream):
            for axis in range(3):
                faces = [c for c in self._comm_container if c.axis == axis]
                if not faces:
                    continue
                pairs = []
                for c in faces:
                    hb = self._halo_buffers(c, beta)
                    origin = c.getRecvOrigin() if beta else c.getSendOrigin()
                    s.haloPack(origin, c.getSendSize(), hb["send_mask"], hb["send"].data_ptr(),
                               comm_stream.cuda_stream)
                    pairs.append((c.getDstId(), hb["send"], hb["recv"]))
                for w in self.backend.exchange(pairs):
                    w.wait()
                for c in faces:
                    hb = self._halo_buffers(c, beta)
                    origin = c.getSendOrigin() if beta else c.getRecvOrigin()
                    s.haloUnpack(origin, c.getSendSize(), hb["recv_mask"], hb["write_mask"],
                                 hb["recv"].data_ptr(), comm_stream.cuda_stream)

    def _torch_comm_stream(self):
        import torch
        if not hasattr(self, "_comm_stream"):
            _, comm = self.cLbmPtr.streams()
            self._comm_stream = torch.cuda.ExternalStream(comm, device=torch.device("cuda", self.device))
        return self._comm_stream

    def connectFaces(self):
        """p2p mode: register every CComm as a face, exchange the CUDA-IPC handles of the receive
        blocks with the neighbours (torch.distributed object gather) and map them."""
        import os
        s = self.cLbmPtr
        if self.axis_order is None and not os.environ.get("LBM_B200_AXIS_ORDER") \
                and any(c.axis == 0 for c in self._comm_container):
            # every rank of an x-cutting decomposition has an x neighbour, so all ranks agree
            s.commSetAxisOrder(capi.LBM_AXIS_ORDER_ZYX)
        self._face_ids = [s.commAddFace(c, self.slots) for c in self._comm_container]
        mine = {}
        for c, fid in zip(self._comm_container, self._face_ids):
            mine[(self._UID, c.getDstId(), c.axis, c.getCommDirection()[c.axis])] = s.commIpcHandle(fid)
        dist = self.backend.dist
        gathered = [None] * self.backend.world
        dist.all_gather_object(gathered, mine, group=self.backend.group)
        table = {}
        for g in gathered:
            table.update(g)
        for c, fid in zip(self._comm_container, self._face_ids):
            key = (c.getDstId(), self._UID, c.axis, -c.getCommDirection()[c.axis])
            s.commConnectIpc(fid, table[key])
        dist.barrier(group=self.backend.group)

    def ghost_faces(self):
        m = 0
        for a in range(3):
            for side in range(2):
                if self._BC[a][side] == FLAG_GHOST_LAYER:
                    m |= 1 << (2 * a + side)
        return m

    def _p2p_sync(self, kind):
        s = self.cLbmPtr
        s.commWaitCompute()
        s.commSync(kind)
        s.computeWaitComm()

    def syncAlpha(self):
        if self.sync_mode == "host":
            self._host_sync(beta=False)
        elif self.sync_mode == "p2p":
            self._p2p_sync(capi.LBM_SYNC_ALPHA)
        else:
            s = self.cLbmPtr
            s.commWaitCompute()
            self._device_sync(beta=False)
            s.computeWaitComm()

    def syncBeta(self):
        if self.sync_mode == "host":
            self._host_sync(beta=True)
        elif self.sync_mode == "p2p":
            self._p2p_sync(capi.LBM_SYNC_BETA)
        else:
            s = self.cLbmPtr
            s.commWaitCompute()
            self._device_sync(beta=True)
            s.computeWaitComm()

    def computeNextStep(self):
        """reference src/CController.hpp:385-391."""
        s = self.cLbmPtr
        if self.sync_mode == "p2p":
            s.commStep()                    # shell -> (push/pull || interior) -> join, in the library
            return
        if self.sync_mode == "overlap" and self._comm_container:
            faces = self.ghost_faces()
            beta_step = (s.simulation_step_counter & 1) == 0
            s.commWaitCompute()             # fork: the comm stream follows the previous step
            s.stepShellComm(faces)          # cells next to ghost faces, on the comm stream ...
            s.stepInterior(faces)           # ... next to the interior kernel on the compute stream
            self._device_sync(beta=beta_step)
            s.computeWaitComm()             # the next step needs the halo
            return
        s.simulationStep()
        if s.simulation_step_counter & 1:
            self.syncBeta()
        else:
            self.syncAlpha()

    # ---------------------------------------------------------------- run loop
    def run(self, quiet=False):
        """reference src/CController.hpp:396-522: loop, MLUPS/bandwidth block, benchmark .ini."""
        cfg = self.config
        domain_size = self._domain.getSize()
        loops = cfg.loops if cfg.loops >= 0 else 100
        floats_per_cell = 19.0 * 2.0 + 1.0
        if cfg.do_visualization or cfg.debug_mode:
            floats_per_cell += 3
        s = self.cLbmPtr
        s.wait()
        t0 = time.perf_counter()
        for i in range(loops):
            self.computeNextStep()
        s.wait()
        seconds = time.perf_counter() - t0
        self.seconds = seconds
        cells = int(np.prod(domain_size))
        fps = loops / seconds if seconds > 0 else float("inf")
        mlups = fps * cells * 1e-6
        self.mlups = mlups
        if not quiet:
            print()
            print("Cube: [%d, %d, %d]" % tuple(domain_size))
            print("Seconds: %g" % seconds)
            print("FPS: %g" % fps)
            print("MLUPS: %g" % mlups)
            print("Bandwidth: %g MB/s (RW, bidirectional)" % (mlups * floats_per_cell * self.dtype.itemsize))
            if cfg.debug_mode:
                self.vector_checksum = s.getVelocityChecksum()
                print("Checksum: %.8f" % (float(self.vector_checksum) * 1000.0))
            print("done.")
        # BENCHMARK block of src/CController.hpp:445-477: rank 0 appends the max-over-ranks time and the
        # whole-domain rates to ./output/benchmark/benchmark_<np>.ini (compile-time switch in the reference,
        # LBM_B200_BENCHMARK in the environment here, like the C++ host)
        import os
        if os.environ.get("LBM_B200_BENCHMARK"):
            gtime = seconds
            if self.backend is not None and getattr(self.backend, "world", 1) > 1:
                gtime = self.backend.max_over_ranks(seconds)
            if self._UID <= 0:
                nproc = int(np.prod(cfg.subdomain_num))
                gfps = loops / gtime if gtime > 0 else float("inf")
                gmlups = gfps * float(np.prod(cfg.domain_size)) * 1e-6
                os.makedirs(os.path.join("output", "benchmark"), exist_ok=True)
                with open(os.path.join("output", "benchmark", "benchmark_%d.ini" % nproc), "a") as f:
                    f.write("CUBE_X : %d\nCUBE_Y : %d\nCUBE_Z : %d\n" % tuple(cfg.domain_size))
                    f.write("SECONDS : %g\nFPS : %g\nMLUPS : %g\nBANDWIDTH : %g\n\n"
                            % (gtime, gfps, gmlups, gmlups * floats_per_cell * self.dtype.itemsize))
        # PROFILE block of src/CController.hpp:503-519 (run-time switch: LBM_B200_PROFILE, or
        # CLbmSolver.profileEnable before run)
        count = getattr(s, "profileEventCount", None)   # an injected solver need not carry the timeline
        if count is not None and count()[0] > 0:
            from .profiler import CProfiler, profile_file_name
            self.profiler = CProfiler()
            self.profiler.collect(s)
            nproc = int(np.prod(cfg.subdomain_num))
            self.profiler.saveProfile(profile_file_name(nproc, self._UID), nproc, self._UID)
        return 0

    def addCommunication(self, comm: CComm):
        self._comm_container.append(comm)

    def setGeometry(self):
        """reference src/CController.hpp:531-546: lid on y = Sy-2, x in [1,Sx-2], z in [1,Sz-2]."""
        S = self._domain.getSize()
        origin = (1, S[1] - 2, 1)
        size = (S[0] - 2, 1, S[2] - 2)
        src = np.full(size[0] * size[2], FLAG_VELOCITY_INJECTION, np.int32)
        self.cLbmPtr.setFlags(src, origin, size)

    def getSolver(self):
        return self.cLbmPtr

    def setSolver(self, solver):
        self.cLbmPtr = solver

    def getDomain(self):
        return self._domain

    def getUid(self):
        return self._UID

    def getComms(self):
        return list(self._comm_container)


class CManager:
    """reference src/CManager.hpp:15-224."""

    def __init__(self, domain: CDomain, subdomainNums, **controller_kw):
        self._domain = domain
        self._lbm_controller = None
        self._controller_kw = controller_kw
        self.setSubdomainNums(subdomainNums)

    def getDomain(self):
        return self._domain

    def setDomain(self, grid):
        self._domain = grid

    def getSubdomainNums(self):
        return self._subdomain_nums

    def setSubdomainNums(self, subdomainNums):
        do_size = self._domain.getSize()
        n = tuple(int(v) for v in subdomainNums)
        if any(do_size[a] % n[a] != 0 for a in range(3)):
            raise ValueError("Number of subdomains does not match with the grid size!")
        self._subdomain_size = tuple(do_size[a] // n[a] for a in range(3))
        self._subdomain_nums = n
        # divided in the simulation type T like the reference (CVector<3,T>, :66-69)
        T = np.dtype(self._controller_kw.get("dtype", np.float32)).type
        L = self._domain.getLength()
        self._subdomain_length = tuple(T(T(L[a]) / T(n[a])) for a in range(3))

    def getSubdomainSize(self):
        return self._subdomain_size

    def layout(self, my_rank):
        """rank -> (nx,ny,nz), BC[3][2], CComm list (reference :78-199)."""
        NX, NY, NZ = self._subdomain_nums
        S = self._subdomain_size
        rid = max(int(my_rank), 0)
        tmp = rid
        nx = tmp % NX
        tmp //= NX
        ny = tmp % NY
        tmp //= NY
        nz = tmp
        coords = (nx, ny, nz)
        BC = [[FLAG_GHOST_LAYER, FLAG_GHOST_LAYER] for _ in range(3)]
        for a in range(3):
            if coords[a] == 0:
                BC[a][0] = FLAG_OBSTACLE
            if coords[a] == self._subdomain_nums[a] - 1:
                BC[a][1] = FLAG_OBSTACLE
        stride = (1, NX, NX * NY)
        comms = []
        for a in range(3):
            face = [S[0], S[1], S[2]]
            face[a] = 1
            face = tuple(face)
            if BC[a][0] == FLAG_GHOST_LAYER:
                so, ro, d = [0, 0, 0], [0, 0, 0], [0, 0, 0]
                so[a], ro[a], d[a] = 1, 0, 1
                comms.append(CComm(rid - stride[a], face, face, tuple(so), tuple(ro), tuple(d)))
            if BC[a][1] == FLAG_GHOST_LAYER:
                so, ro, d = [0, 0, 0], [0, 0, 0], [0, 0, 0]
                so[a], ro[a], d[a] = S[a] - 2, S[a] - 1, -1
                comms.append(CComm(rid + stride[a], face, face, tuple(so), tuple(ro), tuple(d)))
        origin = tuple(coords[a] * S[a] for a in range(3))
        return rid, coords, BC, comms, origin

    def initSimulation(self, my_rank):
        rid, coords, BC, comms, origin = self.layout(my_rank)
        subdomain = CDomain(rid, self._subdomain_size, origin, self._subdomain_length)
        self._lbm_controller = CController(rid, subdomain, BC, **self._controller_kw)
        for c in comms:
            self._lbm_controller.addCommunication(c)
        if coords[1] == self._subdomain_nums[1] - 1:
            self._lbm_controller.setGeometry()
        if self._lbm_controller.sync_mode == "p2p":
            self._lbm_controller.connectFaces()

    def startSimulation(self, **kw):
        if self._lbm_controller is None:
            raise RuntimeError("CManager: Initialize the simulation before starting it!")
        return self._lbm_controller.run(**kw)

    def getController(self):
        return self._lbm_controller

    def setController(self, c):
        self._lbm_controller = c


class InProcessSimulation:
    """All sub-domains of a decomposed run driven by ONE host thread (single-process
    multi-device, or several sub-domains on one device): what the reference does with one MPI
    rank per sub-domain, with the halo moved by ONE peer-copy kernel per face
    (pack + send + unpack fused, NVLink peer stores when the devices differ)."""

    def __init__(self, domain: CDomain, subdomainNums, devices=None, slots=capi.LBM_HALO_SLOTS_MINIMAL,
                 transport="copy", overlap=False, **controller_kw):
        self.nums = tuple(int(v) for v in subdomainNums)
        self.nranks = self.nums[0] * self.nums[1] * self.nums[2]
        if transport == "p2p" and self.nranks > (len(set(devices)) if devices else 1):
            capi.want_hardware_queues()     # sub-domains share a GPU (effective when CUDA is not initialised yet)
        self.slots = slots
        self.transport = transport      # "copy": one peer-copy kernel per face + host phase sync
        self.overlap = overlap          # "p2p": push/flag/pull faces, no host synchronisation
        self.controllers = []
        for r in range(self.nranks):
            kw = dict(controller_kw)
            kw["device"] = (devices[r % len(devices)] if devices else 0)
            kw["sync_mode"] = "inprocess"
            m = CManager(domain, subdomainNums, **kw)
            m.initSimulation(r)
            self.controllers.append(m.getController())
        self.sub_size = m.getSubdomainSize()
        if transport == "p2p":
            self._connect_local()

    def _connect_local(self):
        ids = {}
        for ctrl in self.controllers:
            s = ctrl.getSolver()
            for c in ctrl.getComms():
                ids[(ctrl.getUid(), c.getDstId(), c.axis, c.getCommDirection()[c.axis])] = s.commAddFace(c, self.slots)
        for ctrl in self.controllers:
            s = ctrl.getSolver()
            for c in ctrl.getComms():
                d = c.getCommDirection()[c.axis]
                fid = ids[(ctrl.getUid(), c.getDstId(), c.axis, d)]
                pid = ids[(c.getDstId(), ctrl.getUid(), c.axis, -d)]
                s.commConnectLocal(fid, self.controllers[c.getDstId()].getSolver(), pid)

    def _sync_p2p(self, beta, x_after_interior=False):
        """Every push of an axis is enqueued before any wait of that axis, so the device-side
        flag waits can never be ordered ahead of the push they wait for."""
        kind = capi.LBM_SYNC_BETA if beta else capi.LBM_SYNC_ALPHA
        solvers = [c.getSolver() for c in self.controllers]
        for s in solvers:
            s.commBeginSync(kind)
        zyx = solvers[0].commAxisOrder() == capi.LBM_AXIS_ORDER_ZYX
        for axis in ((2, 1, 0) if zyx else (0, 1, 2)):
            if axis == 0 and x_after_interior:      # lbmCommStep's z,y,x order: x faces follow the interior
                for s in solvers:
                    s.commWaitCompute()
            for s in solvers:
                s.commPush(kind, axis)
            for s in solvers:
                s.commPull(kind, axis)

    def _sync(self, beta):
        kind = capi.LBM_SYNC_BETA if beta else capi.LBM_SYNC_ALPHA
        for axis in range(3):
            issued = False
            for ctrl in self.controllers:
                src = ctrl.getSolver()
                for c in ctrl.getComms():
                    if c.axis != axis:
                        continue
                    peer = self.controllers[c.getDstId()]
                    back = next(k for k in peer.getComms() if k.getDstId() == ctrl.getUid() and k.axis == axis)
                    mask = src.haloSlotMask(kind, back.getCommDirection(), self.slots)
                    if beta:   # my ghost layer -> the peer's outermost real layer
                        so, do = c.getRecvOrigin(), back.getSendOrigin()
                        mask &= src.haloSlotMask(kind, back.getCommDirection(), capi.LBM_HALO_SLOTS_MINIMAL)
                    else:      # my outermost real layer -> the peer's ghost layer
                        so, do = c.getSendOrigin(), back.getRecvOrigin()
                    src.haloCopyPeer(so, peer.getSolver(), do, c.getSendSize(), mask)
                    issued = True
            if issued:          # later axes read the rims this phase wrote
                for ctrl in self.controllers:
                    ctrl.getSolver().wait()

    def computeNextStep(self):
        solvers = [c.getSolver() for c in self.controllers]
        if self.transport == "p2p":
            beta_step = (solvers[0].simulation_step_counter & 1) == 0
            zyx = solvers[0].commAxisOrder() == capi.LBM_AXIS_ORDER_ZYX
            if self.overlap:
                for ctrl, s in zip(self.controllers, solvers):
                    split = ctrl.ghost_faces() & (~3 if zyx else ~0)     # z,y,x: no x shell
                    s.commWaitCompute()
                    s.stepShellComm(split)
                    s.stepInterior(split)
            else:
                for s in solvers:
                    s.simulationStep()
                    s.commWaitCompute()
            self._sync_p2p(beta=beta_step, x_after_interior=self.overlap and zyx)
            for s in solvers:
                s.computeWaitComm()
            return
        for s in solvers:
            s.simulationStep()
        for s in solvers:
            s.wait()
        self._sync(beta=bool(solvers[0].simulation_step_counter & 1))

    def run(self, loops):
        for _ in range(loops):
            self.computeNextStep()
        for ctrl in self.controllers:
            ctrl.getSolver().wait()


def validation_domain_size(domain_size, subdomain_num):
    """reference src/main.cpp:358-361."""
    return tuple(domain_size[a] - 2 * (subdomain_num[a] - 1) for a in range(3))


def validation_sub_origin(rank, subdomain_num, local_size_without_halo):
    """reference src/main.cpp:374-386."""
    NX, NY, _ = subdomain_num
    tmp = rank
    nx = tmp % NX
    tmp //= NX
    ny = tmp % NY
    tmp //= NY
    nz = tmp
    return (1 + nx * local_size_without_halo[0], 1 + ny * local_size_without_halo[1],
            1 + nz * local_size_without_halo[2])


__all__ = ["CController", "CManager", "InProcessSimulation", "LBM_UNITS", "validation_domain_size",
           "validation_sub_origin"]

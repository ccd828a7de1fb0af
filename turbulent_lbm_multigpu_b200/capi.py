"""ctypes binding of liblbm_b200.so -- the same C ABI a C++/cgo/JNI host would bind
(include/lbm_b200.h).  There is no fallback: if the CUDA library is missing or cannot be
loaded this module raises, it never routes to a CPU implementation.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LBM_B200_LIB: tuning hook to load an alternative build of the SAME CUDA library
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(HERE, "lib", "liblbm_b200.so")


def want_hardware_queues(n=32):
    """Sub-domains that SHARE one GPU (tests, InProcessSimulation with more sub-domains than GPUs) keep two
    streams each, and a flag wait of the one-sided exchange spins until the neighbour's kernels have run: with
    CUDA's default of 8 hardware queues a neighbour's stream can end up queued behind such a wait (reproduced,
    profiles/r2/n1_pytest_host_cpp_8_hw_queues.log).  CUDA reads the variable when the context is created, so
    this must run before the first CUDA call of the process; an explicit setting of the user wins.  Not needed --
    and not touched -- with one rank per GPU (bench.py, torchrun)."""
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", str(n))

LBM_OK = 0
LBM_ERR_INVALID, LBM_ERR_CUDA, LBM_ERR_NO_DEVICE, LBM_ERR_TIMEOUT = 1, 2, 3, 4
LBM_F32, LBM_F64 = 0, 1
LBM_BETA_ORDER_SHIPPED, LBM_BETA_ORDER_LINEAR = 0, 1
LBM_HALO_SLOTS_REFERENCE, LBM_HALO_SLOTS_MINIMAL = 0, 1
LBM_SYNC_ALPHA, LBM_SYNC_BETA = 0, 1
LBM_AXIS_ORDER_XYZ, LBM_AXIS_ORDER_ZYX = 0, 1
LBM_BUF_DD, LBM_BUF_FLAGS, LBM_BUF_VELOCITY, LBM_BUF_DENSITY = 0, 1, 2, 3
LBM_PROFILE_EVENTS, LBM_PROFILE_NVTX = 1, 2

c_int3 = ctypes.c_int * 3


class lbm_desc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("device", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("size", ctypes.c_int32 * 3),
        ("bc", ctypes.c_int32 * 6),
        ("inv_tau", ctypes.c_double),
        ("tau", ctypes.c_double),
        ("gravitation", ctypes.c_double * 3),
        ("u_lid", ctypes.c_double),
        ("store_velocity", ctypes.c_int32),
        ("store_density", ctypes.c_int32),
        ("smagorinsky_cs", ctypes.c_double),
        ("beta_order", ctypes.c_int32),
        ("work_group_size", ctypes.c_int32),
        ("block_size", ctypes.c_int32),
        ("vector_width", ctypes.c_int32),
        ("compute_stream", ctypes.c_void_p),
        ("comm_stream", ctypes.c_void_p),
    ]


#: every symbol include/lbm_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u32, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
_ip = ctypes.POINTER(ctypes.c_int)
SYMBOLS = {
    "lbmGetDeviceCount": (_i, [_ip]),
    "lbmGetVersion": (_i, []),
    "lbmGetLastErrorString": (ctypes.c_char_p, [_vp]),
    "lbmCreate": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(lbm_desc)]),
    "lbmDestroy": (_i, [_vp]),
    "lbmReset": (_i, [_vp]),
    "lbmStep": (_i, [_vp]),
    "lbmStepAlpha": (_i, [_vp]),
    "lbmStepBeta": (_i, [_vp]),
    "lbmSteps": (_i, [_vp, _i]),
    "lbmWait": (_i, [_vp]),
    "lbmGetStepCounter": (_i, [_vp, ctypes.POINTER(_u64)]),
    "lbmSetStepCounter": (_i, [_vp, _u64]),
    "lbmSetDrivenCavityVelocity": (_i, [_vp, ctypes.c_double]),
    "lbmStoreDD": (_i, [_vp, _vp, _ip, _ip]),
    "lbmSetDD": (_i, [_vp, _vp, _ip, _ip, _ip]),
    "lbmStoreVelocity": (_i, [_vp, _vp, _ip, _ip]),
    "lbmSetVelocity": (_i, [_vp, _vp, _ip, _ip]),
    "lbmStoreDensity": (_i, [_vp, _vp, _ip, _ip]),
    "lbmSetDensity": (_i, [_vp, _vp, _ip, _ip]),
    "lbmStoreFlags": (_i, [_vp, _vp, _ip, _ip]),
    "lbmSetFlags": (_i, [_vp, _vp, _ip, _ip]),
    "lbmChecksumVelocity": (_i, [_vp, ctypes.POINTER(ctypes.c_double), _i]),
    "lbmHaloSlotMask": (_i, [_i, _ip, _i, ctypes.POINTER(_u32)]),
    "lbmHaloBytes": (_i, [_vp, _ip, _u32, ctypes.POINTER(ctypes.c_size_t)]),
    "lbmHaloPack": (_i, [_vp, _ip, _ip, _u32, _vp, _vp]),
    "lbmHaloUnpack": (_i, [_vp, _ip, _ip, _u32, _u32, _vp, _vp]),
    "lbmHaloCopyPeer": (_i, [_vp, _ip, _vp, _ip, _ip, _u32, _vp]),
    "lbmCommAddFace": (_i, [_vp, _i, _ip, _ip, _ip, _ip, _i, _ip]),
    "lbmCommFaceCount": (_i, [_vp, _ip]),
    "lbmCommGetIpcHandle": (_i, [_vp, _i, _vp]),
    "lbmCommConnectIpc": (_i, [_vp, _i, _vp]),
    "lbmCommConnectLocal": (_i, [_vp, _i, _vp, _i]),
    "lbmCommBeginSync": (_i, [_vp, _i]),
    "lbmCommPush": (_i, [_vp, _i, _i]),
    "lbmCommPull": (_i, [_vp, _i, _i]),
    "lbmCommSync": (_i, [_vp, _i]),
    "lbmCommSetAxisOrder": (_i, [_vp, _i]),
    "lbmCommGetAxisOrder": (_i, [_vp, _ip]),
    "lbmCommStep": (_i, [_vp]),
    "lbmStepShell": (_i, [_vp, _i]),
    "lbmStepShellComm": (_i, [_vp, _i]),
    "lbmCommStepTimed": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "lbmGetSlotStride": (_i, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
    "lbmStepInterior": (_i, [_vp, _i]),
    "lbmStreamWaitStream": (_i, [_vp, _i]),
    "lbmGetStreams": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "lbmGetDevicePointer": (_i, [_vp, _i, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "lbmTimerStart": (_i, [_vp]),
    "lbmTimerStop": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "lbmGetLaunchCount": (_i, [_vp, ctypes.POINTER(_u64)]),
    "lbmGetConfig": (_i, [_vp, _ip, _ip, _ip]),
    "lbmProfileEnable": (_i, [_vp, _i]),
    "lbmProfileClear": (_i, [_vp]),
    "lbmProfileEventCount": (_i, [_vp, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "lbmProfileGetEvent": (_i, [_vp, _u64, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
}

_lib = None


class LbmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("liblbm_b200 status %d: %s" % (status, message))
        self.status = status


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m turbulent_lbm_multigpu_b200.build` "
                "(or __graft_entry__.build()); the framework has no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def i3(v):
    return None if v is None else c_int3(*[int(x) for x in v])


def check(handle, status):
    if status != LBM_OK:
        msg = load().lbmGetLastErrorString(handle)
        raise LbmError(status, (msg or b"").decode("utf-8", "replace"))

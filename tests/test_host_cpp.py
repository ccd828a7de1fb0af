"""The C++ host facade (turbulent_lbm_multigpu_b200/host: CLbmSolver<T>, CController<T>,
CManager<T>, CComm<T>, CDomain<T>, CConfiguration<T>, CLbmSkeleton<T> + the lbm_b200 driver).

CPU part: the driver builds, parses the reference's conf.xml schema, and its decomposition
tables / parametrisation are bit-identical to the Python mirror (which tests/test_host_logic.py
and tests/test_oracle.py pin against the oracle).  GPU part: the reference's validate mode and
field parity against the oracle."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from turbulent_lbm_multigpu_b200.controller import CManager
from turbulent_lbm_multigpu_b200.domain import CDomain
from turbulent_lbm_multigpu_b200.host import build as host_build
from turbulent_lbm_multigpu_b200.skeleton import compute_parameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONF = """<?xml version="1.0" encoding="ISO-8859-1"?>
<lbm-configuration>
  <!-- physics -->
  <physics>
    <viscosity>0.001308</viscosity>
    <gravitation><x>0</x><y>-9.81</y><z>0</z></gravitation>
    <cavity-velocity><x>100</x><y>0</y><z>0</z><w>1</w></cavity-velocity>
    %(extra)s
  </physics>
  <grid>
    <domain-size><x>%(dx)d</x><y>%(dy)d</y><z>%(dz)d</z></domain-size>
    <subdomain-num><x>%(nx)d</x><y>%(ny)d</y><z>%(nz)d</z></subdomain-num>
    <domian-length><x>%(lx)g</x><y>%(ly)g</y><z>%(lz)g</z></domian-length>
  </grid>
  <simulation>
    <loops>%(loops)d</loops>
    <!-- timestep default: -1 for automatic detection -->
    <timestep>-1.0</timestep>
    <visualization><VTK>0</VTK></visualization>
    <validate>%(validate)d</validate>
  </simulation>
  <device>
    <kernel-count>128</kernel-count>
    <device-number>0</device-number>
  </device>
</lbm-configuration>
"""


def write_conf(path, D, nums, L=(0.1, 0.1, 0.1), loops=40, validate=0, extra=""):
    path.write_text(CONF % dict(dx=D[0], dy=D[1], dz=D[2], nx=nums[0], ny=nums[1], nz=nums[2],
                                lx=L[0], ly=L[1], lz=L[2], loops=loops, validate=validate, extra=extra))
    return str(path)


@pytest.fixture(scope="module")
def exe():
    return host_build.build()


def run(exe, *args, ok=(0,)):
    p = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    assert p.returncode in ok, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    return p.stdout


@pytest.mark.parametrize("D,nums", [((96, 32, 32), (3, 1, 1)), ((24, 36, 48), (2, 3, 4)), ((16, 16, 24), (1, 1, 2))])
def test_cpp_decomposition_tables_match(exe, tmp_path, D, nums):
    out = run(exe, "-c", write_conf(tmp_path / "c.xml", D, nums), "--dump-layout")
    m = CManager.__new__(CManager)
    m._domain = CDomain(-1, D, (0, 0, 0), (0.1,) * 3)
    m._controller_kw = {}
    m.setSubdomainNums(nums)
    lines = out.strip().splitlines()
    assert lines[0] == "subdomain_size %d %d %d" % m.getSubdomainSize()
    expected = []
    for r in range(int(np.prod(nums))):
        rid, coords, BC, comms, origin = m.layout(r)
        expected.append("rank %d origin %d %d %d bc %d %d %d %d %d %d ncomm %d" % (
            (r,) + tuple(origin) + (BC[0][0], BC[0][1], BC[1][0], BC[1][1], BC[2][0], BC[2][1], len(comms))))
        for c in comms:
            expected.append("  comm dst %d send_size %d %d %d recv_size %d %d %d send_origin %d %d %d "
                            "recv_origin %d %d %d dir %d %d %d" % (
                                (c.getDstId(),) + tuple(c.getSendSize()) + tuple(c.getRecvSize())
                                + tuple(c.getSendOrigin()) + tuple(c.getRecvOrigin()) + tuple(c.getCommDirection())))
    assert lines[1:] == expected


@pytest.mark.parametrize("size,dbl", [(16, False), (64, False), (128, False), (256, False), (512, False),
                                      (64, True), (384, True)])
def test_cpp_parametrisation_bit_identical(exe, size, dbl):
    args = ["-S", size, "--dump-params"] + (["--double"] if dbl else [])
    out = run(exe, *args)
    got = {}
    for ln in out.splitlines():
        m = re.match(r"(\w+) \S+ 0x([0-9a-f]+)$", ln)
        if m:
            got[m.group(1)] = int(m.group(2), 16)
    dtype = np.float64 if dbl else np.float32
    p = compute_parameters((size,) * 3, (0.1,) * 3, dtype=dtype)

    def bits(v):
        return struct.unpack("<Q", struct.pack("<d", float(v)))[0] if dbl else struct.unpack("<I", struct.pack("<f", float(v)))[0]
    exp = dict(d_cell_length=p.d_cell_length, d_timestep=p.d_timestep, tau=p.tau, inv_tau=p.inv_tau,
               inv_trt_tau=p.inv_trt_tau, gravitation_x=p.gravitation[0], gravitation_y=p.gravitation[1],
               gravitation_z=p.gravitation[2], u_lid=p.u_lid, d_reynolds=p.d_reynolds)
    for k, v in exp.items():
        assert got[k] == bits(v), (k, hex(got[k]), hex(bits(v)))


def test_cpp_rejects_bad_decomposition_and_missing_file(exe, tmp_path):
    p = subprocess.run([exe, "-x", "30", "-X", "4", "--dump-layout"], capture_output=True, text=True)
    assert p.returncode != 0 and "Number of subdomains does not match" in p.stderr
    p = subprocess.run([exe, "-c", str(tmp_path / "nope.xml"), "--dump-layout"], capture_output=True, text=True)
    assert p.returncode != 0 and "Loading XML file failed" in p.stderr


def test_cpp_sample_conf_parses(exe):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = run(exe, "-c", os.path.join(root, "conf.xml"), "--dump-layout")
    assert out.splitlines()[0] == "subdomain_size 256 256 256"


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("D,nums,loops", [((96, 32, 32), (3, 1, 1), 40), ((16, 16, 24), (1, 1, 2), 41),
                                          ((24, 24, 24), (2, 2, 2), 30)])
@pytest.mark.parametrize("sync", ["copy", "host", "auto"])
def test_cpp_validate_mode(exe, tmp_path, D, nums, loops, sync):
    """the reference's own acceptance test (src/main.cpp:309-408): 0 failed cells at 1e-15"""
    args = ["-c", write_conf(tmp_path / "c.xml", D, nums, loops=loops, validate=1)]
    if sync != "auto":
        args += ["--sync", sync]
    out = run(exe, *args)
    m = re.search(r"NUMBER OF FAILED CELLS/TOTAL NUMBER OF CELLS: (\d+)/(\d+)", out)
    assert m, out[-1500:]
    S = [D[a] // nums[a] - 2 for a in range(3)]
    assert int(m.group(1)) == 0 and int(m.group(2)) == S[0] * S[1] * S[2]
    assert out.count("MLUPS:") == int(np.prod(nums)) + 1


@pytest.mark.gpu
@pytest.mark.parametrize("D,nums,loops", [((96, 32, 32), (3, 1, 1), 40), ((16, 16, 24), (1, 1, 2), 41)])
def test_cpp_validate_mode_one_sided_exchange(exe, tmp_path, D, nums, loops):
    """--sync p2p with the rank threads sharing ONE GPU: push/flag/pull over device memory, and the
    z,y,x phase order the controller picks when the decomposition cuts x.  At most 4 ranks here: every
    rank owns two streams and its pull kernels spin on a flag, so the ranks of a box must fit the
    device's hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, default 8) -- which is why `auto` only
    picks p2p with one rank per GPU (8 ranks on one GPU deadlock; measured)."""
    out = run(exe, "-c", write_conf(tmp_path / "c.xml", D, nums, loops=loops, validate=1), "--sync", "p2p")
    m = re.search(r"NUMBER OF FAILED CELLS/TOTAL NUMBER OF CELLS: (\d+)/(\d+)", out)
    assert m, out[-1500:]
    S = [D[a] // nums[a] - 2 for a in range(3)]
    assert int(m.group(1)) == 0 and int(m.group(2)) == S[0] * S[1] * S[2]


@pytest.mark.gpu
@pytest.mark.parametrize("sync", ["copy", "host"])
def test_cpp_decomposed_velocity_equals_oracle(exe, tmp_path, sync):
    from helpers import bits_equal
    from oracle import multi as omulti
    D, nums, steps, L = (24, 24, 12), (2, 2, 1), 21, (0.1, 0.1, 0.1)
    dump = tmp_path / "vel.bin"
    run(exe, "-c", write_conf(tmp_path / "c.xml", D, nums, loops=steps), "-v", "--sync", sync,
        "--beta-order", "linear", "--dump-velocity", dump)
    raw = dump.read_bytes()
    make, po = omulti.make_oracle_factory(D, nums, L, dtype=np.float32, variant=1)
    md = omulti.MultiDomain(D, nums, make, slots="minimal")
    md.run(steps)
    off = 0
    for r in range(int(np.prod(nums))):
        hdr = np.frombuffer(raw, np.int32, 4, off)
        off += 16
        n = int(hdr[1]) * int(hdr[2]) * int(hdr[3])
        vel = np.frombuffer(raw, np.float32, 3 * n, off).reshape(3, hdr[3], hdr[2], hdr[1])
        off += 12 * n
        S = md.sub_size
        exp = md.ranks[r]["solver"].velocity.reshape(3, S[2], S[1], S[0])[:, 1:-1, 1:-1, 1:-1]
        assert hdr[0] == r and bits_equal(vel, exp), r


@pytest.mark.gpu
def test_cpp_known_answer_checksum(exe):
    """64^3, 100 steps, shipped (shared-memory order) beta kernel."""
    out = run(exe, "-S", 64, "-l", 100, "-v")
    m = re.search(r"Checksum: ([-\d.]+)", out)
    assert m and abs(float(m.group(1)) - 20745.41210938) < 1e-6, out[-800:]


@pytest.fixture(scope="module")
def ipc_exe(tmp_path_factory):
    """tests/cpp/ipc_ranks.cpp: one PROCESS per sub-domain (fork + socketpair carrying the 64-byte CUDA-IPC
    handles), C ABI only -- the C/C++ multi-process transport a reference maintainer with MPI would write."""
    libdir = os.path.join(ROOT, "turbulent_lbm_multigpu_b200", "lib")
    exe = str(tmp_path_factory.mktemp("ipc") / "ipc_ranks")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Wextra",
                           os.path.join(ROOT, "tests", "cpp", "ipc_ranks.cpp"), "-o", exe, "-L" + libdir, "-llbm_b200",
                           "-Wl,-rpath," + libdir, "-ldl", "-lrt", "-pthread"])
    return exe


def test_ipc_ranks_builds_and_fails_loudly_without_a_gpu(ipc_exe):
    """no CUDA device: every rank reports it, the parent reports the failed ranks -- no CPU fallback, no hang"""
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([ipc_exe, "32", "16", "16", "2", "1", "1", "4"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "no CUDA device" in p.stderr and "validation: not run (2 ranks failed)" in p.stdout


def _have_gpu():
    import ctypes
    from turbulent_lbm_multigpu_b200 import capi
    n = ctypes.c_int(0)
    return capi.load().lbmGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    "32 16 16 2 1 1 40",                  # the reference's default split (x), x,y,z phase order
    "32 16 16 2 1 1 41 zyx",              # ... x faces after the step kernel
    "16 16 24 1 1 2 40 0.1",              # z-slabs, Smagorinsky
    "768 16 16 2 1 1 20 zyx 0.1",         # rows of 384 cells: fused x exchange inside the vectorised step kernels
    "32 32 24 1 2 2 21 0.1 double",       # pencils, fp64
    "24 24 24 2 2 2 21 zyx 0.1",          # blocks: eight processes
])
def test_ipc_ranks_validate_criterion(ipc_exe, args):
    """fork()ed ranks + CUDA IPC + lbmCommStep; the reference's validate criterion (src/main.cpp:309-408):
    every rank's interior velocity block equals the single domain of size D - 2 (n - 1) bit for bit."""
    p = subprocess.run([ipc_exe] + args.split(), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-1500:])
    assert "validation: 0 failed cells" in p.stdout

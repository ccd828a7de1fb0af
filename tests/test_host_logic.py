"""Host logic without a GPU: configuration, decomposition / CComm tables (bit-exact "halo index
maps"), C ABI surface, error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import multi as omulti
from turbulent_lbm_multigpu_b200 import capi
from turbulent_lbm_multigpu_b200.configuration import CConfiguration, ConfigSingleton
from turbulent_lbm_multigpu_b200.controller import CManager, validation_domain_size, validation_sub_origin
from turbulent_lbm_multigpu_b200.domain import CComm, CDomain
from turbulent_lbm_multigpu_b200.skeleton import LBM_UNITS, SkeletonError, compute_parameters

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONF_XML = """<?xml version="1.0" encoding="ISO-8859-1"?>
<lbm-configuration>
  <physics><viscosity>0.001308</viscosity>
    <gravitation><x>0</x><y>-9.81</y><z>0</z></gravitation>
    <cavity-velocity><x>100</x><y>0</y><z>0</z><w>1</w></cavity-velocity></physics>
  <grid><domain-size><x>96</x><y>32</y><z>32</z></domain-size>
    <subdomain-num><x>3</x><y>1</y><z>1</z></subdomain-num>
    <domian-length><x>0.1</x><y>0.1</y><z>0.1</z></domian-length></grid>
  <simulation><loops>40</loops><timestep>-1.0</timestep>
    <visualization><VTK>0</VTK></visualization><validate>0</validate></simulation>
  <device><kernel-count>128</kernel-count><device-number>0</device-number></device>
</lbm-configuration>
"""


def _manager(D, nums):
    m = CManager.__new__(CManager)
    m._domain = CDomain(-1, D, (0, 0, 0), (0.1, 0.1, 0.1))
    m._controller_kw = {}
    m._lbm_controller = None
    m.setSubdomainNums(nums)
    return m


def test_conf_xml_schema(tmp_path):
    """the reference's conf.xml (same tags incl. `domian-length`), reference conf.xml:1-56"""
    f = tmp_path / "conf.xml"
    f.write_text(CONF_XML)
    c = CConfiguration(str(f))
    assert c.domain_size == (96, 32, 32) and c.subdomain_num == (3, 1, 1)
    assert c.domain_length == (0.1, 0.1, 0.1) and c.viscosity == 0.001308
    assert c.gravitation == (0.0, -9.81, 0.0) and c.drivenCavityVelocity == (100.0, 0.0, 0.0, 1.0)
    assert c.loops == 40 and c.timestep == -1.0 and not c.do_visualization and not c.do_validate
    assert c.computation_kernel_count == 128 and c.device_nr == 0 and c.smagorinsky_constant == 0.0
    with pytest.raises(RuntimeError, match="Loading XML file failed"):
        CConfiguration(str(tmp_path / "missing.xml"))
    assert ConfigSingleton.Instance() is ConfigSingleton.Instance()


def test_conf_xml_smagorinsky_extension(tmp_path):
    f = tmp_path / "conf.xml"
    f.write_text(CONF_XML.replace("<viscosity>0.001308</viscosity>",
                                  "<viscosity>0.001308</viscosity><smagorinsky-constant>0.1</smagorinsky-constant>"))
    assert CConfiguration(str(f)).smagorinsky_constant == 0.1


def test_subdomain_divisibility_error():
    """CManager::setSubdomainNums throws when the grid is not divisible (src/CManager.hpp:49-55)."""
    with pytest.raises(ValueError, match="Number of subdomains does not match with the grid size!"):
        _manager((96, 32, 32), (5, 1, 1))


@pytest.mark.parametrize("D,nums", [((96, 32, 32), (3, 1, 1)), ((24, 36, 48), (2, 3, 4)), ((64, 64, 64), (1, 1, 8)),
                                    ((32, 32, 32), (2, 2, 2))])
def test_comm_tables_bit_exact(D, nums):
    """CComm fields for every rank == the independent restatement of CManager.hpp:78-199."""
    m = _manager(D, nums)
    sub = omulti.decompose(D, nums)
    assert m.getSubdomainSize() == sub
    for r in range(nums[0] * nums[1] * nums[2]):
        rid, coords, BC, comms, origin = m.layout(r)
        ocoords, obc, ocomms, oorigin = omulti.rank_layout(r, nums, sub)
        assert (coords, BC, origin) == (ocoords, obc, oorigin)
        assert [c.as_tuple() for c in comms] == [c.as_tuple() for c in ocomms]


def test_comm_table_literal_values():
    """conf.xml default 96x32x32 / 3x1x1, middle rank (values per src/CManager.hpp:122-145)."""
    rid, coords, BC, comms, origin = _manager((96, 32, 32), (3, 1, 1)).layout(1)
    assert BC == [[8, 8], [1, 1], [1, 1]] and origin == (32, 0, 0)
    assert comms[0].as_tuple() == (0, (1, 32, 32), (1, 32, 32), (1, 0, 0), (0, 0, 0), (1, 0, 0))
    assert comms[1].as_tuple() == (2, (1, 32, 32), (1, 32, 32), (30, 0, 0), (31, 0, 0), (-1, 0, 0))


def test_ccomm_and_cdomain_accessors():
    c = CComm(3, (1, 2, 3), (1, 2, 3), (1, 0, 0), (0, 0, 0), (1, 0, 0))
    c.setDstId(5); c.setCommDirection((0, -1, 0))
    assert c.getDstId() == 5 and c.getCommDirection() == (0, -1, 0) and c.axis == 1
    d = CDomain(7, (4, 5, 6))
    assert d.getUid() == 7 and d.getSize() == (4, 5, 6) and d.getOrigin() == (0, 0, 0)
    assert d.getLength() == (0.05, 0.05, 0.05)       # default of src/CDomain.hpp:30-31


def test_validation_mapping():
    assert validation_domain_size((96, 32, 32), (3, 1, 1)) == (92, 32, 32)
    assert validation_sub_origin(2, (3, 1, 1), (30, 30, 30)) == (61, 1, 1)
    assert validation_sub_origin(5, (2, 2, 2), (10, 10, 10)) == (11, 1, 11)


def test_tau_out_of_range_is_an_error():
    with pytest.raises(SkeletonError):
        compute_parameters((8, 8, 8), (10.0, 10.0, 10.0), viscosity=1e-9)
    p = compute_parameters((8, 8, 8), (10.0, 10.0, 10.0), viscosity=1e-9, strict=False)
    assert "tau has to be within the boundary" in p.error


def test_lattice_vectors():
    assert len(LBM_UNITS) == 19 and LBM_UNITS[18] == (0, 0, 0)
    for f in range(0, 18, 2):
        assert tuple(-v for v in LBM_UNITS[f]) == LBM_UNITS[f + 1]


# ---------------------------------------------------------------- C ABI surface
def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm[A-Z]\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), "liblbm_b200.so does not export %s" % name
    assert sorted(capi.SYMBOLS) == declared, "capi.py binds a different set than the header declares"


def test_no_silent_cpu_fallback():
    """Without a CUDA device creation fails loudly; there is no CPU path behind the ABI."""
    lib = capi.load()
    n = ctypes.c_int()
    if lib.lbmGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    d = capi.lbm_desc()
    d.struct_size = ctypes.sizeof(capi.lbm_desc)
    d.size[:] = (16, 16, 16)
    d.bc[:] = (1,) * 6
    d.inv_tau = 1.5
    h = ctypes.c_void_p()
    rc = lib.lbmCreate(ctypes.byref(h), ctypes.byref(d))
    assert rc == 3 and not h.value          # LBM_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.lbmGetLastErrorString(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "turbulent_lbm_multigpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liblbm_oracle" not in src and "libref" not in src, f


def test_every_environment_switch_is_documented():
    """Each LBM_B200_* variable the library, the hosts or bench.py read appears in INTEGRATION.md section 5."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    seen = set()
    for base in ("turbulent_lbm_multigpu_b200", "bench.py"):
        path = os.path.join(ROOT, base)
        files = [path] if os.path.isfile(path) else [os.path.join(d, f) for d, _, fs in os.walk(path) for f in fs]
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(f, errors="replace").read()
                seen.update(re.findall(r'(?:getenv\(|environ(?:\.get)?[\(\[])\s*"(LBM_B200_[A-Z_]+)"', src))
    assert "LBM_B200_XFUSE" in seen and "LBM_B200_PROFILE" in seen
    missing = sorted(v for v in seen if v not in doc)
    assert not missing, missing


def test_shared_gpu_runs_ask_for_more_hardware_queues(monkeypatch):
    """capi.want_hardware_queues never overrides the user's setting and is not applied at import
    (bench.py with one rank per GPU keeps CUDA's default)."""
    monkeypatch.delenv("CUDA_DEVICE_MAX_CONNECTIONS", raising=False)
    capi.want_hardware_queues()
    assert os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] == "32"
    monkeypatch.setenv("CUDA_DEVICE_MAX_CONNECTIONS", "4")
    capi.want_hardware_queues()
    assert os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] == "4"
    src = open(os.path.join(ROOT, "turbulent_lbm_multigpu_b200", "capi.py")).read()
    assert not re.search(r"^os\.environ", src, flags=re.M)


def test_halo_slot_masks():
    lib = capi.load()
    m = ctypes.c_uint32()

    def mask(kind, d, slots):
        assert lib.lbmHaloSlotMask(kind, capi.i3(d), slots, ctypes.byref(m)) == 0
        return [f for f in range(19) if (m.value >> f) & 1]

    assert mask(capi.LBM_SYNC_BETA, (1, 0, 0), capi.LBM_HALO_SLOTS_MINIMAL) == [0, 4, 6, 8, 10]
    assert mask(capi.LBM_SYNC_BETA, (1, 0, 0), capi.LBM_HALO_SLOTS_REFERENCE) == [0, 4, 6, 8, 10]
    assert mask(capi.LBM_SYNC_ALPHA, (1, 0, 0), capi.LBM_HALO_SLOTS_MINIMAL) == [1, 5, 7, 9, 11]
    assert mask(capi.LBM_SYNC_ALPHA, (0, 0, -1), capi.LBM_HALO_SLOTS_MINIMAL) == [8, 11, 12, 15, 16]
    assert mask(capi.LBM_SYNC_ALPHA, (0, -1, 0), capi.LBM_HALO_SLOTS_REFERENCE) == list(range(19))
    for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
        expect = [f for f in range(19) if sum(a * b for a, b in zip(d, LBM_UNITS[f])) > 0]
        assert mask(capi.LBM_SYNC_BETA, d, capi.LBM_HALO_SLOTS_MINIMAL) == expect



@pytest.mark.parametrize("tname,dtype", [("float", np.float32), ("double", np.float64)])
def test_debug_dumps_follow_the_reference_format(tmp_path, tname, dtype):
    """debug_print / debugDD (reference src/CLbmSolver.hpp:981-1101): the C++ helper
    (host/CLbmDebug.hpp) and the Python twin (debug.py) print the same bytes; the layout is pinned by
    hand-written lines restating the reference's loops ("\\n<row>: " every `wrap` values, precision 4
    fixed, flags dumped byte-wise, blank line every `empty_line` values in debugDD)."""
    import subprocess
    from turbulent_lbm_multigpu_b200 import debug
    exe = str(tmp_path / "debug_print")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall",
                           os.path.join(ROOT, "tests", "cpp", "debug_print.cpp"), "-o", exe])
    got = subprocess.check_output([exe] + (["double"] if tname == "double" else []), text=True)

    cells = 12
    T = dtype
    a = np.arange(19 * cells, dtype=np.uint64)
    dd = (((a * 2654435761) & 0xffffffff) >> 9).astype(np.int64) % 4001 - 2000
    dd = dd.astype(T) / T(7919)
    a = np.arange(3 * cells, dtype=np.uint64)
    vel = ((((a * 40503) & 0xffffffff) >> 2).astype(np.int64) % 2001 - 1000).astype(T) / T(100000)
    rho = T(1) + (np.arange(cells) - 5).astype(T) / T(3000)
    fl = (1 << ((np.arange(cells) * 7) % 4)).astype(np.int32)
    fl[3] = -2
    exp = ("0.123457\n" + debug.debug_print(dd, vel, rho, fl) + "DD 5\n" + debug.debugDD(dd, cells, 5, 16, 6)
           + debug.debugDD(dd, cells, 0) + "0.123457\n")
    assert got == exp
    # the layout itself, restated by hand from the reference's loops
    lines = got.split("\n")
    assert lines[1] == "DENSITY DISTRIBUTIONS:" and lines[2].startswith("0: ") and lines[3].startswith("1: ")
    assert len(lines[2].split()) == 1 + 16 and all(len(t.split(".")[1]) == 4 for t in lines[2].split()[1:])
    i = lines.index("VELOCITY:")
    assert lines[i - 1] == "" and len(lines[i + 1].split()) == 1 + 12
    i = lines.index("DENSITY:")
    assert len(lines[i + 1].split()) == 1 + 4 and lines[i + 1].split()[1] == "%.4f" % float(rho[0])
    i = lines.index("FLAGS:")
    assert lines[i + 1] == "0: 1 0 0 0 8 0 0 0 4 0 0 0 -2 -1 -1 -1 "      # int32 flags byte by byte, signed chars
    i = lines.index("DD 5")
    # slot 5 starts at element 60: row labels count from the array start (60 // 16 = 3 appears at element 64)
    assert lines[i + 1].count(" ") == 4 and not lines[i + 1].startswith("3: ")
    assert lines[i + 2].startswith("4: ") and "" in lines[i + 1:i + 8]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header compiles as strict C99 and a C program (no C++,
    no Python, no torch) links against liblbm_b200.so and drives it.  Without a GPU creation fails
    loudly with LBM_ERR_NO_DEVICE; with one the program steps a small solver."""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "abi_from_c.c")
    exe = str(tmp_path / "abi_from_c")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", src, "-o", exe,
                           "-L" + libdir, "-llbm_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe], text=True, timeout=120)
    lines = dict(l.split(" ", 1) for l in out.strip().split("\n"))
    assert lines["version"] == "100"
    assert lines["mask"].startswith("rc=0 0x") and bin(int(lines["mask"].split("0x")[1], 16)).count("1") == 5
    if lines["create"].startswith("rc=0"):
        assert lines["stepped"] == "rc=0 counter=2" and lines["destroy"] == "rc=0"
    else:
        assert lines["create"] == "rc=3 handle=no" and "no CPU fallback" in lines["error:"]
        assert lines["null-handle"] == "step rc=%d" % 1       # LBM_ERR_INVALID

"""Profiling glue: the per-rank kernel timeline `profile_<np>_<rank>.ini`.

CPU: this repository's writers (C++ host/CProfiler.hpp and Python profiler.py) are byte-identical
to the REFERENCE's own CProfiler/CProfilerEvent (tests/golden/profile_8_5.ini, produced by
tests/golden/make_profile_golden.sh from the reference's classes), and tools/profile.py reads the
file with the reference profile.py's logic.  GPU: the C ABI records one event per kernel launch
under the reference's kernel names, in launch order, without changing the results."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from turbulent_lbm_multigpu_b200.profiler import CProfiler, CProfilerEvent, profile_file_name

GOLDEN = os.path.join(ROOT, "tests", "golden", "profile_8_5.ini")


def _events():
    """tests/cpp/profile_events.h"""
    import re
    txt = open(os.path.join(ROOT, "tests", "cpp", "profile_events.h")).read()
    return [(m.group(1), int(m.group(2)), int(m.group(3)))
            for m in re.finditer(r'\{ "(\w+)", (\d+)ull, (\d+)ull \}', txt)]


def test_python_writer_matches_reference_writer(tmp_path):
    p = CProfiler()
    for name, a, b in _events():
        p.addDeviceKernel(name, a, b)
    fn = str(tmp_path / "profile_8_5.ini")
    p.saveProfile(fn, 8, 5)
    assert open(fn, "rb").read() == open(GOLDEN, "rb").read()
    assert profile_file_name(8, 5) == "./output/profile/profile_8_5.ini"


def test_cpp_writer_matches_reference_writer(tmp_path):
    exe = str(tmp_path / "profile_writer")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "profile_writer.cpp"),
                           "-o", exe])
    fn = str(tmp_path / "out.ini")
    out = subprocess.check_output([exe, fn], text=True).split()
    assert open(fn, "rb").read() == open(GOLDEN, "rb").read()
    assert out == ["7", "0"]


def test_event_rejects_empty_name_and_overlap_rule():
    with pytest.raises(ValueError):
        CProfilerEvent(1, "", 0, 1)
    a, b, c = CProfilerEvent(1, "a", 0, 10), CProfilerEvent(2, "b", 9, 20), CProfilerEvent(3, "c", 10, 12)
    assert a.overlap(b) and b.overlap(a) and not a.overlap(c) and b.overlap(c)


def test_analyser_reads_the_reference_schema(tmp_path):
    import profile as lbm_profile           # tools/profile.py
    total, proc, events = lbm_profile.read_profile(GOLDEN)
    assert (total, proc, len(events)) == (8, 5, 7)
    assert [e.getEventId() for e in events][:3] == ["init_kernel", "copy_buffer_rect", "lbm_kernel_beta"]
    assert [(e.getEventStartTime(), e.getEventEndTime()) for e in events] == [(a, b) for _, a, b in _events()]
    assert lbm_profile.overlapping(events) == []
    # a file with an overlapped exchange: halo kernels under the interior step kernel
    p = CProfiler()
    p.addDeviceKernel("lbm_kernel_beta", 0, 1000)          # shell
    p.addDeviceKernel("lbm_kernel_beta", 1000, 9000)       # interior
    p.addDeviceKernel("halo_push", 1100, 1600)
    p.addDeviceKernel("halo_pull", 1600, 2100)
    p.addDeviceKernel("halo_pull", 9000, 9500)             # exposed
    fn = str(tmp_path / "output" / "profile" / "profile_2_0.ini")
    p.saveProfile(fn, 2, 0)
    p.saveProfile(fn, 2, 0)                                # appended twice (ios::app): last run wins
    r = lbm_profile.analyse([fn])[0]
    assert r["events"] == 5 and r["overlapping_events"] == 2
    assert r["kernels"]["lbm_kernel_beta"]["count"] == 2
    assert abs(r["halo_hidden_frac"] - 1000 / 1500) < 1e-12
    assert lbm_profile.main(["--dir", os.path.dirname(fn)]) == 0


@pytest.mark.gpu
def test_timeline_has_one_event_per_launch_and_leaves_results_unchanged():
    from helpers import bits_equal, make_cuda
    size = (40, 24, 16)
    a = make_cuda(size, np.float32)
    b = make_cuda(size, np.float32)
    b.profileEnable(3)                                     # events + NVTX ranges
    l0 = b.launchCount()
    for s in (a, b):
        for _ in range(6):
            s.simulationStep()
    ev = b.profileEvents()
    assert len(ev) == b.launchCount() - l0
    names = [n for n, _, _ in ev]
    assert names.count("lbm_kernel_alpha") == 3 and names.count("lbm_kernel_beta") == 3
    assert names.count("lbm_kernel_beta.wrap") == 3 and set(names) <= {"lbm_kernel_alpha", "lbm_kernel_beta", "lbm_kernel_beta.wrap"}
    assert all(t1 >= t0 for _, t0, t1 in ev)
    # launch order on the compute stream is time order
    main = [(t0, t1) for n, t0, t1 in ev if n != "lbm_kernel_beta.wrap"]
    assert all(main[i][1] <= main[i + 1][0] + 2000 for i in range(len(main) - 1))     # 2 us event granularity
    assert sum(t1 - t0 for t0, t1 in main) > 0
    assert bits_equal(a.storeDensityDistribution(), b.storeDensityDistribution())
    names2 = [n for n, _, _ in b.profileEvents()]
    assert names2.count("copy_buffer_rect") == 0           # whole-array store is a memcpy, no kernel
    b.profileClear()
    assert b.profileEventCount() == (0, 0)
    b.storeFlags(origin=(1, 1, 1), size=(4, 4, 4))
    assert [n for n, _, _ in b.profileEvents()] == ["copy_buffer_rect"]
    b.profileEnable(0)
    b.simulationStep()
    assert b.profileEventCount()[0] == 1
    a.close(); b.close()


@pytest.mark.gpu
def test_driver_writes_profile_ini_per_rank(tmp_path):
    import profile as lbm_profile
    from turbulent_lbm_multigpu_b200.host import build as host_build
    exe = host_build.build()
    env = dict(os.environ, LBM_B200_PROFILE="1")
    p = subprocess.run([exe, "-x", "32", "-y", "32", "-z", "48", "-Z", "2", "-l", "10"], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    files = sorted(os.listdir(tmp_path / "output" / "profile"))
    assert files == ["profile_2_0.ini", "profile_2_1.ini"]
    rep = lbm_profile.analyse([str(tmp_path / "output" / "profile" / f) for f in files])
    for rank, r in enumerate(rep):
        assert (r["total_num_proc"], r["current_proc_id"]) == (2, rank)
        k = r["kernels"]
        assert k["init_kernel"]["count"] == 1
        assert k["lbm_kernel_alpha"]["count"] + k["lbm_kernel_beta"]["count"] >= 10
        assert r["events"] == sum(c["count"] for c in k.values())

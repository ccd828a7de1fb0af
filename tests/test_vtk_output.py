"""Output path: the legacy-VTK writer of the C++ host facade is byte-identical
to the reference's own writer (src/libvis/CLbmVisualizationVTK.hpp + VTK_Common.cpp).

tests/golden/vtk_{f32,f64}.3.7.vtk were produced by the REFERENCE writer fed with
tests/cpp/vtk_mock_solver.hpp (tests/golden/make_vtk_golden.sh, run where /root/reference is
mounted); here this repository's writer is compiled against the same mock and compared."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def writer(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("vtk") / "vtk_writer")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-Wall", os.path.join(ROOT, "tests", "cpp", "vtk_writer.cpp"),
                           "-o", exe])
    return exe


@pytest.mark.parametrize("tag,extra", [("f32", []), ("f64", ["double"])])
def test_vtk_writer_matches_reference_writer(writer, tmp_path, tag, extra):
    prefix = str(tmp_path / "mine")
    subprocess.check_call([writer, prefix] + extra)
    got = open(prefix + ".3.7.vtk", "rb").read()
    exp = open(os.path.join(ROOT, "tests", "golden", "vtk_%s.3.7.vtk" % tag), "rb").read()
    assert got == exp


@pytest.mark.gpu
def test_driver_writes_one_vtk_file_per_rank_and_step(tmp_path):
    from turbulent_lbm_multigpu_b200.host import build as host_build
    exe = host_build.build()
    p = subprocess.run([exe, "-x", "16", "-y", "16", "-z", "24", "-Z", "2", "-l", "3", "-g"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    files = sorted(os.listdir(tmp_path / "output" / "vtk"))
    assert files == ["OUTPUT.%d.%d.vtk" % (r, s) for r in range(2) for s in range(3)]
    txt = open(tmp_path / "output" / "vtk" / "OUTPUT.1.2.vtk").read()
    assert "DIMENSIONS  17 17 13 " in txt and "CELL_DATA %d " % (16 * 16 * 12) in txt
    # rank 1 sits at z origin 12 cells: first grid point z = 12 * cell length (0.1/2 / 12... = L_sub/Sz)
    first = re.search(r"POINTS \d+ float\n\n(\S+) (\S+) (\S+)\n", txt)
    assert first and float(first.group(1)) == 0.0 and float(first.group(3)) > 0.0
    flags = txt.split("SCALARS flag INT 1 \nLOOKUP_TABLE default \n")[1].split("\n\n")[0].split()
    assert set(flags) <= {"1", "2", "4", "8"} and "8" in flags and "4" in flags

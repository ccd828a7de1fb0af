"""tools/benchmark.py: the recipes generate the reference's command lines (benchmark.py:61-129)
and the .ini averaging reads what CController::run appends (src/CController.hpp:445-477)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("lbm_benchmark", os.path.join(ROOT, "tools", "benchmark.py"))
bm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bm)


def _args(argv):
    return dict(zip(argv[1::2], argv[2::2]))


def test_reference_recipes():
    n, argv = bm.point("weak-1d", 4, reference_recipe=True)
    a = _args(argv)
    assert n == 4 and (a["-x"], a["-y"], a["-z"], a["-X"], a["-Y"], a["-Z"]) == ("4096", "1024", "32", "4", "1", "1")
    assert float(a["-n"]) == 0.4 and float(a["-m"]) == 0.1
    n, argv = bm.point("strong-1d", 3, reference_recipe=True)
    a = _args(argv)
    assert n == 8 and (a["-x"], a["-y"], a["-X"]) == ("1024", "1024", "8") and float(a["-n"]) == 0.1
    n, argv = bm.point("weak-2d", 2, reference_recipe=True, grid=1024)
    a = _args(argv)
    assert n == 4 and (a["-x"], a["-y"], a["-X"], a["-Y"]) == ("2048", "2048", "2", "2")


def test_z_slab_recipes():
    n, argv = bm.point("weak-1d", 8, axis="z", grid=256)
    a = _args(argv)
    assert n == 8 and (a["-x"], a["-y"], a["-z"], a["-Z"]) == ("256", "256", "2048", "8") and abs(float(a["-p"]) - 0.8) < 1e-12
    n, argv = bm.point("strong-2d", 2, axis="z", grid=512)
    a = _args(argv)
    assert n == 4 and (a["-Y"], a["-Z"], a["-X"]) == ("2", "2", "1")


def test_ini_averaging(tmp_path):
    for nproc, mlups in ((1, (100.0, 110.0)), (2, (200.0, 190.0))):
        with open(tmp_path / ("benchmark_%d.ini" % nproc), "w") as f:
            for i, m in enumerate(mlups, 1):
                f.write("[EXP%d]\nNP : %d\nCUBE_X : 64\nCUBE_Y : 64\nCUBE_Z : 64\nSECONDS : 1.5\nFPS : 66\n"
                        "MLUPS : %g\nBANDWIDTH : 1\n\n" % (i, nproc, m))
    res = bm.analyse(sorted(str(p) for p in tmp_path.glob("*.ini")), peak_gbs=1000.0)
    assert res[1]["MLUPS"] == 105.0 and res[2]["MLUPS"] == 195.0 and res[2]["NUM_EXP"] == 2
    assert abs(res[2]["SPEEDUP"] - 195.0 / 105.0) < 1e-12 and abs(res[2]["EFFICIENCY"] - 195.0 / 210.0) < 1e-12
    assert abs(res[1]["ROOFLINE_FRAC_PER_GPU"] - 105.0 * 156.0 / 1e3 / 1000.0) < 1e-12


def test_visualize_writes_results_txt_and_plots(tmp_path):
    """reference benchmark.py:131-185: pretty-printed results and one plot per key plus the speed-up."""
    res = {1: {"SECONDS": 2.0, "MLUPS": 100.0, "FPS": 5.0}, 2: {"SECONDS": 1.0, "MLUPS": 200.0, "FPS": 10.0},
           4: {"SECONDS": 0.5, "MLUPS": 390.0, "FPS": 20.0}}
    files = bm.visualize(res, str(tmp_path))
    names = sorted(os.path.basename(f) for f in files)
    assert names == ["plot_FPS.svg", "plot_MLUPS.svg", "plot_SECONDS.svg", "plot_speedup.svg"]
    txt = (tmp_path / "results.txt").read_text()
    assert "profiling results pretty print" in txt and "390.0" in txt
    svg = (tmp_path / "plot_speedup.svg").read_text()
    assert svg.startswith("<svg") and "Speedup Scaling" in svg and svg.count("<path") == 3


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: ONE JSON line on stdout (libraries that print there are redirected to
    stderr), the reference arm's extra keys, and -- under torchrun with N > 1 -- ranks other than 0
    exit 0 without work or output.  The reference arm runs on the host cores, so this is a CPU test."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--size", "32"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-1500:]
    lines = p.stdout.strip().split("\n")
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "MLUPS" and j["unit"] == "MLUPS" and j["higher_is_better"] is True
    assert j["steps"] == 1 and j["value"] > 0 and j["n_gpus"] == 1 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout == ""

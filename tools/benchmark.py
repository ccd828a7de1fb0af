#!/usr/bin/env python3
"""Scaling-benchmark harness: py3 successor of the reference's benchmark.py.

Same recipes (reference benchmark.py:40-129): weak/strong scaling in 1-D and 2-D, `num_exp`
repetitions per point, every run appends to output/benchmark/benchmark_<np>.ini, which is
averaged afterwards (reference :11-34) into results.json -- extended with speed-up, parallel
efficiency and the HBM-roofline fraction.  `mpirun -n N ./lbm_opencl ...` becomes ONE
`lbm_b200 ...` process (ranks are host threads, one GPU each).

    python tools/benchmark.py weak-1d --max-num 8 --axis z          # 256x256x(256 n), (1,1,n)
    python tools/benchmark.py strong-1d --max-num 3 --grid 512       # np = 1,2,4,8
    python tools/benchmark.py weak-1d --reference-recipe             # x = 1024 n, y = 1024, z = 32, -X n
    python tools/benchmark.py analyse
"""
from __future__ import annotations

import argparse
import configparser
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LBM_COMMAND = os.path.join(ROOT, "turbulent_lbm_multigpu_b200", "host", "lbm_b200")
INI_FILE_DIR = os.path.join("output", "benchmark")
SECTION_BASE = "EXP"
KEYS = ["CUBE_X", "CUBE_Y", "CUBE_Z", "SECONDS", "FPS", "MLUPS", "BANDWIDTH"]
AXES = "xyz"


def point(kind, num_increase, axis="z", grid=256, loops=100, base_length=0.1, reference_recipe=False, extra=()):
    """(num_proc, argv) of one benchmark point."""
    size, nums, length = [grid] * 3, [1, 1, 1], [base_length] * 3
    a = AXES.index(axis)
    b = (a + 2) % 3 if axis == "z" else (a + 1) % 3        # second axis of the 2-D recipes
    if reference_recipe:                                    # benchmark.py:61-83,107-129
        a, b = 0, 1
        size = [1024, 1024, 32] if kind.endswith("1d") else [grid, grid, 32]
    if kind == "weak-1d":
        nproc = num_increase
        size[a] *= nproc; nums[a] = nproc; length[a] *= nproc
    elif kind == "weak-2d":
        nproc = num_increase * num_increase
        for ax in (a, b):
            size[ax] *= num_increase; nums[ax] = num_increase; length[ax] *= num_increase
    elif kind == "strong-1d":
        nproc = 2 ** num_increase
        nums[a] = nproc
    elif kind == "strong-2d":
        nproc = num_increase * num_increase
        nums[a] = nums[b] = num_increase
    else:
        raise ValueError(kind)
    argv = [LBM_COMMAND, "-x", size[0], "-y", size[1], "-z", size[2], "-X", nums[0], "-Y", nums[1], "-Z", nums[2],
            "-l", loops, "-n", length[0], "-m", length[1], "-p", length[2]] + list(extra)
    return nproc, [str(v) for v in argv]


def run_recipe(kind, max_num, num_exp, dry_run=False, **kw):
    os.makedirs(INI_FILE_DIR, exist_ok=True)
    first = 0 if kind == "strong-1d" else 1
    commands = []
    for k in range(first, max_num + 1):
        nproc, argv = point(kind, k, **kw)
        ini = os.path.join(INI_FILE_DIR, "benchmark_%d.ini" % nproc)
        for exp in range(1, num_exp + 1):
            commands.append(argv)
            if dry_run:
                continue
            with open(ini, "a") as f:                      # section header, then the run appends its keys
                f.write("[%s%d]\nNP : %d\n" % (SECTION_BASE, exp, nproc))
            print("executing command:", " ".join(argv), flush=True)
            subprocess.run(argv, env=dict(os.environ, LBM_B200_BENCHMARK="1"), check=False)
    return commands


def analyse(filenames, bytes_per_lup=156.0, peak_gbs=None):
    """average every key over the experiments of each file; add speed-up / efficiency / roofline"""
    res = {}
    for fn in filenames:
        cfg = configparser.ConfigParser(delimiters=(":",))
        cfg.read(fn)
        secs = [s for s in cfg.sections() if s.startswith(SECTION_BASE)]
        if not secs:
            continue
        nproc = cfg.getint(secs[0], "NP")
        res[nproc] = {k: sum(cfg.getfloat(s, k) for s in secs) / len(secs) for k in KEYS if cfg.has_option(secs[0], k)}
        res[nproc]["NUM_EXP"] = len(secs)
    if 1 in res:
        for nproc, r in res.items():
            r["SPEEDUP"] = r["MLUPS"] / res[1]["MLUPS"]
            r["EFFICIENCY"] = r["SPEEDUP"] / nproc
    if peak_gbs:
        for nproc, r in res.items():
            r["ROOFLINE_FRAC_PER_GPU"] = r["MLUPS"] * bytes_per_lup / 1e3 / nproc / peak_gbs
    return res


def _svg_plot(path, title, xlabel, ylabel, xs, ys):
    """One line plot as a dependency-free SVG (matplotlib is not part of the B200 image)."""
    W, H, L, R, T, B = 640, 420, 80, 20, 40, 50
    x0, x1 = min(xs) - 1, max(xs) + 1
    y0, y1 = min(0.0, min(ys)), max(ys) * 1.08 if max(ys) > 0 else 1.0

    def px(x):
        return L + (x - x0) / (x1 - x0) * (W - L - R)

    def py(y):
        return H - B - (y - y0) / (y1 - y0) * (H - T - B)
    out = ['<svg xmlns="http://www.w3.org/2000/svg" width="%d" height="%d" font-family="sans-serif" font-size="12">' % (W, H),
           '<rect width="100%" height="100%" fill="white"/>',
           '<text x="%d" y="22" text-anchor="middle" font-size="15">%s</text>' % (W // 2, title),
           '<text x="%d" y="%d" text-anchor="middle">%s</text>' % (W // 2, H - 10, xlabel),
           '<text x="16" y="%d" text-anchor="middle" transform="rotate(-90 16 %d)">%s</text>' % (H // 2, H // 2, ylabel)]
    for k in range(6):
        y = y0 + (y1 - y0) * k / 5
        out.append('<line x1="%d" y1="%.1f" x2="%d" y2="%.1f" stroke="#ccc"/>' % (L, py(y), W - R, py(y)))
        out.append('<text x="%d" y="%.1f" text-anchor="end">%.4g</text>' % (L - 6, py(y) + 4, y))
    for x in xs:
        out.append('<line x1="%.1f" y1="%d" x2="%.1f" y2="%d" stroke="#ccc"/>' % (px(x), T, px(x), H - B))
        out.append('<text x="%.1f" y="%d" text-anchor="middle">%d</text>' % (px(x), H - B + 16, x))
    pts = " ".join("%.1f,%.1f" % (px(x), py(y)) for x, y in zip(xs, ys))
    out.append('<polyline points="%s" fill="none" stroke="green" stroke-dasharray="6,4" stroke-width="2"/>' % pts)
    for x, y in zip(xs, ys):
        out.append('<path d="M %.1f %.1f l 6 10 l -12 0 z" fill="green"/>' % (px(x), py(y) - 6))
    out.append("</svg>")
    with open(path, "w") as f:
        f.write("\n".join(out))


def visualize(res, outdir=INI_FILE_DIR):
    """reference benchmark.py:131-185: pretty-print the averaged results (-> results.txt), one
    '<KEY> scaling' plot per key over the number of GPUs and a speed-up plot (SECONDS[first] / SECONDS)."""
    import pprint
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "results.txt"), "w") as f:
        f.write("profiling results pretty print:\n" + pprint.pformat(res, indent=4) + "\n")
    ngpus = sorted(res)
    if not ngpus:
        return []
    files = []
    for key in KEYS:
        if not all(key in res[n] for n in ngpus):
            continue
        fn = os.path.join(outdir, "plot_%s.svg" % key)
        _svg_plot(fn, key + " scaling", "# GPUs", key, ngpus, [float(res[n][key]) for n in ngpus])
        files.append(fn)
    if all("SECONDS" in res[n] for n in ngpus):
        t = [float(res[n]["SECONDS"]) for n in ngpus]
        fn = os.path.join(outdir, "plot_speedup.svg")
        _svg_plot(fn, "Speedup Scaling", "# GPUs", "speedup", ngpus, [t[0] / v for v in t])
        files.append(fn)
    return files


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return None


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("recipe", choices=["weak-1d", "weak-2d", "strong-1d", "strong-2d", "analyse"])
    ap.add_argument("--max-num", type=int, default=1)
    ap.add_argument("--num-exp", type=int, default=1)
    ap.add_argument("--loops", type=int, default=100)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--axis", default="z", choices=list(AXES))
    ap.add_argument("--reference-recipe", action="store_true", help="the reference's sizes and x/y split")
    ap.add_argument("--smagorinsky", type=float, default=0.0)
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args(argv)
    if a.recipe != "analyse":
        extra = (["--smagorinsky", a.smagorinsky] if a.smagorinsky else []) + (["--double"] if a.double else [])
        cmds = run_recipe(a.recipe, a.max_num, a.num_exp, dry_run=a.dry_run, axis=a.axis, grid=a.grid, loops=a.loops,
                          reference_recipe=a.reference_recipe, extra=extra)
        if a.dry_run:
            for c in cmds:
                print(" ".join(c))
            return 0
    res = analyse(sorted(glob.glob(os.path.join(INI_FILE_DIR, "benchmark_*.ini"))),
                  bytes_per_lup=308.0 if a.double else 156.0, peak_gbs=measured_peak())
    with open("results.json", "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    for fn in visualize(res):
        print("saved graph:", fn)
    for nproc in sorted(res):
        r = res[nproc]
        print("np %2d  MLUPS %10.1f  speed-up %5.2f  efficiency %5.1f %%  roofline/GPU %s" % (
            nproc, r["MLUPS"], r.get("SPEEDUP", float("nan")), 100 * r.get("EFFICIENCY", float("nan")),
            ("%.1f %%" % (100 * r["ROOFLINE_FRAC_PER_GPU"])) if "ROOFLINE_FRAC_PER_GPU" in r else "-"))
    return 0


if __name__ == "__main__":
    sys.exit(main())

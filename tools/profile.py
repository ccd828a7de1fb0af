#!/usr/bin/env python3
"""py3 successor of the reference's profile.py (reference profile.py:1-139): reads the per-rank
timelines `output/profile/profile_<np>_<rank>.ini` that a run with LBM_B200_PROFILE=1 writes
(lbm_b200 driver or the Python CController), reports the number of events and of OVERLAPPING
events per rank like the reference does (profile.py:36-77), and adds what the reference left as a
TODO: time per kernel name and how much of the halo kernels' time is hidden under a step kernel.

    LBM_B200_PROFILE=1 turbulent_lbm_multigpu_b200/host/lbm_b200 -x 256 -y 256 -z 512 -Z 2 -l 100
    python tools/profile.py [--dir output/profile] [--json]
"""
from __future__ import annotations

import argparse
import configparser
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from turbulent_lbm_multigpu_b200.profiler import CProfilerEvent  # noqa: E402

METADATA_SECTION = "METADATA"
STEP_KERNELS = ("lbm_kernel_alpha", "lbm_kernel_beta")
HALO_KERNELS = ("halo_push", "halo_pull", "copy_buffer_rect")


def read_profile(filename):
    """-> (total_num_proc, current_proc_id, [CProfilerEvent]); sections EVENT1..EVENTn as
    profile.py:36-47 reads them.  A file appended to by several runs (the writers open with
    ios::app like the reference) keeps the last run."""
    text = open(filename).read()
    last = text.rfind("[METADATA]")
    config = configparser.ConfigParser(delimiters=(":",), comment_prefixes=("#",))
    config.read_string(text[last:] if last >= 0 else text)
    total = config.getint(METADATA_SECTION, "TOTAL_NUM_PROC")
    proc = config.getint(METADATA_SECTION, "CURRENT_PROC_ID")
    events = []
    for k in range(1, len(config.sections())):
        sec = "EVENT%d" % k
        ev = CProfilerEvent(k, config.get(sec, "NAME").strip(), config.getint(sec, "START"), config.getint(sec, "END"),
                            1 if config.get(sec, "TYPE").strip() == "DEVICE_KERNEL" else 2)
        ev.file_duration = config.getfloat(sec, "DURATION")
        events.append(ev)
    return total, proc, events


def overlapping(events):
    """pairs of events whose [start, end) intervals intersect (profile.py:49-58), found by a sweep
    over start-sorted events instead of the reference's all-pairs loop"""
    ev = sorted(events, key=lambda e: (e.getEventStartTime(), e.getEventEndTime()))
    out = []
    for i, a in enumerate(ev):
        for b in ev[i + 1:]:
            if b.getEventStartTime() >= a.getEventEndTime():
                break
            if a.overlap(b):
                out.append((a, b))
    return out


def hidden_ns(events):
    """ns of halo-kernel time that runs while a step kernel of the same rank is running"""
    steps = sorted((e.getEventStartTime(), e.getEventEndTime()) for e in events if e.getEventId() in STEP_KERNELS)
    total = hidden = 0
    for e in events:
        if e.getEventId() not in HALO_KERNELS:
            continue
        s, t = e.getEventStartTime(), e.getEventEndTime()
        total += t - s
        for a, b in steps:
            if a >= t:
                break
            hidden += max(0, min(t, b) - max(s, a))
    return total, hidden


def analyse(filenames):
    report = []
    for fn in sorted(filenames):
        total, proc, events = read_profile(fn)
        per = {}
        for e in events:
            c = per.setdefault(e.getEventId(), dict(count=0, ms=0.0))
            c["count"] += 1
            c["ms"] += e.getEventDuration()
        halo_total, halo_hidden = hidden_ns(events)
        span = (max(e.getEventEndTime() for e in events) - min(e.getEventStartTime() for e in events)) if events else 0
        report.append(dict(file=fn, total_num_proc=total, current_proc_id=proc, events=len(events),
                           overlapping_events=len(overlapping(events)), kernels=per, span_ms=span / 1e6,
                           halo_ms=halo_total / 1e6,
                           halo_hidden_frac=(halo_hidden / halo_total) if halo_total else None))
    return report


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="./output/profile")
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args(argv)
    filenames = glob.glob(os.path.join(args.dir, "*.ini"))
    report = analyse(filenames)
    if args.json:
        print(json.dumps(report))
        return 0
    print("start analysing ", len(filenames), " files.")
    for r in report:
        print("\nanalysing file: ", r["file"])
        print("TOTAL NUMBER OF PROCESSES: ", r["total_num_proc"])
        print("CURRENT PROCCESSOR ID: ", r["current_proc_id"])
        print("# EVENTS: ", r["events"])
        print("# OEVERLAPPING EVENTS FOUND: ", r["overlapping_events"])
        for name, c in sorted(r["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
            print("  %-22s %7d launches %12.3f ms  %9.4f ms each" % (name, c["count"], c["ms"], c["ms"] / c["count"]))
        if r["halo_hidden_frac"] is not None:
            print("  halo kernels: %.3f ms, %.1f %% of it under a step kernel" % (r["halo_ms"], 100 * r["halo_hidden_frac"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())

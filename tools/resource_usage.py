#!/usr/bin/env python3
"""Registers / stack / shared memory of every kernel in liblbm_b200.so (cuobjdump
--dump-resource-usage), as a markdown table: the CPU-side check before GPU time is spent.

usage: python tools/resource_usage.py [lib.so] [name-filter-regex]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "turbulent_lbm_multigpu_b200", "lib", "liblbm_b200.so")
    flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
    rows = []
    for m in re.finditer(r"Function (\S+):\n\s*(.*)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void lbm::", "").replace("void ", "")
        if flt and not flt.search(name):
            continue
        r = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
        rows.append((name, int(r["REG"]), int(r["STACK"]), int(r["SHARED"])))
    rows.sort()
    print("| kernel | registers | stack B | shared B | resident 128-thread blocks/SM (register limit) |")
    print("|---|---|---|---|---|")
    for name, reg, stack, shared in rows:
        alloc = (reg + 7) // 8 * 8
        print("| `%s` | %d | %d | %d | %d |" % (name, reg, stack, shared, min(16, 65536 // (alloc * 128))))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Print one table row per bench.py JSON line found in the given files (tuning/reporting aid)."""
import json
import sys

print("| file | N | workload | MLUPS | ms/step | roofline frac per GPU | halo exposed | e2e MLUPS | parity | graph |")
print("|---|---|---|---|---|---|---|---|---|---|")
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print("| %s | unreadable: %s |" % (f, e))
        continue
    if j.get("impl") == "reference":
        continue
    h = j.get("halo") or {}
    c = j["config"]
    par = j.get("parity")
    print("| %s | %d | %s %s %s | %.0f | %.4f | %.3f | %s | %.0f | %s | %s |" % (
        f.split("/")[-1], j["n_gpus"], c["workload"].split(" D3Q19")[0].replace("lid-driven cavity ", ""),
        "x".join(str(v) for v in c["subdomain_num"]), j["dtype"], j["value"], j["ms_per_step"], j["roofline"]["frac"],
        ("%.1f %%" % (100 * h["exposed_frac"])) if h else "-", j["e2e"]["value"],
        "-" if not par else ("ok" if par.get("ok") else "FAIL" if par.get("ok") is False else "n/a"),
        "yes" if j.get("cuda_graph") else "no"))

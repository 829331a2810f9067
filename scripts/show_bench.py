#!/usr/bin/env python
"""Print the essentials of bench.py JSON lines: python scripts/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        r = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "UNREADABLE", e)
        continue
    if "unavailable" in r:
        print(path, "unavailable:", r["unavailable"])
        continue
    line = "%s: n_gpus %s value %.0f %s, %.4f ms/step" % (path, r.get("n_gpus"), r["value"], r.get("unit", ""), r.get("ms_per_step", 0.0))
    if "e2e" in r:
        line += ", e2e %.0f" % r["e2e"]["value"]
    if "reps" in r and isinstance(r["reps"], dict):
        line += ", reps %.4f-%.4f" % (r["reps"]["ms_per_step_min"], r["reps"]["ms_per_step_max"])
    print(line)
    if "kernel_us" in r:
        print("   kernel_us", {k: round(v, 1) for k, v in r["kernel_us"].items()})
        print("   roofline", {k: (round(v["frac"], 3), round(v["frac_l2"], 3)) for k, v in r.get("roofline_all", {}).items()},
              "l2_copy_gbs %.0f" % r["roofline"].get("l2_copy_gbs", 0), "clocks", r.get("clocks"))
    for k in ("ensemble_c4", "banded"):
        if k in r:
            print("   %s" % k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in r[k].items() if a not in ("workload", "collectives", "initial_field")})
    if "warning" in r:
        print("   WARNING", r["warning"])

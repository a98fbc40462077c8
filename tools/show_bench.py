#!/usr/bin/env python
"""Print the headline numbers and the per-call breakdown of a bench.py JSON line (development helper)."""
import json, sys
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(path, "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "launches", d.get("gpu_launches"),
          "n_gpus", d["n_gpus"])
    r = d.get("roofline") or {}
    print("  roofline", r.get("kernel"), "frac", round(r.get("frac", 0), 3), "share", r.get("share_of_step"))
    for k, v in list((d.get("kernels") or {}).items())[:int(sys.argv[0] and 16)]:
        print("   ", k.ljust(28), v)

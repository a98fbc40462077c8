#!/usr/bin/env python
"""Per-launch DRAM traffic of ONE HeteroConv layer forward + backward from an ncu metrics pass:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
      --log-file gpurun_out/r2_layer_traffic_c4s8.csv python tools/layer_bench.py --workload C4s8 --reps 1 --only-layer
  python tools/layer_traffic.py gpurun_out/r2_layer_traffic_c4s8.csv C4s8 > profiles/r2_traffic.json

layer_bench --only-layer --reps 1 runs the layer three times (two warm-ups + one timed); the launches of the last run are
the ones after the last ATen fill of the L2-flush buffer (`vectorized_elementwise_kernel ... FillFunctor<unsigned char>`)."""
import csv
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path, workload = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    order = []
    for r in rows[1:]:
        if not r[ix["ID"]].isdigit():
            continue
        k = int(r[ix["ID"]])
        if k not in per:
            per[k] = {"name": r[ix["Kernel Name"]]}
            order.append(k)
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
        per[k][r[ix["Metric Name"]]] = v * mult
    last_fill = max(i for i, k in enumerate(order) if "FillFunctor<unsigned char>" in per[k]["name"])
    launches = [per[k] for k in order[last_fill + 1:] if "sleep" not in per[k]["name"].lower() and "spin" not in per[k]["name"].lower()]
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches)
    spec = importlib.import_module("multi-modal-gnn_b200.synth").SPECS[workload]
    bytes_min = 5 * spec.n_patient * 128 * 4 + 2 * (spec.e_lab + spec.e_dx + spec.e_med) * 4 + 6 * (spec.n_patient + 1) * 4
    top = sorted(launches, key=lambda l: -l.get("gpu__time_duration.sum", 0.0))[:6]
    out = {"hetero_layer_fwd_bwd:" + workload: {
        "dram_bytes_per_layer": rd + wr, "dram_read": rd, "dram_write": wr, "n_kernels": len(launches), "bytes_min": bytes_min,
        "ratio_to_bytes_min": (rd + wr) / bytes_min,
        "top_kernels_us_readMB_writeMB": [[l["name"].split("(")[0][-40:], round(l.get("gpu__time_duration.sum", 0.0), 1),
                                           round(l.get("dram__bytes_read.sum", 0.0) / 1e6, 1), round(l.get("dram__bytes_write.sum", 0.0) / 1e6, 1)] for l in top],
        "source": "profiles/" + os.path.basename(path) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  "--clock-control none over python tools/layer_bench.py --workload " + workload + " --reps 1 --only-layer; the launches of the "
                  "last layer forward+backward; tools/layer_traffic.py)"}}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

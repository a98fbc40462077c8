#!/usr/bin/env python
"""Device timeline of one CUDA-graph replay of the training step (development helper, 1 GPU):
per-kernel busy time and the idle gaps between consecutive kernels, from torch.profiler (CUPTI) records.

    python tools/graph_timeline.py [workload] > gpurun_out/timeline.txt
"""
import importlib
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multi-modal-gnn_b200"


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "C2"
    import bench
    pkg = importlib.import_module(PKG)
    M, T = importlib.import_module(PKG + ".model"), importlib.import_module(PKG + ".trainer")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    spec = pkg.synth.SPECS[workload]
    cfg = bench._cfg(dropout=0.2, loss="mse")
    g = pkg.synth.make_graph(spec, seed=42)
    masker = T.EdgeMasker(g, 0.7, 0.15, 0.15, 0.2, 42)
    torch.manual_seed(0)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None)
    trainer = T.Trainer(model, g, masker, cfg, dev)
    model._init_embeddings(trainer.data)
    pi, li = masker.split_rows("train")
    _, ev = masker.split_edges("train")
    sup = masker.supervision_mask("train", seed=1).to(dev)
    trainer.enable_cuda_graph()
    model.train()
    for _ in range(4):
        trainer.train_step(pi, li, ev, sup)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            trainer.train_step(pi, li, ev, sup)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    print(f"{len(evs)} device events over 3 steps")
    t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"span {(t1 - t0) / 3:.1f} us/step, busy {busy / 3:.1f} us/step, idle {(t1 - t0 - busy) / 3:.1f} us/step")
    agg = defaultdict(lambda: [0, 0.0, 0.0])          # name -> count, busy, gap after
    for a, b in zip(evs, evs[1:] + [None]):
        n = a.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
        agg[n][0] += 1
        agg[n][1] += a.time_range.end - a.time_range.start
        if b is not None:
            agg[n][2] += max(0.0, b.time_range.start - a.time_range.end)
    print(f"{'kernel':70s} {'n/step':>7s} {'busy us/step':>13s} {'us/launch':>10s} {'gap-after us/step':>18s}")
    for n, (c, bu, ga) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print(f"{n:70s} {c / 3:7.1f} {bu / 3:13.1f} {bu / c:10.2f} {ga / 3:18.1f}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Skip experiments / wait-cycle counters of k_layer_tf32 (B2G_LAYER_DBG bits: 1 skip the bit expansion, 2 skip the B loads,
4 skip the x loads, 8 print per-role wait cycles of CTA 0; B2G_LAYER_STAGES) at the C4 shard, TF32 and fp16 adjacency tiles.

  python tools/layer_probe.py [--workload C4s8]
"""
import argparse
import importlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multi-modal-gnn_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4s8")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    M = importlib.import_module(PKG + ".model")
    dev = torch.device("cuda:0")
    spec = pkg.synth.SPECS[args.workload]
    g = pkg.synth.make_graph(spec, seed=42, device=dev).to(dev)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2, "use_batch_norm": True, "activation": "relu"}}
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    model._init_embeddings(g)
    gi = model._graph_index(g)
    pb = gi.hub_bits("patient")
    d, m = 128, spec.n_patient
    gen = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(m, d, device=dev, generator=gen)
    ys = [torch.randn(n, d, device=dev, generator=gen) for n in pb.sizes]
    w = torch.randn(d, d, device=dev, generator=gen) / d ** 0.5
    b = torch.randn(d, device=dev, generator=gen)
    wcat, bias = ops.layer_cat_weights_([w], False, ys, [None] * len(ys), pb.offs, d, d + 32 * pb.nw, d, [b])
    wcat_h = wcat.clone()
    hv = ops.layer_cat_half_(wcat_h, d)
    out = torch.empty(m, d, device=dev)
    rs = pb.rscale_in()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for i in range(args.reps):
            flush.fill_(i)
            torch.cuda._sleep(int(0.004 * 1.9e9))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return statistics.median(ts)

    os.environ["B2G_LAYER_DBG"] = "0"
    gout = torch.randn(m, d, device=dev, generator=gen)
    adj = {"adjT fwd": lambda: ops.layer_adjT_tc_(x, pb.bits_out, pb, [None] * len(ys), pb.col_scale_out()),
           "adjT bwd (+dW, +db)": lambda: ops.layer_adjT_tc_(gout, pb.bits_in, pb, rs, None, with_colsum=True, dense_b=x)}
    for name, fn in adj.items():
        for bs in (2, 3):
            os.environ["B2G_ADJT_BSTAGES"] = str(bs)
            for dbg in (0, 1):
                os.environ["B2G_ADJT_DBG"] = str(dbg)
                print(f"{name} bstages={bs} dbg={dbg}: {timed(fn):.4f} ms", flush=True)
    os.environ.pop("B2G_ADJT_BSTAGES", None)
    os.environ["B2G_ADJT_DBG"] = "0"
    fns = {"tf32": lambda: ops.layer_fwd_tc_(x, wcat, bias, pb.bits_in, pb, rs, out),
           "f16": lambda: ops.layer_fwd_tc_(x, wcat_h, bias, pb.bits_in, pb, rs, out, None, hv)}
    for mode, fn in ():
        for dbg in (0,):
            os.environ["B2G_LAYER_DBG"] = str(dbg)
            print(f"{mode} dbg={dbg}: {timed(fn):.4f} ms", flush=True)
        for st in ():
            os.environ["B2G_LAYER_DBG"] = "0"
            os.environ["B2G_LAYER_STAGES"] = str(st)
            print(f"{mode} stages={st}: {timed(fn):.4f} ms", flush=True)
        os.environ.pop("B2G_LAYER_STAGES", None)
        for order in ():
            os.environ["B2G_LAYER_ORDER"] = str(order)
            os.environ["B2G_LAYER_DBG"] = "0"
            print(f"{mode} order={order}: {timed(fn):.4f} ms", flush=True)
        os.environ.pop("B2G_LAYER_ORDER", None)
        for d2 in (9, 10, 12):
            os.environ["B2G_LAYER_DBG"] = str(d2)
            fn()
            torch.cuda.synchronize()
        os.environ["B2G_LAYER_DBG"] = "8"
        fn()
        torch.cuda.synchronize()
        os.environ["B2G_LAYER_DBG"] = "0"


if __name__ == "__main__":
    main()

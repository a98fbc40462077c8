#!/usr/bin/env python
"""Micro-benchmark of the HeteroConv patient-side kernels (csrc/layer_tc.cu) and of one whole layer, fused vs per-relation.

  python tools/layer_bench.py [--workload C4s8] [--reps 7] [--out gpurun_out/layer_bench.json]

Times are CUDA events on the launching stream, inputs larger than L2 (or L2 flushed), after warm-up.  Algorithmic bytes:
k_layer_tf32  = read x_p [M,128] + write out_p [M,128] + bits [M,nw]                  (the forward half of bytes_min)
k_adjT_tf32   = read x_p [M,128] + bits
layer fwd+bwd = bytes_min = 5 N_p d 4 + 2 E 4 + 6 (N_p + 1) 4
"""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multi-modal-gnn_b200"


def timed(fn, reps, flush):
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        flush.fill_(i & 0xFF)
        torch.cuda._sleep(int(0.004 * 1.9e9))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4s8")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only-layer", action="store_true", help="skip the per-kernel timings (for ncu captures of one layer)")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    M = importlib.import_module(PKG + ".model")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda:0")
    spec = pkg.synth.SPECS[args.workload]
    g = pkg.synth.make_graph(spec, seed=42, device=dev if spec.n_patient >= 500_000 else "cpu").to(dev)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2, "use_batch_norm": True, "activation": "relu"}}
    torch.manual_seed(0)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    model._init_embeddings(g)
    model.train()
    ops.set_precision("tf32")
    gi = model._graph_index(g)
    pb = gi.hub_bits("patient")
    d = 128
    m = spec.n_patient
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    res = {"workload": args.workload, "n_patient": m, "nw": pb.nw if pb else None, "hbm_peak_gbs": peaks}
    if pb is not None and not args.only_layer:
        x = torch.randn(m, d, device=dev, generator=gen)
        ys = [torch.randn(n, d, device=dev, generator=gen) for n in pb.sizes]
        w = torch.randn(d, d, device=dev, generator=gen) / d ** 0.5
        b = torch.randn(d, device=dev, generator=gen)
        wcat, bias = ops.layer_cat_weights_([w], False, ys, [None] * len(ys), pb.offs, d, d + 32 * pb.nw, d, [b])
        out = torch.empty(m, d, device=dev)
        sums = torch.zeros(2 * d, dtype=torch.float64, device=dev)
        rs = pb.rscale_in()
        wcat_h = wcat.clone()
        hv = ops.layer_cat_half_(wcat_h, d)
        for name, fn, nbytes in [
            ("k_layer_tf32 (f16 adjacency)", lambda: ops.layer_fwd_tc_(x, wcat_h, bias, pb.bits_in, pb, rs, out, None, hv), 8 * m * d + 4 * m * pb.nw),
            ("k_layer_tf32 (f16 adjacency)+stats", lambda: ops.layer_fwd_tc_(x, wcat_h, bias, pb.bits_in, pb, rs, out, sums, hv), 8 * m * d + 4 * m * pb.nw),
            ("k_layer_tf32", lambda: ops.layer_fwd_tc_(x, wcat, bias, pb.bits_in, pb, rs, out), 8 * m * d + 4 * m * pb.nw),
            ("k_layer_tf32+stats", lambda: ops.layer_fwd_tc_(x, wcat, bias, pb.bits_in, pb, rs, out, sums), 8 * m * d + 4 * m * pb.nw),
            ("k_adjT_tf32", lambda: ops.layer_adjT_tc_(x, pb.bits_out, pb, [None] * len(ys), pb.col_scale_out()), 4 * m * d + 4 * m * pb.nw),
            ("k_linear_tf32 (x W^T + b only, round-1 kernel)", lambda: ops.linear_fwd_(x, w, b, out), 8 * m * d),
        ]:
            med, mn = timed(fn, args.reps, flush)
            res[name] = {"ms": med, "ms_min": mn, "algorithmic_GBps": nbytes / med / 1e6, "frac_of_hbm_peak": nbytes / med / 1e6 / peaks}
            print(name, res[name], flush=True)
    # whole layer, forward + backward (the bench's layer_microbench), fused and per-relation
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    xs = {nt: torch.randn(n, d, device=dev, generator=gen).requires_grad_(True) for nt, n in counts.items()}
    gout = {nt: torch.randn(n, d, device=dev, generator=gen) for nt, n in counts.items()}
    params = list(model.convs[0].parameters())

    def once():
        for t in list(xs.values()) + params:
            t.grad = None
        o = model._layer(0, xs, gi)
        torch.autograd.backward([o[nt] for nt in o], [gout[nt] for nt in o])

    e_und = spec.e_lab + spec.e_dx + spec.e_med
    bytes_min = 5 * m * d * 4 + 2 * e_und * 4 + 6 * (m + 1) * 4
    saved = M.HeteroRGCN._layer_fused
    for label, fn in [("layer_fused", saved)] + ([] if args.only_layer else [("layer_per_relation", lambda self, *a, **k: None)]):
        M.HeteroRGCN._layer_fused = fn
        try:
            med, mn = timed(once, args.reps, flush)
        finally:
            M.HeteroRGCN._layer_fused = saved
        res[label] = {"ms": med, "ms_min": mn, "bytes_min": bytes_min, "frac_of_hbm_peak": bytes_min / med / 1e6 / peaks,
                      "directed_edges_per_s": spec.directed_edges_per_layer / (med * 1e-3)}
        print(label, res[label], flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

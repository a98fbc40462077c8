#!/usr/bin/env python
"""Exact-fp32 message passing onto the patient rows at the C4 shard: b2g_gather_reduce (source rows through L1 / L2) against
b2g_gather_reduce_staged (source tables staged in shared memory by bulk copies, north_star (b)), and the BatchNorm column
reduction (k_col_reduce), each against its algorithmic HBM bytes.  Also the target of the ncu captures of these kernels.

  python tools/gather_bench.py [--workload C4s8] [--reps 7] [--out gpurun_out/gather_bench.json]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C4s8")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + ".ops")
    M = importlib.import_module(PKG + ".model")
    L = importlib.import_module(PKG + "._lib")
    lib = L.load()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.isfile(pk) else 6650.0
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
    xs = [torch.randn(n, d, device=dev, generator=gen) for n in pb.sizes]
    csrs = [r.by_dst for r in pb.in_rel]
    rsc = [r.by_dst.inv_deg for r in pb.in_rel]
    out = torch.empty(m, d, device=dev)
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

    e_in = sum(c.n_edges for c in csrs)
    nbytes = 4 * m * d + 4 * e_in + 4 * 3 * (m + 1) + 4 * m * 3 + sum(4 * x.numel() for x in xs)      # out + col + rowptr + 1/deg + tables
    res = {"workload": args.workload, "n_patient": m, "edges_in": e_in, "hbm_peak_gbs": peak, "algorithmic_bytes": nbytes}
    for staged, stream in ((False, False), (False, True), (True, True)):
        ops.GATHER_STAGED = staged
        ops.GATHER_STREAM, ops.STREAM_MIN_ROWS, ops.STREAM_MAX_AVG_DEG = stream, 1, 10 ** 9
        ms = timed(lambda: ops.gather_reduce_(csrs, xs, rsc, [None] * len(xs), out, False))
        key = "b2g_gather_reduce_staged" if staged else ("b2g_gather_reduce (k_gather_reduce_stream)" if stream else "b2g_gather_reduce (k_gather_reduce, warp per row)")
        res[key] = {"ms": ms, "algorithmic_GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak, "edges_per_s": e_in / (ms * 1e-3)}
        print(key, res[key], flush=True)
    # BatchNorm statistics of a [m, 128] activation (k_col_reduce<0>): reads the array once
    x = torch.randn(m, d, device=dev, generator=gen)
    bn = torch.nn.BatchNorm1d(d).to(dev)
    bn.train()
    ops.set_precision("fp32")
    ms = timed(lambda: ops.BNActDropFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, True, 1, 0.0, 0, 0, 1e-5, 0.1))
    res["bn_fwd (k_col_reduce<0> + k_bn_apply)"] = {"ms": ms, "algorithmic_GBps": 3 * 4 * m * d / ms / 1e6, "frac_of_hbm_peak": 3 * 4 * m * d / ms / 1e6 / peak}
    print("bn_fwd", res["bn_fwd (k_col_reduce<0> + k_bn_apply)"], flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run-to-run determinism of the 100-epoch C1 training used by tests/test_gpu_e2e_parity.py, GPU side and CPU-oracle side
separately: the same run repeated N times in one process must give bit-identical loss sequences.

  python tools/determinism_check.py [--runs 3] [--epochs 40]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
PKG = "multi-modal-gnn_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import hetero_rgcn_ref as R
    pkg = importlib.import_module(PKG)
    M, T, ops = (importlib.import_module(PKG + m) for m in (".model", ".trainer", ".ops"))
    dev = torch.device("cuda:0")
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True, "activation": "relu"},
           "train": {"loss": "mae", "epochs": args.epochs, "early_stopping_patience": 1000, "optimizer": {"type": "adam", "lr": 1e-3, "weight_decay": 1e-5},
                     "lr_scheduler": {"enabled": False}}}
    g = pkg.synth.make_graph("C1", seed=42)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = list(g.edge_types)
    sd = R.init_state(counts, ets, seed=3)

    def gpu_run():
        model = M.build_model(cfg, (g.node_types, g.edge_types), None)
        masker = T.EdgeMasker(pkg.synth.make_graph("C1", seed=42), 0.7, 0.15, 0.15, 0.2, 42)
        trainer = T.Trainer(model, masker.data, masker, cfg, dev)
        model._init_embeddings(trainer.data)
        model.load_state_dict(sd)
        out = []
        for epoch in range(args.epochs):
            out.append(float(trainer.train_epoch(seed=1000 + epoch)))
            out.append(float(trainer.validate("val")))
        return out

    def cpu_run():
        masker = T.EdgeMasker(pkg.synth.make_graph("C1", seed=42), 0.7, 0.15, 0.15, 0.2, 42)
        ei = g["patient", "has_lab", "lab"].edge_index
        attr = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
        tr = masker.train_mask
        pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr]
        w = R.lab_weights(li, tgt, counts["lab"])
        sd_ref = {k: v.clone() for k, v in sd.items()}
        params = [sd_ref[k].requires_grad_(True) for k in R.trainable_keys(sd_ref)]
        opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
        torch.set_num_threads(os.cpu_count() or 1)
        out = []
        for epoch in range(min(args.epochs, 15)):
            sup = R.supervision_mask(int(tr.sum()), 0.2, 1000 + epoch)
            opt.zero_grad()
            pred = R.predict_lab_values(sd_ref, counts, ets, g.edge_index_dict, pi, li, True, p_drop=0.0, mask_fn=None)
            loss = R.weighted_loss(pred, tgt, li, w, sup, "mae")
            loss.backward()
            opt.step()
            out.append(float(loss.detach()))
        return out

    for mode in ("tf32", "fp32"):
        ops.set_precision(mode)
        runs = [gpu_run() for _ in range(args.runs)]
        same = all(r == runs[0] for r in runs)
        first = next((i for i in range(len(runs[0])) if any(r[i] != runs[0][i] for r in runs)), None)
        print(f"GPU {mode}: {args.runs} runs of {args.epochs} epochs bit-identical: {same}" + ("" if same else f" (first difference at value {first}: {[r[first] for r in runs]})"),
              flush=True)
    if not args.no_cpu:
        runs = [cpu_run() for _ in range(args.runs)]
        same = all(r == runs[0] for r in runs)
        first = next((i for i in range(len(runs[0])) if any(r[i] != runs[0][i] for r in runs)), None)
        print(f"CPU oracle ({torch.get_num_threads()} threads): {args.runs} runs bit-identical: {same}" + ("" if same else f" (first difference at epoch {first}: {[r[first] for r in runs]})"),
              flush=True)


if __name__ == "__main__":
    main()

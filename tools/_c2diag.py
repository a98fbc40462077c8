import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
from oracle import hetero_rgcn_ref as R
PKG = "multi-modal-gnn_b200"
pkg = importlib.import_module(PKG); ops = importlib.import_module(PKG + ".ops"); M = importlib.import_module(PKG + ".model")
dev = torch.device("cuda:0")
g = pkg.synth.make_graph("C2", seed=42)
counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}; ets = list(g.edge_types)
sd = R.init_state(counts, ets, seed=11)
ei = g["patient", "has_lab", "lab"].edge_index; attr = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
tr = R.split_masks(ei.shape[1])[0]
pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr]
w = R.lab_weights(li, tgt, counts["lab"]); sup = R.supervision_mask(int(tr.sum()), 0.2, 1234)
torch.set_num_threads(os.cpu_count())
l32, p32, g32 = R.train_step_grads({k: v.clone() for k, v in sd.items()}, counts, ets, g.edge_index_dict, pi, li, tgt, sup, w, "mse", 0.0)
sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
l64, p64, g64 = R.train_step_grads(sd64, counts, ets, g.edge_index_dict, pi, li, tgt.double(), sup, w.double(), "mse", 0.0)
cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True, "activation": "relu"}}
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for mode in ("fp32", "tf32"):
    ops.set_precision(mode)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    gd = pkg.synth.make_graph("C2", seed=42).to(dev)
    model._init_embeddings(gd); model.load_state_dict(sd); model.train()
    pred = model.predict_lab_values(gd, pi.to(dev), li.to(dev))
    loss = ops.weighted_loss(pred, tgt.to(dev), li.to(dev), w.to(dev), sup.to(dev), "mse"); loss.backward()
    params = dict(model.named_parameters())
    print(f"== {mode}: loss gpu {float(loss):.7f} cpu32 {float(l32):.7f} cpu64 {float(l64):.7f}; pred vs64 gpu {rel(pred, p64):.2e} cpu32 {rel(p32, p64):.2e}")
    rows = []
    for k, gr in g64.items():
        if gr is None or k.endswith("lin_l.bias") or k in ("patient_transform.0.bias", "patient_transform.4.bias"): continue
        rows.append((rel(params[k].grad, gr), rel(g32[k], gr), k))
    rows.sort(reverse=True)
    for eg, ec, k in rows[:12]: print(f"   {k:60s} gpu-vs-64 {eg:.2e}   cpu32-vs-64 {ec:.2e}")
    print("   median gpu", sorted(r[0] for r in rows)[len(rows)//2], "median cpu32", sorted(r[1] for r in rows)[len(rows)//2])
    k = "embeddings.patient.weight"
    d = (params[k].grad.double().cpu() - g64[k]).abs().max(1)[0]; mx = float(g64[k].abs().max())
    bad = (d > 1e-3 * mx).nonzero().squeeze(1)
    deg = torch.bincount(ei[0], minlength=counts["patient"])
    print("   patient rows with error > 1e-3 max:", bad.numel(), "of", d.numel(), "degrees of worst:", deg[d.topk(5).indices].tolist(), "their |g64| row max:", g64[k].abs().max(1)[0][d.topk(5).indices].tolist(), "max", mx)

"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Each fixture = inputs + what the reference's own src/model.py / src/train.py produced for them on the
PyG shim, with dropout 0 (device Philox streams cannot match ATen's CPU stream bit-wise; dropout is
validated separately with replayed masks).  Fixtures are consumed by tests/ on machines without
/root/reference (the GPU box).
"""
from __future__ import annotations

import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def graph_blob(g):
    blob = {"num_nodes": {nt: int(g[nt].num_nodes) for nt in g.node_types}, "edge_types": list(g.edge_types),
            "edge_index": {"__".join(et): g[et].edge_index.to(torch.int32) for et in g.edge_types},
            "edge_attr": g["patient", "has_lab", "lab"].edge_attr.clone()}
    return blob


def run_case(spec_name, loss, store_state, store_all_grads):
    pkg = importlib.import_module("multi-modal-gnn_b200")
    clock = H.FakeClock()
    M, T = H.load_reference(clock)
    g = pkg.synth.make_graph(spec_name, seed=42)
    d = H.to_shim_data(g)
    cfg = H.make_config(dropout=0.0, loss=loss)
    torch.manual_seed(1234)
    masker = T.EdgeMasker(d, 0.7, 0.15, 0.15, 0.2, 42)
    model = M.build_model(cfg, (d.node_types, d.edge_types), None)
    n_params_before_tables = sum(p.numel() for p in model.parameters())
    trainer = T.Trainer(model, d, masker, cfg, torch.device("cpu"))   # optimizer built w/o tables (N2)
    model._init_embeddings(d)
    # de-trivialise BN affine params so BN gradients are exercised
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if ("batch_norms" in name or "patient_transform.1" in name or "patient_transform.5" in name):
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}

    sup_seed = int(clock.now) + 1                 # the seed train_epoch's get_masked_data will use
    loss_train = trainer.train_epoch()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}

    torch.manual_seed(sup_seed)
    ei_train = masker.edge_index[:, masker.train_mask]
    sup_mask = torch.rand(int(masker.train_mask.sum())) < 0.2
    # eval-mode products at a fully stored state: the *before* parameters + the BN buffers after the step
    loss_val = trainer.validate("val")
    loss_test = trainer.validate("test")
    eval_state = dict(before)
    eval_state.update({k: v for k, v in after.items() if "running" in k or "num_batches" in k})
    model3 = M.build_model(cfg, (d.node_types, d.edge_types), None)
    model3._init_embeddings(d)
    model3.load_state_dict(eval_state)
    model3.eval()
    with torch.no_grad():
        ei_val = masker.edge_index[:, masker.val_mask]
        pred_val = model3.predict_lab_values(d, ei_val[0], ei_val[1])
        enc = model3.encode_nodes(d)
        fwd = model3(d)
        ei_all = masker.edge_index
        pred_all = model3.predict_lab_values(d, ei_all[0], ei_all[1])
        eval_loss_val = float(M.compute_regression_loss(pred_val, masker.edge_attr[masker.val_mask].squeeze(-1), loss))
    # train-mode prediction replay from the *before* state (fresh model so BN buffers start equal)
    model2 = M.build_model(cfg, (d.node_types, d.edge_types), None)
    model2._init_embeddings(d)
    model2.load_state_dict(before)
    model2.train()
    with torch.no_grad():
        pred_train = model2.predict_lab_values(d, ei_train[0], ei_train[1])

    blob = {
        "spec": spec_name, "loss_fn": loss, "graph": graph_blob(g),
        "n_params_before_tables": n_params_before_tables,
        "split": {"train": masker.train_mask, "val": masker.val_mask, "test": masker.test_mask},
        "sup_seed": sup_seed, "sup_mask": sup_mask, "lab_weights": trainer.lab_weights.clone(),
        "loss_train": float(loss_train), "loss_val": float(loss_val), "loss_test": float(loss_test),
        "pred_train": pred_train, "pred_val": pred_val, "pred_all": pred_all, "eval_loss_val": eval_loss_val,
        "degree": torch.bincount(masker.edge_index[0], minlength=int(g["patient"].num_nodes)),
        "state_keys": list(before.keys()),
        "state_checksum": {k: float(v.double().sum()) for k, v in before.items()},
        "after_buffers": {k: v for k, v in after.items() if "running" in k or "num_batches" in k},
        "grad_is_none": sorted(n for n, v in grads.items() if v is None),
        "grad_norm": {n: float(v.double().norm()) for n, v in grads.items() if v is not None},
        "optimizer_param_count": sum(p.numel() for grp in trainer.optimizer.param_groups for p in grp["params"]),
    }
    blob["state_before"] = before
    if store_state:
        blob["encode_eval"] = enc
        blob["forward_eval"] = fwd
        blob["after_params"] = {k: after[k] for k in ("patient_transform.0.weight", "edge_predictor.mlp.0.weight",
                                                      "convs.0.convs.lab__has_lab_rev__patient.lin_l.weight",
                                                      "batch_norms.0.patient.weight")}
    if store_all_grads:
        blob["grads"] = {n: v for n, v in grads.items() if v is not None}
    else:
        keep = ("patient_transform.0.weight", "patient_transform.8.bias", "edge_predictor.mlp.0.weight",
                "edge_predictor.mlp.6.bias", "convs.0.convs.lab__has_lab_rev__patient.lin_l.weight",
                "convs.1.convs.patient__has_lab__lab.lin_r.weight", "batch_norms.0.patient.weight",
                "embeddings.lab.weight")
        blob["grads"] = {n: grads[n] for n in keep if grads.get(n) is not None}
    return blob


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, spec, loss, st, ag in (("tiny_mae", "tiny", "mae", True, True),
                                     ("tiny_mse", "tiny", "mse", False, False),
                                     ("c1_mae", "C1", "mae", False, False)):
        blob = run_case(spec, loss, st, ag)
        if name == "tiny_mse":               # same before-state as tiny_mae? no: fresh init -> keep it
            pass
        path = os.path.join(OUT, name + ".pt")
        torch.save(blob, path)
        print(name, os.path.getsize(path) / 1e6, "MB", "loss", blob["loss_train"], blob["loss_val"],
              "opt params", blob["optimizer_param_count"], "none-grads", len(blob["grad_is_none"]))


if __name__ == "__main__":
    main()

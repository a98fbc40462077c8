"""End-to-end parity of the BENCHMARKED mode (ops.PRECISION = 'tf32': tcgen05 linears, bit-adjacency layer kernels, tensor-core
decoder) against the CPU oracle (oracle/hetero_rgcn_ref.py, pinned to the unmodified reference by tests/golden/*):

  * north_star end criterion "final R^2 / MAE within 1e-3": the reference's training loop (train.py:433-544: train_epoch +
    validate + ReduceLROnPlateau, 100 epochs) on the C1-shaped graph, then the reference's evaluation arithmetic
    (evaluate.py:397-445: test-split predictions, per-lab +-3 sigma winsorisation, MAE / RMSE / R^2) -- the CUDA Trainer in
    its default precision vs the oracle + torch.optim.Adam on the CPU, with dropout 0 and with dropout 0.2 (the device's
    Philox masks replayed into the oracle every step);
  * the tensor-core decoder kernels against the ORACLE's EdgeRegressionHead (not against the repo's own SIMT kernel);
  * one whole training step on the benchmarked-size C2 graph (46,520 patients, 5 M lab edges) in both precision modes.
"""
import importlib
import os
import time

import numpy as np
import pytest
import torch

from oracle import eval_metrics_ref as E
from oracle import hetero_rgcn_ref as R

pytestmark = pytest.mark.gpu
PKG = "multi-modal-gnn_b200"


def _mods():
    return (importlib.import_module(PKG), importlib.import_module(PKG + ".graph"), importlib.import_module(PKG + ".ops"),
            importlib.import_module(PKG + ".model"), importlib.import_module(PKG + ".trainer"), importlib.import_module(PKG + ".metrics"))


@pytest.fixture
def tf32_mode():
    ops = importlib.import_module(PKG + ".ops")
    old = ops.PRECISION
    ops.set_precision("tf32")
    yield ops
    ops.set_precision(old)


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def normerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _cfg(dropout, loss="mae"):
    return {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": dropout, "use_batch_norm": True,
                      "activation": "relu"},
            "train": {"loss": loss, "epochs": 100, "early_stopping_patience": 15, "mask_fraction": 0.2,
                      "optimizer": {"type": "adam", "lr": 1e-3, "weight_decay": 1e-5},
                      "lr_scheduler": {"enabled": True, "type": "reduce_on_plateau", "factor": 0.5, "patience": 10}}}


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_final_metrics_after_100_epochs_match_oracle(dropout, tf32_mode):
    """|delta MAE|, |delta RMSE|, |delta R^2| <= 1e-3 on the test split after the reference's 100-epoch loop (config.yaml:236:
    epochs 100, loss 'mae' config.yaml:232, Adam 1e-3 / wd 1e-5, ReduceLROnPlateau x0.5 patience 10), default precision."""
    pkg, G, ops, M, T, MET = _mods()
    assert ops.PRECISION == "tf32"
    dev = torch.device("cuda:0")
    epochs = int(os.environ.get("B2G_E2E_EPOCHS", "100"))
    g = pkg.synth.make_graph("C1", seed=42)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = list(g.edge_types)
    sd = R.init_state(counts, ets, seed=3)
    cfg = _cfg(dropout)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None)
    masker = T.EdgeMasker(pkg.synth.make_graph("C1", seed=42), 0.7, 0.15, 0.15, 0.2, 42)
    trainer = T.Trainer(model, masker.data, masker, cfg, dev)          # optimizer before the lazy tables exist (note N2)
    model._init_embeddings(trainer.data)
    model.load_state_dict(sd)

    ei = g["patient", "has_lab", "lab"].edge_index
    attr = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
    tr, va, te = masker.train_mask, masker.val_mask, masker.test_mask
    pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr]
    w = R.lab_weights(li, tgt, counts["lab"])
    sd_ref = {k: v.clone() for k, v in sd.items()}
    keys = R.trainable_keys(sd_ref)
    params = [sd_ref[k].requires_grad_(True) for k in keys]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=10)
    # ONE thread for the oracle: torch's multi-threaded CPU kernels are not reproducible run to run (tools/determinism_check.py:
    # the oracle's loss differs in the 7th digit after 3 epochs between two runs with 16 threads, and by 5e-5 after 6 epochs between
    # 4 and 8 threads, while the CUDA path is bit-identical run to run in both precision modes); over 100 epochs that noise alone
    # moves the oracle's final R^2 by several 1e-4, i.e. a 1e-3 comparison against a multi-threaded oracle is flaky by construction
    n_threads_before = torch.get_num_threads()
    torch.set_num_threads(1)

    lr_gpu, lr_ref, max_loss_gap, max_val_gap = [], [], 0.0, 0.0
    t0 = time.time()
    for epoch in range(epochs):
        torch.manual_seed(10_000 + epoch)                              # the dropout seed is drawn from torch's CPU generator
        loss_gpu = trainer.train_epoch(seed=1000 + epoch)
        streams = model._last_streams
        tags = dict(streams.log)

        def mask_fn(tag, x):
            return ops.dropout_mask(x.numel(), dropout, streams.seed, tags[tag], dev).cpu().view_as(x).to(x.dtype)

        val_gpu = trainer.validate("val")

        sup = R.supervision_mask(int(tr.sum()), 0.2, 1000 + epoch)
        opt.zero_grad()
        pred = R.predict_lab_values(sd_ref, counts, ets, g.edge_index_dict, pi, li, True, p_drop=dropout,
                                    mask_fn=mask_fn if dropout > 0 else None)
        loss = R.weighted_loss(pred, tgt, li, w, sup, "mae")
        loss.backward()
        opt.step()
        with torch.no_grad():
            frozen = {k: v.detach() for k, v in sd_ref.items()}
            pv = R.predict_lab_values(frozen, counts, ets, g.edge_index_dict, ei[0][va], ei[1][va], False)
            val_ref = float(R.regression_loss(pv, attr[va], "mae"))
        # ReduceLROnPlateau (train.py:490-494) compares validation losses that are nearly tied on this graph (the frozen
        # tables, note N2, leave little to learn), so a 1e-4 difference can flip a plateau decision and fork the two runs'
        # hyper-parameters.  Both sides therefore follow the REFERENCE's decisions; the validation losses themselves are
        # compared every epoch.
        sched.step(val_ref)
        trainer.scheduler.step(val_ref)
        lr_ref.append(opt.param_groups[0]["lr"])
        lr_gpu.append(trainer.optimizer.param_groups[0]["lr"])
        max_val_gap = max(max_val_gap, abs(val_gpu - val_ref) / abs(val_ref))
        max_loss_gap = max(max_loss_gap, abs(loss_gpu - float(loss.detach())) / abs(float(loss.detach())))
    torch.set_num_threads(n_threads_before)      # (the final evaluation below is one forward pass: thread count does not matter there)
    print(f"[e2e dropout={dropout}] {epochs} epochs in {time.time() - t0:.1f}s, max relative gap: train loss {max_loss_gap:.2e}, "
          f"validation loss {max_val_gap:.2e}; final lr gpu/ref {lr_gpu[-1]:.2e}/{lr_ref[-1]:.2e}")
    assert lr_gpu == lr_ref
    assert max_loss_gap <= 5e-3 and max_val_gap <= 5e-3

    # evaluate.py:397-445 on the test split
    model.eval()
    pt, lt, tt = ei[0][te].to(dev), ei[1][te].to(dev), attr[te].to(dev)
    with torch.no_grad():
        pred_gpu = model.predict_lab_values(trainer.data, pt, lt)
        frozen = {k: v.detach() for k, v in sd_ref.items()}
        pred_ref = R.predict_lab_values(frozen, counts, ets, g.edge_index_dict, ei[0][te], ei[1][te], False)
    got = MET.evaluate_predictions(pred_gpu, tt, lt, counts["lab"])["overall"]
    pw, _ = E.winsorize(pred_ref.numpy().astype(np.float64), attr[te].numpy().astype(np.float64), ei[1][te].numpy())
    want = E.regression_metrics(pw, attr[te].numpy().astype(np.float64))
    print(f"[e2e dropout={dropout}] test metrics  gpu: mae {got['mae']:.5f} rmse {got['rmse']:.5f} r2 {got['r2']:.5f}   "
          f"oracle: mae {want['mae']:.5f} rmse {want['rmse']:.5f} r2 {want['r2']:.5f}")
    for k in ("mae", "rmse", "r2"):
        assert abs(got[k] - want[k]) <= 1e-3, (k, got[k], want[k])
    # individual predictions after 100 chaotic optimizer steps: root-mean-square deviation relative to the predictions' spread
    # (individual predictions drift apart over 100 chaotic optimizer steps -- reported, not asserted; the criterion is the metrics)
    rms = float((pred_gpu.cpu().double() - pred_ref.double()).pow(2).mean().sqrt() / attr[te].double().std())
    print(f"[e2e dropout={dropout}] prediction deviation: rms {rms:.2e} of std(target), max {relerr(pred_gpu, pred_ref):.2e} of max|pred|")


@pytest.mark.parametrize("m,p_drop,frac_active", [(5000, 0.0, 1.0), (43038, 0.2, 0.2), (300000, 0.2, 0.2)])
def test_tensor_core_decoder_matches_oracle_head(m, p_drop, frac_active, tf32_mode):
    """k_decoder_fwd_tc / k_decoder_bwd_tc (tf32 mode) vs the oracle's EdgeRegressionHead on cat([h_p[pi], h_l[li]])
    (model.py:305-333,373-386) in float64, with the device's dropout masks replayed and a sparse upstream gradient."""
    pkg, G, ops, M, T, MET = _mods()
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(m)
    n_p, n_l, d = 5000, 160, 128
    hp, hl = torch.randn(n_p, d, generator=gen), torch.randn(n_l, d, generator=gen)
    pi, li = torch.randint(0, n_p, (m,), generator=gen), torch.randint(0, n_l, (m,), generator=gen)
    head = M.EdgeRegressionHead(2 * d, [64, 32], 1, p_drop).to(dev)
    head.train()
    sd = {"h.mlp." + k.split("mlp.")[1]: v.detach().cpu().clone() for k, v in head.state_dict().items()}
    go = torch.randn(m, generator=gen) * (torch.rand(m, generator=gen) < frac_active)
    pairs = G.PairIndex(pi.to(dev), li.to(dev), n_p, n_l)
    streams = M._DropoutStreams(p_drop > 0)
    hpd, hld = hp.to(dev).requires_grad_(True), hl.to(dev).requires_grad_(True)
    ops.PROFILE = []
    pred = head.forward_pairs(hpd, hld, pairs, streams, "h")
    pred.backward(go.to(dev))
    names = [p[0] for p in ops.PROFILE]
    ops.PROFILE = None
    assert "b2g_decoder_fwd_tc" in names and "b2g_decoder_bwd_tc" in names, "the tensor-core decoder kernels must be the ones tested"
    tags = dict(streams.log)

    def mask_fn(tag, x):
        return ops.dropout_mask(x.numel(), p_drop, streams.seed, tags[tag], dev).cpu().view_as(x).to(x.dtype)

    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    hpr, hlr = hp.double().requires_grad_(True), hl.double().requires_grad_(True)
    ref = R.edge_head(sdr, "h", torch.cat([hpr[pi], hlr[li]], 1), True, p_drop, mask_fn if p_drop > 0 else None)
    ref.backward(go.double())
    e_pred = relerr(pred, ref)
    errs = {"h_p": normerr(hpd.grad, hpr.grad), "h_l": normerr(hld.grad, hlr.grad)}
    for name, prm in head.named_parameters():
        errs[name] = normerr(prm.grad, sdr["h." + name].grad)
    print(f"[tc decoder m={m}] pred max-rel {e_pred:.2e}; gradient norm errors", {k: round(v, 5) for k, v in errs.items()})
    assert e_pred <= 1e-2                      # north_star: tensor-core outputs <= 1e-2
    # gradients: a TF32-sized perturbation (1e-3) of a pre-activation that lies within 1e-3 of zero flips its ReLU derivative,
    # i.e. a fraction ~1e-3 of the hidden units changes its contribution by O(1): a relative error ~sqrt(1e-3) = 3e-2 in norm
    # is inherent to ANY reduced-precision forward through ReLU (measured 2.4e-2 .. 3.3e-2); the exact-fp32 kernels are held
    # to 1e-4 in test_gpu_parity.py::test_fused_decoder_matches_oracle_head
    assert max(errs.values()) <= 5e-2, errs


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_whole_training_step_on_C2_matches_oracle(mode):
    """One Trainer-style step (forward + weighted MSE + backward, dropout 0) on the benchmarked C2 graph (46,520 patients, 5 M lab
    edges): loss, predictions and every parameter gradient, in the exact-fp32 mode and in the benchmarked tf32 mode.

    At this size the gradients are ill-conditioned in fp32 itself (BatchNorm's backward subtracts column means over 46 k rows; the
    weight gradients are sums of 46 k cancelling terms): the reference's OWN fp32 arithmetic (the oracle on the CPU) deviates from
    a float64 evaluation of the same step by 1e-1 of max|grad| on the patient table and ~2e-3 in the median.  The float64 oracle
    is therefore the yardstick, and the fp32 oracle's distance to it is the scale:
      fp32 mode  every gradient at least as close to float64 as the reference's fp32 arithmetic (factor 2 + 2e-4 slack);
      tf32 mode  predictions <= 1e-2, loss <= 2e-3 relative, every gradient <= 8e-2 in norm (measured: ~2.6e-2 median, the
                 truncation of tcgen05's kind::tf32 operands is biased and does not average out over cancelling sums); what this
                 does to training is pinned by test_final_metrics_after_100_epochs_match_oracle."""
    pkg, G, ops, M, T, MET = _mods()
    dev = torch.device("cuda:0")
    old = ops.PRECISION
    ops.set_precision(mode)
    try:
        g = pkg.synth.make_graph("C2", seed=42)
        counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
        ets = list(g.edge_types)
        sd = R.init_state(counts, ets, seed=11)
        ei = g["patient", "has_lab", "lab"].edge_index
        attr = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
        tr = R.split_masks(ei.shape[1])[0]
        pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr]
        w = R.lab_weights(li, tgt, counts["lab"])
        sup = R.supervision_mask(int(tr.sum()), 0.2, 1234)
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.time()
        loss32, pred32, grads32 = R.train_step_grads({k: v.clone() for k, v in sd.items()}, counts, ets, g.edge_index_dict, pi, li, tgt,
                                                     sup, w, "mse", 0.0)
        sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        loss64, pred64, grads64 = R.train_step_grads(sd64, counts, ets, g.edge_index_dict, pi, li, tgt.double(), sup, w.double(), "mse", 0.0)
        t_ref = time.time() - t0

        model = M.build_model(_cfg(0.0, "mse"), (g.node_types, g.edge_types), None).to(dev)
        gd = pkg.synth.make_graph("C2", seed=42).to(dev)
        model._init_embeddings(gd)
        model.load_state_dict(sd)
        model.train()
        pred = model.predict_lab_values(gd, pi.to(dev), li.to(dev))
        loss = ops.weighted_loss(pred, tgt.to(dev), li.to(dev), w.to(dev), sup.to(dev), "mse")
        loss.backward()
        params = dict(model.named_parameters())
        gmax = max(float(gr.abs().max()) for gr in grads64.values() if gr is not None)
        e_gpu, e_cpu, n_gpu = {}, {}, {}
        for k, gr in grads64.items():
            if gr is None:
                assert params[k].grad is None, f"{k}: the reference leaves this gradient None (note N8)"
                continue
            if k in ("patient_transform.0.bias", "patient_transform.4.bias") or k.endswith("lin_l.bias"):
                # a bias in front of a BatchNorm: its true gradient is exactly zero (the batch mean is subtracted)
                assert float(params[k].grad.abs().max()) <= 1e-4 * gmax, k
                continue
            e_gpu[k], e_cpu[k], n_gpu[k] = relerr(params[k].grad, gr), relerr(grads32[k], gr), normerr(params[k].grad, gr)
        worst = max(e_gpu, key=e_gpu.get)
        med = lambda d: sorted(d.values())[len(d) // 2]
        print(f"[C2 step {mode}] oracle steps (fp32 + fp64) {t_ref:.1f}s; loss gpu {float(loss.detach()):.7f} fp32-ref {float(loss32):.7f} "
              f"fp64-ref {float(loss64):.7f}; pred vs fp64: gpu {relerr(pred, pred64):.2e}, fp32-ref {relerr(pred32, pred64):.2e}; gradients vs "
              f"fp64 (max-element error / max|grad|): gpu worst {worst} {e_gpu[worst]:.2e} (fp32-ref there {e_cpu[worst]:.2e}), gpu median "
              f"{med(e_gpu):.2e}, fp32-ref median {med(e_cpu):.2e}; gpu norm error worst {max(n_gpu.values()):.2e} median {med(n_gpu):.2e}")
        if mode == "fp32":
            assert abs(float(loss) - float(loss64)) <= 1e-5 * abs(float(loss64))
            assert relerr(pred, pred64) <= 1e-4
            for k in e_gpu:
                assert e_gpu[k] <= 2.0 * e_cpu[k] + 2e-4, (k, e_gpu[k], e_cpu[k])
        else:
            assert abs(float(loss) - float(loss64)) <= 2e-3 * abs(float(loss64))
            assert relerr(pred, pred64) <= 1e-2
            assert max(n_gpu.values()) <= 8e-2, max(n_gpu, key=n_gpu.get)
    finally:
        ops.set_precision(old)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_bulk_imputation_hidden_256_matches_oracle(mode):
    """BASELINE config 5 (inference-only bulk imputation, 256-d hidden; inference.py:140-159, advanced_visualizations.py:439-461):
    eval-mode predictions for the held-out (val + test) pairs and for every never-measured pair of a patient range, d = 256,
    against the oracle on a graph with both gate branches."""
    pkg, G, ops, M, T, MET = _mods()
    dev = torch.device("cuda:0")
    old = ops.PRECISION
    ops.set_precision(mode)
    try:
        spec = pkg.synth.GraphSpec("c5-small", 3000, 50, 114, 100, 90000, 8000, 24000, 0.1, hidden_dim=256)
        g = pkg.synth.make_graph(spec, seed=6)
        counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
        ets = list(g.edge_types)
        sd = R.init_state(counts, ets, hidden=256, seed=21)
        for k in list(sd):                       # non-trivial running statistics (eval mode uses them)
            if k.endswith("running_mean"):
                sd[k] = torch.randn_like(sd[k]) * 0.1
            if k.endswith("running_var"):
                sd[k] = torch.rand_like(sd[k]) + 0.5
        cfg = _cfg(0.2)
        cfg["model"]["hidden_dim"] = 256
        model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
        gd = pkg.synth.make_graph(spec, seed=6).to(dev)
        model._init_embeddings(gd)
        model.load_state_dict(sd)
        model.eval()
        ei = g["patient", "has_lab", "lab"].edge_index
        _, va, te = R.split_masks(ei.shape[1])
        held = va | te
        pi, li = ei[0][held], ei[1][held]
        with torch.no_grad():
            pred = model.predict_lab_values(gd, pi.to(dev), li.to(dev))
            ref = R.predict_lab_values(sd, counts, ets, g.edge_index_dict, pi, li, False)
            some = torch.arange(100, 160)
            mp, ml, mv = model.impute_missing(gd, some.to(dev))
            ref_m = R.predict_lab_values(sd, counts, ets, g.edge_index_dict, mp.cpu(), ml.cpu(), False)
        deg = torch.bincount(ei[0], minlength=counts["patient"])
        assert bool((deg[pi] < 6).any()) and bool((deg[pi] >= 6).any()), "both heads must be exercised"
        have = set(zip(ei[0].tolist(), ei[1].tolist()))
        assert list(zip(mp.tolist(), ml.tolist())) == [(p, l) for p in some.tolist() for l in range(counts["lab"]) if (p, l) not in have]
        tol = 1e-4 if mode == "fp32" else 1e-2
        e1, e2 = relerr(pred, ref), relerr(mv, ref_m)
        print(f"[config 5, d=256, {mode}] held-out pairs {pi.numel()}: max-rel {e1:.2e}; never-measured pairs {mp.numel()}: max-rel {e2:.2e}")
        assert e1 <= tol and e2 <= tol
    finally:
        ops.set_precision(old)


def test_reference_call_sequence_with_host_indices(tf32_mode):
    """The reference Trainer's call sequence (train.py:210-219, 347-392) against the drop-in on the GPU: model and graph moved to
    the device, Adam built before the lazy tables exist (N2), and -- as the reference's EdgeMasker does -- the pair indices, the
    targets and the supervision mask left on the HOST (train.py:86,173).  predict_lab_values stages the index lists itself;
    loss / backward / optimizer step follow train.py:364-390 with the two device moves INTEGRATION.md section 1 lists."""
    pkg, G, ops, M, T, MET = _mods()
    dev = torch.device("cuda:0")
    g = pkg.synth.make_graph("C1", seed=42)
    cfg = _cfg(0.2)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None)
    model = model.to(dev)                                                   # train.py:210
    data = pkg.synth.make_graph("C1", seed=42).to(dev)                      # train.py:211
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)  # train.py:219,255-260 (torch's own Adam)
    assert sum(p.numel() for grp in opt.param_groups for p in grp["params"]) == 483970
    masker = T.EdgeMasker(g, 0.7, 0.15, 0.15, 0.2, 42)                      # host tensors, like the reference's
    losses = []
    for epoch in range(3):
        model.train()                                                       # train.py:347
        ei, ev, _, sup = masker.get_masked_data("train", seed=5 + epoch)    # train.py:350 (all on the host)
        assert not ei.is_cuda and not sup.is_cuda
        opt.zero_grad()                                                     # train.py:356
        pred = model.predict_lab_values(data, ei[0], ei[1])                 # train.py:358-362: HOST index tensors
        assert pred.is_cuda and pred.shape == (ei.shape[1],)
        p_sup, t_sup = pred[sup.to(dev)], ev[sup].to(dev)                   # train.py:366-368
        loss = (p_sup - t_sup).abs().mean()                                 # train.py:380 (mae, unweighted branch)
        loss.backward()                                                     # train.py:389
        opt.step()                                                          # train.py:390
        losses.append(float(loss))
    assert all(l == l and l < 10 for l in losses) and losses[-1] <= losses[0] * 1.05
    assert model.embeddings["patient"].weight.grad is not None               # N2: tables get gradients but are not optimised

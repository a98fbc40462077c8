"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through libb2g's C ABI
(ctypes) and is compared with the CPU oracle (oracle/hetero_rgcn_ref.py) or with the golden vectors the
unmodified reference produced (tests/golden/).

Tolerances (BASELINE.json north_star): bit-exact for CSR / degrees / masks; <= 1e-5 relative for fp32
aggregation; fp32 dense layers <= 1e-4 relative (different summation order only)."""
import importlib

import pytest
import torch

from conftest import golden_graph
from oracle import hetero_rgcn_ref as R

pytestmark = pytest.mark.gpu

PKG = "multi-modal-gnn_b200"


def _mods():
    return (importlib.import_module(PKG), importlib.import_module(PKG + ".graph"), importlib.import_module(PKG + ".ops"),
            importlib.import_module(PKG + ".model"), importlib.import_module(PKG + ".trainer"), importlib.import_module(PKG + "._lib"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _exact_fp32_by_default():
    """Parity tests run the exact-fp32 kernels unless a test opts into the tensor-core (TF32) path."""
    ops = importlib.import_module(PKG + ".ops")
    old = ops.PRECISION
    ops.set_precision("fp32")
    yield
    ops.set_precision(old)


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------------
# (a) CSR / degrees / gate: bit-exact
# ------------------------------------------------------------------------------------------------------
def _check_csr(G, key, val, n_rows, n_vals, dev):
    csr = G.CSR(key.to(dev), val.to(dev), n_rows, n_vals)
    order = torch.argsort(key, stable=True)
    deg = torch.bincount(key, minlength=n_rows)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64)
    rowptr[1:] = deg.cumsum(0)
    assert torch.equal(csr.rowptr.cpu().long(), rowptr)
    e = key.numel()
    if e:
        assert torch.equal(csr.eid.cpu().long()[:e], order)
        assert torch.equal(csr.col.cpu().long()[:e], val[order])
    assert csr.deg.dtype == torch.int64 and torch.equal(csr.deg.cpu(), deg)
    torch.testing.assert_close(csr.inv_deg.cpu(), 1.0 / deg.clamp(min=1).float(), rtol=0, atol=0)
    return csr


@pytest.mark.parametrize("e,n_rows,n_vals", [(0, 7, 5), (1, 1, 1), (33, 5, 9), (4097, 50, 300), (100_000, 1834, 50),
                                             (300_000, 70_000, 200), (2_000_000, 300_000, 160), (50_000, 1 << 20, 3)])
def test_csr_build_bit_exact(e, n_rows, n_vals, dev):
    _, G, *_ = _mods()
    g = torch.Generator().manual_seed(e + n_rows)
    key = torch.randint(0, n_rows, (e,), generator=g)
    val = torch.randint(0, n_vals, (e,), generator=g)
    if e > 100:  # leave some rows empty and one row heavy
        key[key % 7 == 3] = 0
    _check_csr(G, key, val, n_rows, n_vals, dev)


def test_csr_rejects_out_of_range(dev):
    pkg, G, *_ = _mods()
    key = torch.tensor([0, 1, 5], device=dev)
    val = torch.tensor([0, 0, 0], device=dev)
    with pytest.raises(RuntimeError):
        G.CSR(key, val, 5, 1)
    g = pkg.synth.make_graph("tiny").to(dev)
    g["patient", "has_lab", "lab"].edge_index[1, 0] = 10_000
    with pytest.raises(ValueError):
        G.GraphIndex(g)


def test_graph_index_degrees_and_gate(dev):
    pkg, G, ops, M, T, L = _mods()
    g = pkg.synth.make_graph("C1")
    gi = G.GraphIndex(pkg.synth.make_graph("C1").to(dev))
    ei = g["patient", "has_lab", "lab"].edge_index
    deg = torch.bincount(ei[0], minlength=1834)
    assert torch.equal(gi.patient_lab_degree.cpu(), deg)
    # forward CSR of has_lab_rev == transposed CSR of has_lab (same edge set, SURVEY.md section 3.5 iii)
    a, b = gi.relations[("lab", "has_lab_rev", "patient")].by_dst, gi.relations[("patient", "has_lab", "lab")].by_src
    assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.col, b.col)
    pi = ei[0][::7].contiguous()
    low = torch.empty(pi.numel(), dtype=torch.uint8, device=dev)
    L.check(L.load().b2g_degree_gate(gi.patient_lab_degree.data_ptr(), pi.to(dev).data_ptr(), pi.numel(), 6, low.data_ptr(), None))
    assert torch.equal(low.cpu().bool(), deg[pi] < 6)
    for et, rel in gi.relations.items():
        e = g[et].edge_index.shape[1]
        assert int(rel.by_dst.rowptr[-1]) == e and int(rel.by_src.rowptr[-1]) == e
        assert int(rel.by_dst.deg.sum()) == e


# ------------------------------------------------------------------------------------------------------
# (b) message passing
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec,d", [("tiny", 128), ("C1", 128), ("C1", 64), ("tiny", 256), ("tiny", 32)])
def test_mean_aggregate_fwd_bwd(spec, d, dev):
    pkg, G, ops, *_ = _mods()
    g = pkg.synth.make_graph(spec)
    gi = G.GraphIndex(pkg.synth.make_graph(spec).to(dev))
    gen = torch.Generator().manual_seed(0)
    for et in g.edge_types:
        rel = gi.relations[et]
        x = torch.randn(rel.n_src, d, generator=gen)
        go = torch.randn(rel.n_dst, d, generator=gen)
        xr = x.clone().double().requires_grad_(True)
        ref = R.mean_aggregate(xr, g[et].edge_index, rel.n_dst)
        ref.backward(go.double())
        xd = x.to(dev).requires_grad_(True)
        out = ops.MeanAggFn.apply(xd, rel)
        out.backward(go.to(dev))
        assert relerr(out, ref.detach()) <= 1e-5, et
        assert relerr(xd.grad, xr.grad) <= 1e-5, et
        out2 = ops.MeanAggFn.apply(xd, rel)
        assert torch.equal(out, out2), "aggregation must be run-to-run bit-identical"


def test_aggregate_isolated_destinations_are_zero(dev):
    pkg, G, ops, *_ = _mods()
    g = pkg.synth.make_graph("C1").to(dev)
    gi = G.GraphIndex(g)
    rel = gi.relations[("diagnosis", "has_diagnosis_rev", "patient")]
    iso = (rel.by_dst.deg == 0)
    assert int(iso.sum()) > 0
    out = ops.MeanAggFn.apply(torch.randn(rel.n_src, 128, device=dev), rel)
    assert float(out[iso].abs().max()) == 0.0


def test_aggregate_linearity_and_ones_at_bench_size(dev):
    """Size-independent properties at the benchmark configuration (C2)."""
    pkg, G, ops, *_ = _mods()
    g = pkg.synth.make_graph("C2").to(dev)
    gi = G.GraphIndex(g)
    for et in [("lab", "has_lab_rev", "patient"), ("patient", "has_lab", "lab"), ("patient", "has_medication", "medication")]:
        rel = gi.relations[et]
        for csr in (rel.by_dst, rel.by_src):
            e = csr.n_edges
            assert int(csr.rowptr[-1]) == e
            assert bool((csr.rowptr[1:] >= csr.rowptr[:-1]).all())
            assert torch.equal(torch.sort(csr.eid[:e].long())[0], torch.arange(e, device=dev))
            interior = torch.ones(e, dtype=torch.bool, device=dev)
            interior[csr.rowptr[:-1][csr.deg > 0].long()] = False
            assert bool((csr.eid[1:e][interior[1:]] > csr.eid[: e - 1][interior[1:]]).all()), "rows must keep edge order"
        ones = torch.ones(rel.n_src, 128, device=dev)
        m = ops.MeanAggFn.apply(ones, rel)
        has = rel.by_dst.deg > 0
        assert float((m[has] - 1).abs().max()) <= 1e-5 and float(m[~has].abs().sum()) == 0.0
        a, b = torch.randn(rel.n_src, 128, device=dev), torch.randn(rel.n_src, 128, device=dev)
        lhs = ops.MeanAggFn.apply(2.0 * a - 3.0 * b, rel)
        rhs = 2.0 * ops.MeanAggFn.apply(a, rel) - 3.0 * ops.MeanAggFn.apply(b, rel)
        assert float((lhs - rhs).abs().max()) <= 2e-5 * float(rhs.abs().max().clamp_min(1))


# ------------------------------------------------------------------------------------------------------
# (d) dense, BN, L2 norm, dropout, loss
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n,bias", [(1, 128, 128, True), (300, 128, 128, True), (1834, 128, 64, False), (5000, 64, 32, True),
                                        (5000, 32, 1, True), (777, 256, 64, True), (129, 128, 128, False), (4099, 256, 256, True)])
def test_linear_fwd_bwd(m, k, n, bias, dev):
    *_, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(m + n)
    x, w = torch.randn(m, k, generator=gen), torch.randn(n, k, generator=gen) / k ** 0.5
    b = torch.randn(n, generator=gen) if bias else None
    go = torch.randn(m, n, generator=gen)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    br = b.double().requires_grad_(True) if bias else None
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(go.double())
    xd, wd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
    bd = b.to(dev).requires_grad_(True) if bias else None
    y = ops.linear(xd, wd, bd)
    y.backward(go.to(dev))
    assert relerr(y, yr.detach()) <= 1e-5
    assert relerr(xd.grad, xr.grad) <= 1e-5
    assert relerr(wd.grad, wr.grad) <= 2e-5
    if bias:
        assert relerr(bd.grad, br.grad) <= 2e-5


@pytest.mark.parametrize("m,d,act,p", [(50, 128, 1, 0.0), (1834, 128, 1, 0.0), (20000, 64, 1, 0.0), (300, 256, 0, 0.0),
                                       (1834, 128, 2, 0.0), (1834, 128, 3, 0.0), (1834, 128, 1, 0.2)])
def test_batchnorm_act_dropout(m, d, act, p, dev):
    *_, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(m)
    x = torch.randn(m, d, generator=gen) * 2 + 0.5
    gamma, beta = 1 + 0.1 * torch.randn(d, generator=gen), 0.1 * torch.randn(d, generator=gen)
    go = torch.randn(m, d, generator=gen)
    seed, sid = 12345, 3
    mask = ops.dropout_mask(m * d, p, seed, sid, dev).cpu().view(m, d).double() if p > 0 else None
    if p > 0:
        keep = float((mask > 0).double().mean())
        assert abs(keep - (1 - p)) < 0.01 and abs(float(mask.max()) - 1 / (1 - p)) < 1e-6
    acts = {0: lambda t: t, 1: torch.relu, 2: lambda t: torch.nn.functional.leaky_relu(t, 0.01), 3: torch.nn.functional.elu}
    for training in (True, False):
        rm, rv = torch.zeros(d) + 0.1, torch.ones(d) * 1.5
        rm_d, rv_d = rm.clone().to(dev), rv.clone().to(dev)
        xr, gr, br = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
        rm_r, rv_r = rm.double(), rv.double()
        yr = acts[act](torch.nn.functional.batch_norm(xr, rm_r, rv_r, gr, br, training, 0.1, 1e-5))
        if training and p > 0:
            yr = yr * mask
        yr.backward(go.double())
        xd, gd, bd = x.to(dev).requires_grad_(True), gamma.to(dev).requires_grad_(True), beta.to(dev).requires_grad_(True)
        y = ops.BNActDropFn.apply(xd, gd, bd, rm_d, rv_d, training, act, p, seed, sid, 1e-5, 0.1)
        y.backward(go.to(dev))
        assert relerr(y, yr.detach()) <= 2e-5
        assert relerr(xd.grad, xr.grad) <= 1e-4
        assert relerr(gd.grad, gr.grad) <= 1e-4 and relerr(bd.grad, br.grad) <= 1e-4
        assert relerr(rm_d, rm_r) <= 1e-6 and relerr(rv_d, rv_r) <= 1e-6


@pytest.mark.parametrize("m,d", [(1, 128), (1834, 128), (999, 64), (100, 256)])
def test_l2norm(m, d, dev):
    *_, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(d)
    x = torch.randn(m, d, generator=gen)
    x[0] = 0.0                      # clamp branch: x / max(0, eps) = 0
    go = torch.randn(m, d, generator=gen)
    xr = x.double().requires_grad_(True)
    yr = torch.nn.functional.normalize(xr, p=2.0, dim=1, eps=1e-12)
    yr.backward(go.double())
    xd = x.to(dev).requires_grad_(True)
    y = ops.L2NormFn.apply(xd, 1e-12)
    y.backward(go.to(dev))
    assert relerr(y, yr.detach()) <= 1e-6
    assert float(y[0].abs().max()) == 0.0
    if m > 1:
        assert relerr(xd.grad[1:], xr.grad[1:]) <= 1e-5


@pytest.mark.parametrize("kind", ["mae", "mse", "huber"])
def test_weighted_loss(kind, dev):
    *_, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(1)
    m, nl = 43038, 50
    p, t = torch.randn(m, generator=gen), torch.randn(m, generator=gen)
    lab = torch.randint(0, nl, (m,), generator=gen)
    w = torch.rand(nl, generator=gen) + 0.5
    sup = torch.rand(m, generator=gen) < 0.2
    pr = p.double().requires_grad_(True)
    ref = R.weighted_loss(pr, t.double(), lab, w.double(), sup, kind)
    ref.backward()
    pd = p.to(dev).requires_grad_(True)
    if kind == "huber":
        loss = ops.weighted_loss(pd, t.to(dev), None, None, sup.to(dev), kind)
    else:
        loss = ops.weighted_loss(pd, t.to(dev), lab.to(dev), w.to(dev), sup.to(dev), kind)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-6 * abs(float(ref))
    assert relerr(pd.grad, pr.grad) <= 1e-6
    # unweighted, all pairs == compute_regression_loss (model.py:579-612)
    l2 = M.compute_regression_loss(p.to(dev), t.to(dev), kind)
    assert abs(float(l2) - float(R.regression_loss(p.double(), t.double(), kind))) <= 1e-6
    with pytest.raises(ValueError):
        M.compute_regression_loss(p.to(dev), t.to(dev), "nope")


# ------------------------------------------------------------------------------------------------------
# whole path against golden vectors of the unmodified reference
# ------------------------------------------------------------------------------------------------------
def _model_from_golden(blob, dev, dropout=0.0, state=None):
    pkg, G, ops, M, T, L = _mods()
    counts, ets, eid, attr = golden_graph(blob)
    g = pkg.HeteroGraph()
    for nt, n in counts.items():
        g[nt].num_nodes = n
    for et in ets:
        g[et].edge_index = eid[et].to(dev)
    g["patient", "has_lab", "lab"].edge_attr = attr.to(dev)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": dropout, "use_batch_norm": True,
                     "activation": "relu"}}
    model = M.build_model(cfg, (list(counts), ets), None).to(dev)
    model._init_embeddings(g)
    model.load_state_dict(state if state is not None else blob["state_before"])
    return model, g, counts, ets, eid, attr


@pytest.mark.parametrize("fixture", ["golden_tiny_mae", "golden_tiny_mse", "golden_c1"])
def test_train_step_matches_reference_golden(fixture, request, dev):
    pkg, G, ops, M, T, L = _mods()
    blob = request.getfixturevalue(fixture)
    model, g, counts, ets, eid, attr = _model_from_golden(blob, dev)
    assert sum(p.numel() for n, p in model.named_parameters() if not n.startswith("embeddings.")) == 483970   # KA-1
    ei = eid[("patient", "has_lab", "lab")]
    tr = blob["split"]["train"]
    pi, li, tgt = ei[0][tr].to(dev), ei[1][tr].to(dev), attr[tr].squeeze(-1).to(dev)
    w = T.compute_lab_weights(li, tgt, counts["lab"])
    torch.testing.assert_close(w.cpu(), blob["lab_weights"], rtol=1e-5, atol=1e-7)
    model.train()
    pred = model.predict_lab_values(g, pi, li)
    loss = ops.weighted_loss(pred, tgt, li, w, blob["sup_mask"].to(dev), blob["loss_fn"])
    loss.backward()
    assert relerr(pred, blob["pred_train"]) <= 5e-5
    assert abs(float(loss) - blob["loss_train"]) <= 2e-5 * abs(blob["loss_train"])
    params = dict(model.named_parameters())
    none_keys = sorted(k for k, p in params.items() if p.grad is None)
    assert none_keys == blob["grad_is_none"]                                  # N8 dead branches
    for k, gref in blob["grads"].items():
        if k in ("patient_transform.0.bias", "patient_transform.4.bias") or k.endswith("lin_l.bias"):
            assert float(params[k].grad.abs().max()) < 1e-6, k                # exact zero up to rounding
            continue
        assert relerr(params[k].grad, gref) <= 2e-4, k
    for k, gn in blob["grad_norm"].items():
        if k in ("patient_transform.0.bias", "patient_transform.4.bias") or k.endswith("lin_l.bias"):
            continue
        assert abs(float(params[k].grad.double().norm()) - gn) <= 2e-4 * gn, k
    sd = model.state_dict()
    for k, v in blob["after_buffers"].items():                                # N3: patient MLP BNs updated twice
        if v.dtype == torch.long:
            assert int(sd[k]) == int(v), k
        else:
            assert relerr(sd[k], v) <= 1e-5, k


def test_eval_products_match_reference_golden(golden_tiny_mae, dev):
    pkg, G, ops, M, T, L = _mods()
    blob = golden_tiny_mae
    state = dict(blob["state_before"])
    state.update(blob["after_buffers"])
    model, g, counts, ets, eid, attr = _model_from_golden(blob, dev, dropout=0.2, state=state)
    model.eval()
    ei = eid[("patient", "has_lab", "lab")]
    with torch.no_grad():
        enc = model.encode_nodes(g)
        fwd = model(g)
        va = blob["split"]["val"]
        pred_val = model.predict_lab_values(g, ei[0][va].to(dev), ei[1][va].to(dev))
        pred_all = model.predict_lab_values(g, ei[0].to(dev), ei[1].to(dev))
        lv = M.compute_regression_loss(pred_val, attr[va].squeeze(-1).to(dev), "mae")
    for nt in counts:
        assert relerr(enc[nt], blob["encode_eval"][nt]) <= 2e-5, nt
        assert relerr(fwd[nt], blob["forward_eval"][nt]) <= 5e-5, nt
    assert relerr(pred_val, blob["pred_val"]) <= 5e-5
    assert relerr(pred_all, blob["pred_all"]) <= 5e-5
    assert abs(float(lv) - blob["eval_loss_val"]) <= 2e-5 * abs(blob["eval_loss_val"])
    # the gate really split this call across both heads
    low = blob["degree"][ei[0]] < 6
    assert 0 < int(low.sum()) < low.numel()


def test_dropout_training_step_with_replayed_masks(dev):
    """dropout > 0: device Philox masks are replayed into the CPU oracle (quirk N3: two encodes)."""
    pkg, G, ops, M, T, L = _mods()
    g = pkg.synth.make_graph("tiny", seed=9)
    gd = pkg.synth.make_graph("tiny", seed=9).to(dev)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = list(g.edge_types)
    sd = R.init_state(counts, ets, seed=5)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2, "use_batch_norm": True,
                     "activation": "relu"}}
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    model._init_embeddings(gd)
    model.load_state_dict(sd)
    model.train()
    ei = g["patient", "has_lab", "lab"].edge_index
    tgt = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
    pi, li = ei[0], ei[1]
    torch.manual_seed(77)
    pred = model.predict_lab_values(gd, pi.to(dev), li.to(dev))
    streams = model._last_streams
    tags = dict(streams.log)
    assert "init.enc.drop0" in tags and "fwd.enc.drop0" in tags and "edge_predictor.drop1" in tags

    def mask_fn(tag, x):
        return ops.dropout_mask(x.numel(), 0.2, streams.seed, tags[tag], dev).cpu().view_as(x).to(x.dtype)

    sd_ref = {k: v.clone() for k, v in sd.items()}
    ref = R.predict_lab_values(sd_ref, counts, ets, g.edge_index_dict, pi, li, True, 0.2, mask_fn=mask_fn)
    assert relerr(pred, ref) <= 1e-4
    after = model.state_dict()
    for k in ("patient_transform.1.running_mean", "patient_transform.5.running_var", "batch_norms.1.lab.running_var"):
        assert relerr(after[k], sd_ref[k]) <= 1e-5, k
    assert int(after["patient_transform.1.num_batches_tracked"]) == 2
    # statistical check of the drop rate on a large site
    mk = ops.dropout_mask(1 << 22, 0.2, 99, 0, dev)
    assert abs(float((mk == 0).float().mean()) - 0.2) < 2e-3


def test_trainer_tracks_oracle_adam(dev):
    """5 optimizer steps (mse + lab weights): the CUDA Trainer vs oracle + torch.optim.Adam on the CPU."""
    pkg, G, ops, M, T, L = _mods()
    g = pkg.synth.make_graph("tiny", seed=4)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = list(g.edge_types)
    sd = R.init_state(counts, ets, seed=8)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True,
                     "activation": "relu"},
           "train": {"loss": "mse", "epochs": 5, "early_stopping_patience": 15, "optimizer": {"type": "adam", "lr": 1e-3,
                     "weight_decay": 1e-5}, "lr_scheduler": {"enabled": True, "type": "reduce_on_plateau"}}}
    model = M.build_model(cfg, (g.node_types, g.edge_types), None)
    masker = T.EdgeMasker(pkg.synth.make_graph("tiny", seed=4), 0.7, 0.15, 0.15, 0.2, 42)
    for m_, ref_m in zip((masker.train_mask, masker.val_mask, masker.test_mask), R.split_masks(masker.num_edges)):
        assert torch.equal(m_, ref_m)                                          # bit-exact splits
    trainer = T.Trainer(model, masker.data, masker, cfg, dev)                  # optimizer built before tables exist (N2)
    model._init_embeddings(trainer.data)
    model.load_state_dict(sd)
    assert sum(p.numel() for grp in trainer.optimizer.param_groups for p in grp["params"]) == 483970

    ei = g["patient", "has_lab", "lab"].edge_index
    attr = g["patient", "has_lab", "lab"].edge_attr
    tr = masker.train_mask
    pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr].squeeze(-1)
    w = R.lab_weights(li, tgt, counts["lab"])
    torch.testing.assert_close(trainer.lab_weights.cpu(), w, rtol=1e-5, atol=1e-7)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    keys = R.trainable_keys(sd_ref)
    params = [sd_ref[k].requires_grad_(True) for k in keys]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-5)
    for step in range(5):
        loss_gpu = trainer.train_epoch(seed=1000 + step)
        sup = R.supervision_mask(int(tr.sum()), 0.2, 1000 + step)
        opt.zero_grad()
        pred = R.predict_lab_values(sd_ref, counts, ets, g.edge_index_dict, pi, li, True, p_drop=0.0)
        loss = R.weighted_loss(pred, tgt, li, w, sup, "mse")
        loss.backward()
        opt.step()
        assert abs(loss_gpu - float(loss)) <= 1e-3 * abs(float(loss)), (step, loss_gpu, float(loss))
    val_gpu = trainer.validate("val")
    va = masker.val_mask
    with torch.no_grad():
        pv = R.predict_lab_values({k: v.detach() for k, v in sd_ref.items()}, counts, ets, g.edge_index_dict, ei[0][va], ei[1][va], False)
        val_ref = float(R.regression_loss(pv, attr[va].squeeze(-1), "mse"))
    assert abs(val_gpu - val_ref) <= 2e-3 * abs(val_ref)


def test_embedding_gather_and_sparse_gradient(dev):
    """nn.Embedding lookup for an arbitrary index list + its scatter-add gradient (sorted segments, no atomics)."""
    *_, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(3)
    for n, d, m in [(50, 128, 1000), (1834, 128, 43038), (300, 64, 7), (10, 256, 5000)]:
        table = torch.randn(n, d, generator=gen)
        idx = torch.randint(0, n, (m,), generator=gen)
        go = torch.randn(m, d, generator=gen)
        tr = table.double().requires_grad_(True)
        ref = torch.nn.functional.embedding(idx, tr)
        ref.backward(go.double())
        td = table.to(dev).requires_grad_(True)
        out = ops.GatherRowsFn.apply(td, idx.to(dev))
        out.backward(go.to(dev))
        assert torch.equal(out.detach().cpu(), table[idx])              # a gather is exact
        assert relerr(td.grad, tr.grad) <= 1e-5
        td.grad = None
        out = ops.GatherRowsFn.apply(td, idx.to(dev))
        out.backward(go.to(dev))
        g2 = td.grad.clone()
        td.grad = None
        ops.GatherRowsFn.apply(td, idx.to(dev)).backward(go.to(dev))
        assert torch.equal(g2, td.grad), "scatter-add gradient must be deterministic"


def test_model_rejects_cpu_and_bad_config(dev):
    pkg, G, ops, M, T, L = _mods()
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2, "use_batch_norm": True,
                     "activation": "relu"}}
    md = (pkg.synth.NODE_TYPES, pkg.synth.EDGE_TYPES)
    model = M.build_model(cfg, md, None)
    g = pkg.synth.make_graph("tiny")
    with pytest.raises(RuntimeError):
        model(g)                                             # CPU module: no CPU path
    with pytest.raises(ValueError):
        M.build_model({"model": dict(cfg["model"], architecture="nope")}, md, None)
    with pytest.raises(ValueError):
        M.build_model({"model": dict(cfg["model"], activation="nope")}, md, None)
    # reference smoke (model.py:619-660): hidden 64, edge_predictor on a [10, 128] tensor
    m64 = M.HeteroRGCN(metadata=md, hidden_dim=64, num_layers=2, dropout=0.2, patient_feature_dim=3).to(dev)
    m64.eval()
    out = m64.edge_predictor(torch.randn(10, 128, device=dev))
    assert tuple(out.shape) == (10, 1)
    # PyG >= 2.4 key format is accepted on load
    model = model.to(dev)
    model._init_embeddings(g)
    sd = {k.replace("patient__has_lab__lab", "<patient___has_lab___lab>"): v for k, v in model.state_dict().items()}
    model.load_state_dict(sd)


@pytest.mark.parametrize("m,p_drop,frac_active", [(1, 0.0, 1.0), (257, 0.0, 1.0), (5000, 0.0, 0.2), (43038, 0.2, 0.2), (3000, 0.2, 0.0)])
def test_fused_decoder_matches_oracle_head(m, p_drop, frac_active, dev):
    """csrc/decoder.cu vs the oracle's EdgeRegressionHead on cat([h_p[pi], h_l[li]]) incl. replayed dropout masks and a
    sparse upstream gradient (only 'supervised' pairs carry gradient, as in train.py:366-368)."""
    pkg, G, ops, M, T, L = _mods()
    gen = torch.Generator().manual_seed(m)
    n_p, n_l, d = 700, 37, 128
    hp, hl = torch.randn(n_p, d, generator=gen), torch.randn(n_l, d, generator=gen)
    pi, li = torch.randint(0, n_p, (m,), generator=gen), torch.randint(0, n_l, (m,), generator=gen)
    head = M.EdgeRegressionHead(2 * d, [64, 32], 1, p_drop).to(dev)
    head.train()
    sd = {"h.mlp." + k.split("mlp.")[1]: v.detach().cpu().clone() for k, v in head.state_dict().items()}
    go = torch.randn(m, generator=gen) * (torch.rand(m, generator=gen) < frac_active)
    pairs = G.PairIndex(pi.to(dev), li.to(dev), n_p, n_l)
    streams = M._DropoutStreams(p_drop > 0)
    hpd, hld = hp.to(dev).requires_grad_(True), hl.to(dev).requires_grad_(True)
    pred = head.forward_pairs(hpd, hld, pairs, streams, "h")
    pred.backward(go.to(dev))
    tags = dict(streams.log)

    def mask_fn(tag, x):
        return ops.dropout_mask(x.numel(), p_drop, streams.seed, tags[tag], dev).cpu().view_as(x).to(x.dtype)

    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    hpr, hlr = hp.double().requires_grad_(True), hl.double().requires_grad_(True)
    ref = R.edge_head(sdr, "h", torch.cat([hpr[pi], hlr[li]], 1), True, p_drop, mask_fn if p_drop > 0 else None)
    ref.backward(go.double())
    assert relerr(pred, ref) <= 2e-5
    if frac_active > 0:
        assert relerr(hpd.grad, hpr.grad) <= 1e-4 and relerr(hld.grad, hlr.grad) <= 1e-4
        for name, prm in head.named_parameters():
            assert relerr(prm.grad, sdr["h." + name].grad) <= 1e-4, name
    else:
        assert float(hpd.grad.abs().max()) == 0.0 and float(head.mlp[3].weight.grad.abs().max()) == 0.0
    # the unfused composition (generic kernels) computes the same function
    head.fused = False
    hp2, hl2 = hp.to(dev).requires_grad_(True), hl.to(dev).requires_grad_(True)
    streams2 = M._DropoutStreams(False)
    streams2.seed = streams.seed
    pred2 = head.forward_pairs(hp2, hl2, pairs, streams2, "h")
    assert relerr(pred2, pred) <= 1e-5
    # determinism
    head.fused = True
    head.zero_grad()
    hp3, hl3 = hp.to(dev).requires_grad_(True), hl.to(dev).requires_grad_(True)
    s3 = M._DropoutStreams(False)
    s3.seed = streams.seed
    p3 = head.forward_pairs(hp3, hl3, pairs, s3, "h")
    p3.backward(go.to(dev))
    assert torch.equal(p3, pred) and torch.equal(hp3.grad, hpd.grad)


# ------------------------------------------------------------------------------------------------------
# (d) tcgen05 tensor-core path (TF32 operands, fp32 accumulate): tolerance 1e-2 per BASELINE.json north_star
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,k,n,bias,acc", [(512, 128, 128, True, False), (127, 128, 128, True, False), (129, 128, 128, False, False),
                                            (46520, 128, 128, True, False), (46520, 128, 64, False, False), (5000, 64, 32, True, False),
                                            (70001, 128, 128, True, True), (3000, 256, 64, True, False), (1000, 32, 32, False, False),
                                            (20000, 64, 256, True, False)])
def test_tcgen05_linear_forward(m, k, n, bias, acc, dev):
    *_, ops, M, T, L = _mods()
    lib = L.load()
    assert lib.b2g_linear_fwd_tc_supported(m, n, k) == 1
    gen = torch.Generator().manual_seed(m + n + k)
    x, w = torch.randn(m, k, generator=gen), torch.randn(n, k, generator=gen) / k ** 0.5
    b = torch.randn(n, generator=gen) if bias else None
    y0 = torch.randn(m, n, generator=gen) if acc else None
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double() if bias else None)
    if acc:
        ref = ref + y0.double()
    xd, wd = x.to(dev), w.to(dev)
    bd = b.to(dev) if bias else None
    y = y0.to(dev).clone() if acc else torch.full((m, n), float("nan"), device=dev)
    L.check(lib.b2g_linear_fwd_tc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr() if bias else None, m, n, k, y.data_ptr(), int(acc), None))
    torch.cuda.synchronize()
    err = relerr(y, ref)
    assert err <= 3e-3, err                       # TF32: 10-bit mantissa operands
    assert err >= 1e-7 or k <= 32                 # ... and it really is not the fp32 kernel
    y2 = y0.to(dev).clone() if acc else torch.empty((m, n), device=dev)
    L.check(lib.b2g_linear_fwd_tc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr() if bias else None, m, n, k, y2.data_ptr(), int(acc), None))
    assert torch.equal(y, y2), "tensor-core path must be deterministic"


def test_tcgen05_linear_exact_on_tf32_representable_inputs(dev):
    """With operands that are exactly representable in TF32 the tensor-core result equals fp32 math up to
    accumulation order: pins the smem descriptors / swizzle / TMEM lane mapping, independent of rounding."""
    *_, ops, M, T, L = _mods()
    lib = L.load()
    gen = torch.Generator().manual_seed(0)
    m, k, n = 1000, 128, 128
    x = torch.randint(-8, 9, (m, k), generator=gen).float() / 8
    w = torch.randint(-8, 9, (n, k), generator=gen).float() / 16
    ref = x.double() @ w.double().t()
    y = torch.empty(m, n, device=dev)
    xd, wd = x.to(dev), w.to(dev)
    L.check(lib.b2g_linear_fwd_tc(xd.data_ptr(), wd.data_ptr(), None, m, n, k, y.data_ptr(), 0, None))
    assert relerr(y, ref) <= 1e-6
    assert lib.b2g_linear_fwd_tc_supported(1000, 16, 128) == 0 and lib.b2g_linear_fwd_tc_supported(1000, 128, 48) == 0


def test_tcgen05_autograd_and_training_step(dev, golden_tiny_mae, golden_c1):
    pkg, G, ops, M, T, L = _mods()
    ops.set_precision("tf32")
    gen = torch.Generator().manual_seed(5)
    m, k, n = 46520, 128, 128
    x, w, b = torch.randn(m, k, generator=gen), torch.randn(n, k, generator=gen) / k ** 0.5, torch.randn(n, generator=gen)
    go = torch.randn(m, n, generator=gen)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    torch.nn.functional.linear(xr, wr, br).backward(go.double())
    xd, wd, bd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    ops.PROFILE = []
    ops.linear(xd, wd, bd).backward(go.to(dev))
    names = [p[0] for p in ops.PROFILE]
    ops.PROFILE = None
    assert "b2g_linear_fwd_tc" in names and "b2g_linear_bwd_input_tc" in names
    assert relerr(xd.grad, xr.grad) <= 3e-3 and relerr(wd.grad, wr.grad) <= 3e-3
    assert "b2g_linear_bwd_weight_tc" in names
    # whole training step on the C1 golden with tensor-core linears: within the 1e-2 band of the reference
    blob = golden_c1
    model, g, counts, ets, eid, attr = _model_from_golden(blob, dev)
    ei = eid[("patient", "has_lab", "lab")]
    tr = blob["split"]["train"]
    pi, li, tgt = ei[0][tr].to(dev), ei[1][tr].to(dev), attr[tr].squeeze(-1).to(dev)
    model.train()
    pred = model.predict_lab_values(g, pi, li)
    loss = ops.weighted_loss(pred, tgt, li, blob["lab_weights"].to(dev), blob["sup_mask"].to(dev), blob["loss_fn"])
    loss.backward()
    assert relerr(pred, blob["pred_train"]) <= 1e-2
    assert abs(float(loss) - blob["loss_train"]) <= 1e-3 * abs(blob["loss_train"])
    params = dict(model.named_parameters())
    # gradients pass through ~10 chained TF32 products and the cancellation inside BatchNorm's backward: the deepest
    # ones (first MLP layer) carry the largest error; measured 6.7e-2 .. 7.8e-2 of max|grad| on this fixture (the CUDA path is
    # bit-reproducible, so the bound below is a margin over a fixed number, not over noise).  The claim made for this mode is the
    # final-metric one (tests/test_gpu_e2e_parity.py); in float64 terms the reference's own fp32 gradients are no closer.
    errs = {k_: relerr(params[k_].grad, blob["grads"][k_]) for k_ in blob["grads"]
            if not (k_ in ("patient_transform.0.bias", "patient_transform.4.bias") or k_.endswith("lin_l.bias"))}
    print("TF32 gradient errors (fraction of max|grad|):", {k_: round(v, 4) for k_, v in errs.items()})
    assert max(errs.values()) <= 1.2e-1, errs


@pytest.mark.parametrize("m,n,k", [(64, 128, 128), (1000, 128, 128), (46520, 128, 128), (46521, 64, 128), (30000, 128, 64),
                                   (9999, 128, 256), (5000, 256, 128), (777, 32, 128)])
def test_tcgen05_weight_gradient(m, n, k, dev):
    *_, ops, M, T, L = _mods()
    lib = L.load()
    assert lib.b2g_linear_bwd_weight_tc_supported(m, n, k) == 1
    gen = torch.Generator().manual_seed(m + n + k)
    # TF32-representable operands: the result must equal exact math up to fp32 accumulation order, which pins the
    # MN-major descriptors / swizzle / transposed write-back independently of operand rounding
    dy = (torch.randint(-8, 9, (m, n), generator=gen).float() / 8).to(dev)
    x = (torch.randint(-8, 9, (m, k), generator=gen).float() / 16).to(dev)
    ref = dy.double().cpu().t() @ x.double().cpu()
    dw = torch.full((n, k), float("nan"), device=dev)
    ws = torch.empty(lib.b2g_linear_bwd_weight_tc_ws_bytes(m, n, k), dtype=torch.uint8, device=dev)
    L.check(lib.b2g_linear_bwd_weight_tc(dy.data_ptr(), x.data_ptr(), m, n, k, dw.data_ptr(), ws.data_ptr(), ws.numel(), None))
    assert relerr(dw, ref) <= 2e-6
    # random operands: TF32 rounding only
    dy2, x2 = torch.randn(m, n, generator=gen).to(dev), torch.randn(m, k, generator=gen).to(dev)
    ref2 = dy2.double().cpu().t() @ x2.double().cpu()
    L.check(lib.b2g_linear_bwd_weight_tc(dy2.data_ptr(), x2.data_ptr(), m, n, k, dw.data_ptr(), ws.data_ptr(), ws.numel(), None))
    assert relerr(dw, ref2) <= 3e-3
    dw2 = torch.empty_like(dw)
    L.check(lib.b2g_linear_bwd_weight_tc(dy2.data_ptr(), x2.data_ptr(), m, n, k, dw2.data_ptr(), ws.data_ptr(), ws.numel(), None))
    assert torch.equal(dw, dw2)
    # bias gradient
    db = torch.empty(n, device=dev)
    ws2 = torch.empty(lib.b2g_bn_ws_bytes(n), dtype=torch.uint8, device=dev)
    L.check(lib.b2g_col_sums(dy2.data_ptr(), m, n, db.data_ptr(), ws2.data_ptr(), ws2.numel(), None))
    assert relerr(db, dy2.double().sum(0)) <= 1e-6


def test_dense_adjacency_tensor_core_aggregation(dev):
    """tf32 mode: the four products of a relation on tcgen05 with the dense adjacency, against the oracle's
    index_select/index_add mean (fp64).  Entries of the adjacency are exact; the features are rounded to TF32."""
    pkg, G, ops, M, T, L = _mods()
    ops.set_precision("tf32")
    spec = pkg.synth.GraphSpec("mid", 6000, 50, 114, 100, 200_000, 17_000, 52_000, low_degree_frac=0.05)
    g = pkg.synth.make_graph(spec, seed=1)
    gi = G.GraphIndex(pkg.synth.make_graph(spec, seed=1).to(dev))
    gen = torch.Generator().manual_seed(0)
    used = 0
    for et in g.edge_types:
        rel = gi.relations[et]
        dn = rel.dense(128)
        if dn is None:
            continue
        used += 1
        ei = g[et].edge_index
        dense_ref = torch.zeros(dn.n_big, dn.pad)
        big, small = (ei[1], ei[0]) if dn.big_is_dst else (ei[0], ei[1])
        dense_ref[big, small] = 1.0
        if dn.big_is_dst:
            dense_ref = dense_ref / torch.bincount(big, minlength=dn.n_big).clamp(min=1).float().unsqueeze(1)
        torch.testing.assert_close(dn.mat.cpu(), dense_ref, rtol=0, atol=0)      # adjacency itself: exact
        x = torch.randn(rel.n_src, 128, generator=gen)
        go = torch.randn(rel.n_dst, 128, generator=gen)
        xr = x.double().requires_grad_(True)
        ref = R.mean_aggregate(xr, ei, rel.n_dst)
        ref.backward(go.double())
        xd = x.to(dev).requires_grad_(True)
        ops.PROFILE = []
        out = ops.MeanAggFn.apply(xd, rel)
        out.backward(go.to(dev))
        names = {p[0] for p in ops.PROFILE}
        ops.PROFILE = None
        assert {"b2g_adjacency_mma_fwd", "b2g_adjacency_mma_bwd"} <= names, names
        assert relerr(out, ref.detach()) <= 2e-3, et
        assert relerr(xd.grad, xr.grad) <= 2e-3, et
    assert used >= 4, "the mid-size graph should qualify lab and medication relations in both directions"


def test_cuda_graph_step_equals_eager_step(dev):
    """Trainer.enable_cuda_graph(): replaying the captured forward+loss+backward gives bit-identical losses and
    parameters to issuing the same launches one by one (dropout 0), and fresh dropout masks per replay (dropout > 0)."""
    pkg, G, ops, M, T, L = _mods()
    g = pkg.synth.make_graph("tiny", seed=4)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    sd = R.init_state(counts, list(g.edge_types), seed=8)

    def run(use_graph, dropout):
        cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": dropout, "use_batch_norm": True,
                         "activation": "relu"},
               "train": {"loss": "mse", "epochs": 5, "early_stopping_patience": 15,
                         "optimizer": {"type": "adam", "lr": 1e-3, "weight_decay": 1e-5}, "lr_scheduler": {"enabled": False}}}
        model = M.build_model(cfg, (g.node_types, g.edge_types), None)
        masker = T.EdgeMasker(pkg.synth.make_graph("tiny", seed=4), 0.7, 0.15, 0.15, 0.2, 42)
        trainer = T.Trainer(model, masker.data, masker, cfg, dev)
        model._init_embeddings(trainer.data)
        model.load_state_dict(sd)
        if use_graph:
            trainer.enable_cuda_graph()
        losses = [trainer.train_epoch(seed=500 + i) for i in range(4)]
        return losses, {k: v.detach().clone() for k, v in model.state_dict().items()}, trainer

    le, se, _ = run(False, 0.0)
    lg, sg, tg = run(True, 0.0)
    assert le == lg, (le, lg)
    for k in se:
        assert torch.equal(se[k], sg[k]), k
    assert tg._graph is not None
    assert int(sg["patient_transform.1.num_batches_tracked"]) == 8 and int(sg["batch_norms.0.lab.num_batches_tracked"]) == 4
    # dropout > 0: every replay must see a new mask (the seed lives in device memory)
    ld, _, td = run(True, 0.2)
    assert all(torch.isfinite(torch.tensor(ld))) and len(set(ld)) == 4
    with torch.no_grad():
        pi, li = td.masker.split_rows("train")
        td.model.train()
        a = td.model.predict_lab_values(td.data, pi, li)
        td.model._seed_buffer.fill_(12345)
        b = td.model.predict_lab_values(td.data, pi, li)
        td.model._seed_buffer.fill_(12345)
        c = td.model.predict_lab_values(td.data, pi, li)
    assert not torch.equal(a, b) and torch.equal(b, c)


@pytest.mark.parametrize("m,p_drop", [(512, 0.0), (1000, 0.0), (43038, 0.2), (300000, 0.2)])
def test_fused_decoder_tensor_core_forward(m, p_drop, dev):
    """tf32 mode: k_decoder_fwd_tc (thread-written swizzled A tile + tcgen05) against the exact-fp32 SIMT kernel with
    the same dropout streams."""
    pkg, G, ops, M, T, L = _mods()
    lib = L.load()
    gen = torch.Generator().manual_seed(m)
    n_p, n_l = 5000, 160
    U, V = torch.randn(n_p, 64, generator=gen).to(dev), torch.randn(n_l, 64, generator=gen).to(dev)
    pi, li = torch.randint(0, n_p, (m,), generator=gen).to(dev), torch.randint(0, n_l, (m,), generator=gen).to(dev)
    W2, b2 = (torch.randn(32, 64, generator=gen) / 8).to(dev), torch.randn(32, generator=gen).to(dev)
    w3, b3 = torch.randn(32, generator=gen).to(dev), torch.randn(1, generator=gen).to(dev)
    ref, out = torch.empty(m, device=dev), torch.full((m,), float("nan"), device=dev)
    args = (U.data_ptr(), V.data_ptr(), pi.data_ptr(), li.data_ptr(), W2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), m, p_drop,
            777, 3, 4)
    L.check(lib.b2g_decoder_fwd(*args, ref.data_ptr(), None))
    L.check(lib.b2g_decoder_fwd_tc(*args, out.data_ptr(), None))
    assert relerr(out, ref) <= 3e-3
    out2 = torch.empty(m, device=dev)
    L.check(lib.b2g_decoder_fwd_tc(*args, out2.data_ptr(), None))
    assert torch.equal(out, out2)
    # TF32-representable operands -> equal up to accumulation order
    Uq, Vq = (torch.randint(-8, 9, (n_p, 64), generator=gen).float() / 16).to(dev), (torch.randint(-8, 9, (n_l, 64), generator=gen).float() / 16).to(dev)
    W2q = (torch.randint(-8, 9, (32, 64), generator=gen).float() / 8).to(dev)
    argsq = (Uq.data_ptr(), Vq.data_ptr(), pi.data_ptr(), li.data_ptr(), W2q.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), m, 0.0, 0, 0, 0)
    L.check(lib.b2g_decoder_fwd(*argsq, ref.data_ptr(), None))
    L.check(lib.b2g_decoder_fwd_tc(*argsq, out.data_ptr(), None))
    assert relerr(out, ref) <= 2e-6


@pytest.mark.parametrize("m,p_drop,frac", [(600, 0.0, 1.0), (43038, 0.2, 0.2), (300000, 0.2, 0.2), (5000, 0.0, 0.0)])
def test_fused_decoder_tensor_core_backward(m, p_drop, frac, dev):
    """tf32 mode: k_decoder_bwd_tc against the exact-fp32 SIMT backward (same dropout streams, same compaction)."""
    pkg, G, ops, M, T, L = _mods()
    lib = L.load()
    gen = torch.Generator().manual_seed(m + 1)
    n_p, n_l = 5000, 160
    U, V = torch.randn(n_p, 64, generator=gen).to(dev), torch.randn(n_l, 64, generator=gen).to(dev)
    pi, li = torch.randint(0, n_p, (m,), generator=gen).to(dev), torch.randint(0, n_l, (m,), generator=gen).to(dev)
    W2, b2 = (torch.randn(32, 64, generator=gen) / 8).to(dev), torch.randn(32, generator=gen).to(dev)
    w3 = torch.randn(32, generator=gen).to(dev)
    dpred = (torch.randn(m, generator=gen) * (torch.rand(m, generator=gen) < frac)).to(dev)
    ws = torch.empty(lib.b2g_decoder_bwd_ws_bytes(m), dtype=torch.uint8, device=dev)

    def run(fn):
        g = torch.zeros(m, 64, device=dev)
        flags = torch.empty(m, device=dev)
        dW2, db2, dw3, db3 = torch.empty(32, 64, device=dev), torch.empty(32, device=dev), torch.empty(32, device=dev), torch.empty(1, device=dev)
        L.check(fn(U.data_ptr(), V.data_ptr(), pi.data_ptr(), li.data_ptr(), W2.data_ptr(), b2.data_ptr(), w3.data_ptr(), dpred.data_ptr(), m,
                   p_drop, 99, 1, 2, g.data_ptr(), flags.data_ptr(), dW2.data_ptr(), db2.data_ptr(), dw3.data_ptr(), db3.data_ptr(),
                   ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        return g * flags.unsqueeze(1), flags, dW2, db2, dw3, db3

    ref, out = run(lib.b2g_decoder_bwd), run(lib.b2g_decoder_bwd_tc)
    assert torch.equal(ref[1], out[1]) and torch.equal(ref[1], (dpred != 0).float())
    if frac == 0.0:
        for a in out[2:]:
            assert float(a.abs().max()) == 0.0
        return
    # TF32 can flip the sign of a near-zero layer-2 pre-activation, which toggles that unit's ReLU derivative for the
    # pair: a few rows differ visibly by construction, so compare in norm and bound the fraction of such rows
    for name, a, b in zip(("g", "dW2", "db2", "dw3", "db3"), (out[0], *out[2:]), (ref[0], *ref[2:])):
        err = float((a - b).norm() / b.norm().clamp_min(1e-30))
        assert err <= 8e-2, (name, err)
    row_err = (out[0] - ref[0]).abs().max(1)[0]
    assert float((row_err > 5e-3 * ref[0].abs().max()).float().mean()) <= 0.02
    out2 = run(lib.b2g_decoder_bwd_tc)
    assert torch.equal(out[0], out2[0]) and torch.equal(out[2], out2[2])


# ------------------------------------------------------------------------------------------------------
# (f) peer-memory communicator: on one GPU (world = 1) the exchange degenerates to "sum of my own slice", which pins the
#     slot / parity / sequence bookkeeping and the fused BatchNorm kernels; the 2-GPU run is tools/dist_check.py
# ------------------------------------------------------------------------------------------------------
def test_peer_comm_single_rank_allreduce_and_fused_batchnorm(dev):
    import ctypes
    pkg, G, ops, M, T, L = _mods()
    lib = L.load()
    region, handle, comm = ctypes.c_void_p(), ctypes.create_string_buffer(64), ctypes.c_void_p()
    L.check(lib.b2g_comm_local_alloc(ctypes.byref(region), handle))
    L.check(lib.b2g_comm_create(0, 1, region, bytes(handle.raw), ctypes.byref(comm)))
    try:
        gen = torch.Generator().manual_seed(5)
        for n, dt in ((4, torch.float32), (512, torch.float64), (460 * 128, torch.float32), (1_000_000, torch.float32)):
            a = torch.randn(n, generator=gen, dtype=dt).to(dev)
            fn = lib.b2g_comm_allreduce_f32 if dt == torch.float32 else lib.b2g_comm_allreduce_f64
            for _ in range(3):                      # both parities, advancing sequence numbers
                out = torch.full_like(a, float("nan"))
                L.check(fn(comm, a.data_ptr(), out.data_ptr(), n, None))
                assert torch.equal(out, a)
        big = torch.zeros(int(lib.b2g_comm_max_bytes()) // 4 + 4, device=dev)
        assert lib.b2g_comm_allreduce_f32(comm, big.data_ptr(), big.data_ptr(), big.numel(), None) == -1   # over the one-shot limit
        odd = torch.zeros(6, device=dev)
        assert lib.b2g_comm_allreduce_f32(comm, odd.data_ptr(), odd.data_ptr(), 6, None) == -1             # not a multiple of 16 bytes

        m, d = 3000, 128
        x, dy = torch.randn(m, d, generator=gen).to(dev), torch.randn(m, d, generator=gen).to(dev)
        gamma, beta = torch.rand(d, generator=gen).to(dev) + 0.5, torch.randn(d, generator=gen).to(dev)
        ws = torch.empty(lib.b2g_bn_ws_bytes(d), dtype=torch.uint8, device=dev)

        def stats(sync):
            mean, rstd = torch.empty(d, device=dev), torch.empty(d, device=dev)
            rm, rv = torch.zeros(d, device=dev), torch.ones(d, device=dev)
            if sync:
                L.check(lib.b2g_bn_stats_sync(comm, x.data_ptr(), m, m, d, 1e-5, 0.1, mean.data_ptr(), rstd.data_ptr(), rm.data_ptr(),
                                              rv.data_ptr(), ws.data_ptr(), ws.numel(), None))
            else:
                L.check(lib.b2g_bn_stats(x.data_ptr(), m, d, 1e-5, 0.1, mean.data_ptr(), rstd.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                         ws.data_ptr(), ws.numel(), None))
            return mean, rstd, rm, rv

        a, b = stats(False), stats(True)
        assert all(torch.equal(u, v) for u, v in zip(a, b))
        mean, rstd = a[0], a[1]

        def bwd(sync):
            dx, dg, db, cs = torch.empty_like(x), torch.empty(d, device=dev), torch.empty(d, device=dev), torch.empty(d, device=dev)
            if sync:
                L.check(lib.b2g_bn_bwd_sync(comm, x.data_ptr(), dy.data_ptr(), m, m, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                            beta.data_ptr(), 1, 0.2, 77, 5, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), cs.data_ptr(),
                                            ws.data_ptr(), ws.numel(), None))
            else:
                L.check(lib.b2g_bn_bwd(x.data_ptr(), dy.data_ptr(), m, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                       1, 0.2, 77, 5, 1, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), cs.data_ptr(), ws.data_ptr(),
                                       ws.numel(), None))
            torch.cuda.synchronize()
            ref = dx.double().sum(0)                       # the fused column sums of dx = the preceding Linear's bias gradient
            assert float((cs.double() - ref).abs().max()) <= 1e-6 * float(dx.abs().sum(0).max())
            return dx, dg, db

        a, b = bwd(False), bwd(True)
        assert all(torch.equal(u, v) for u, v in zip(a, b))
        torch.cuda.synchronize()
        assert lib.b2g_comm_error(comm) == 0
    finally:
        L.check(lib.b2g_comm_destroy(comm))


# ------------------------------------------------------------------------------------------------------
# (g) one-launch Adam == torch.optim.Adam (train.py:255-260), including skipped (grad is None) parameters,
#     a parameter that starts receiving gradients later, and state_dict exchange in both directions
# ------------------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_adam(dev):
    O = importlib.import_module(PKG + ".optim")
    gen = torch.Generator().manual_seed(3)
    shapes = [(128, 128), (128,), (64, 256), (1, 32), (1,), (4099,), (460, 128), (7, 3)]
    mine = [torch.randn(*s, generator=gen).to(dev).requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in mine]
    a, b = O.FusedAdam(mine, lr=1e-3, weight_decay=1e-5), torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-5)
    for step in range(6):
        for i, (p, q) in enumerate(zip(mine, ref)):
            if i == 3 or (i == 5 and step < 2):       # never / late gradients
                p.grad = q.grad = None
                continue
            g = torch.randn(*shapes[i], generator=gen).to(dev) * (10.0 ** (i - 4))
            p.grad, q.grad = g.clone(), g.clone()
        if step == 3:
            b.param_groups[0]["lr"] = a.param_groups[0]["lr"] = 5e-4          # what ReduceLROnPlateau does
        a.step()
        b.step()
    assert a.launches == 6 + 4                        # one launch per step, plus one for the late cohort
    for i, (p, q) in enumerate(zip(mine, ref)):
        assert relerr(p, q) <= 2e-6, (i, relerr(p, q))
        if i == 3:
            assert len(a.state[p]) == 0 and len(b.state[q]) == 0
        else:
            assert float(a.state[p]["step"]) == float(b.state[q]["step"])
            assert relerr(a.state[p]["exp_avg"], b.state[q]["exp_avg"]) <= 2e-6
            assert relerr(a.state[p]["exp_avg_sq"], b.state[q]["exp_avg_sq"]) <= 2e-6
    # checkpoints move between the two implementations
    b2 = torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-5)
    b2.load_state_dict(a.state_dict())
    a2 = O.FusedAdam(mine, lr=1e-3, weight_decay=1e-5)
    a2.load_state_dict(b.state_dict())
    for p, q in zip(mine, ref):
        if p.grad is not None:
            g = torch.randn(p.shape, generator=gen).to(dev)
            p.grad, q.grad = g.clone(), g.clone()
    a2.step()
    b2.step()
    for i, (p, q) in enumerate(zip(mine, ref)):
        assert relerr(p, q) <= 1e-5, (i, relerr(p, q))      # the two runs entered this step 2e-6 apart


# ------------------------------------------------------------------------------------------------------
# inference: eval-mode embedding cache (inference.py:92-159 calls predict_lab_values per patient) and the bulk
# "all never-measured pairs" imputer (inference.py:140-159), against the reference's own eval-mode predictions
# ------------------------------------------------------------------------------------------------------
def test_embedding_cache_and_bulk_imputation(golden_tiny_mae, dev):
    pkg, G, ops, M, T, L = _mods()
    blob = golden_tiny_mae
    state = dict(blob["state_before"])
    state.update(blob["after_buffers"])
    model, g, counts, ets, eid, attr = _model_from_golden(blob, dev, dropout=0.2, state=state)
    model.eval()
    ei = eid[("patient", "has_lab", "lab")]
    pi, li = ei[0].to(dev), ei[1].to(dev)
    lib = L.load()
    with torch.no_grad():
        plain = model.predict_lab_values(g, pi, li)
        model.enable_embedding_cache()
        lib.b2g_reset_launch_count()
        first = model.predict_lab_values(g, pi, li)
        n_first = lib.b2g_launch_count()
        lib.b2g_reset_launch_count()
        second = model.predict_lab_values(g, pi, li)
        n_second = lib.b2g_launch_count()
        assert torch.equal(plain, first) and torch.equal(first, second)
        assert n_second < n_first / 4                     # only the decoder ran the second time
        assert relerr(second, blob["pred_all"]) <= 5e-5   # == the unmodified reference's eval-mode predictions
        # a modified parameter invalidates the cache
        model.edge_predictor.mlp[0].bias.add_(0.25)
        third = model.predict_lab_values(g, pi, li)
        model.enable_embedding_cache(False)
        fresh = model.predict_lab_values(g, pi, li)
        assert torch.equal(third, fresh) and not torch.equal(third, second)
        model.edge_predictor.mlp[0].bias.sub_(0.25)
        # training mode never uses the cache
        model.enable_embedding_cache()
        model.train()
        with torch.enable_grad():
            model.predict_lab_values(g, pi, li).sum().backward()
        model.eval()

        # bulk imputation of every never-measured pair of a patient subset, and of all patients
        n_p, n_l = counts["patient"], counts["lab"]
        have = set(zip(ei[0].tolist(), ei[1].tolist()))
        some = torch.tensor([0, 5, 17, 17, n_p - 1])
        mp, ml, mv = model.impute_missing(g, some)
        want = [(p, l) for p in sorted(set(some.tolist())) for l in range(n_l) if (p, l) not in have]
        assert list(zip(mp.tolist(), ml.tolist())) == want
        assert torch.equal(mv, model.predict_lab_values(g, mp, ml))
        ap, al, av = model.impute_missing(g, chunk_pairs=1000)        # several chunks
        assert ap.numel() == n_p * n_l - len(have) and av.shape == ap.shape and bool(torch.isfinite(av).all())
        # against the CPU oracle on the same pairs
        sd = {k: v.cpu() for k, v in model.state_dict().items()}
        ref = R.predict_lab_values(sd, counts, ets, eid, mp.cpu(), ml.cpu(), False)
        assert relerr(mv, ref) <= 5e-5
    with pytest.raises(L.B2GError):
        model.train()
        model.impute_missing(g, some)


# ------------------------------------------------------------------------------------------------------
# (h) on-device evaluation metrics == the reference's evaluate.py functions (golden) and the numpy oracle
# ------------------------------------------------------------------------------------------------------
def test_eval_metrics_match_reference(dev):
    import math
    import os
    from oracle import eval_metrics_ref as E
    MX = importlib.import_module(PKG + ".metrics")
    blob = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.pt"), weights_only=False)
    p, t, lab = blob["pred"].to(dev), blob["target"].to(dev), blob["lab"].to(dev)

    def close(a, b, tol=1e-5):
        return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= tol * max(1.0, abs(b))

    for tag, wins in (("raw", False), ("winsorized", True)):
        res = MX.evaluate_predictions(p, t, lab, blob["n_labs"], winsorize=wins, return_winsorized=True)
        want = blob[f"overall_{tag}"]
        assert all(close(res["overall"][k], want[k]) for k in ("mae", "rmse", "r2", "mape")), (tag, res["overall"], want)
        wrows = {r["lab_index"]: r for r in blob[f"per_lab_{tag}"]}
        assert sorted(r["lab_index"] for r in res["per_lab"]) == sorted(wrows)
        maes = [r["mae"] for r in res["per_lab"]]
        assert maes == sorted(maes)
        for r in res["per_lab"]:
            w = wrows[r["lab_index"]]
            assert r["num_samples"] == w["num_samples"]
            assert all(close(r[k], w[k]) for k in ("mae", "rmse", "r2", "mape")), (tag, r, w)
        if wins:
            # the cap bounds come from fp64 moments here and float32 moments in numpy: a residual that sits on a bound may
            # be counted differently, the capped VALUES agree to float32 rounding
            assert abs(res["num_capped"] - blob["num_capped"]) <= 3
            assert float((res["predictions"].cpu() - blob["pred_winsorized"]).abs().max()) <= 2e-5
        else:
            assert res["num_capped"] == 0 and torch.equal(res["predictions"], p)
    # stratified analysis (evaluate.py:237-341) against what the UNMODIFIED reference functions returned (golden)
    ei = blob["has_lab_edge_index"].to(dev)
    deg = torch.bincount(ei[0], minlength=blob["n_patients"])
    cnt = torch.bincount(ei[1], minlength=blob["n_labs"])
    res = MX.evaluate_predictions(p, t, lab, blob["n_labs"], patient_indices=blob["patient"].to(dev), patient_lab_degree=deg, lab_counts=cnt)
    for key in ("by_patient_degree", "by_lab_frequency"):
        got, want = res["stratified"][key], blob[key]
        assert list(got) == list(want), (key, list(got), list(want))
        for grp in want:
            assert got[grp]["num_samples"] == want[grp]["num_samples"], (key, grp)
            assert all(close(got[grp][k], want[grp][k], 2e-5) for k in ("mae", "rmse", "r2", "mape")), (key, grp, got[grp], want[grp])
    # run-to-run identical (fixed-order reductions)
    a = MX.evaluate_predictions(p, t, lab, blob["n_labs"])["records"]
    b = MX.evaluate_predictions(p, t, lab, blob["n_labs"])["records"]
    assert torch.equal(a, b)
    # degenerate inputs: no pairs at all
    e = torch.zeros(0, device=dev)
    res = MX.evaluate_predictions(e, e, torch.zeros(0, dtype=torch.int64, device=dev), 5)
    assert res["per_lab"] == [] and math.isnan(res["overall"]["mae"])

"""GPU tests of the single-launch HeteroConv patient-side kernels (csrc/layer_tc.cu) through the C ABI.

The kernels feed fp32 words to tcgen05.mma.kind::tf32, which ignores the low 13 mantissa bits of every operand.  Two
references are used: (1) a float64 product of the operands truncated to TF32 the same way -- what the tensor core computes
up to fp32 accumulation order, held to 2e-5 of max|ref| (catches any layout / swizzle / barrier bug); (2) the plain fp32
product of the untruncated operands (the reference's arithmetic, model.py:125-131,256 via PyG SAGEConv mean), held to the
north_star tolerance for tensor-core outputs (1e-2) and in practice ~1e-3."""
import importlib

import pytest
import torch

PKG = "multi-modal-gnn_b200"


def _mods():
    return (importlib.import_module(PKG + ".graph"), importlib.import_module(PKG + ".ops"), importlib.import_module(PKG + ".model"),
            importlib.import_module(PKG + ".synth"), importlib.import_module(PKG + "._lib"))


def tf32_trunc(t):
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def relmax(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- host logic (no GPU) ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes", [[50, 200, 100], [20, 30, 25], [160, 200, 100], [5, 5, 5, 40], [1], [33, 31], [256, 256, 250]])
def test_bit_layout_one_boundary_per_word(sizes):
    G = importlib.import_module(PKG + ".graph")
    offs, nw, rel_a, rel_b, split = G.bit_layout(sizes)
    assert len(offs) == len(sizes) and nw == len(rel_a) == len(rel_b) == len(split)
    owner = {}
    for i, (o, n) in enumerate(zip(offs, sizes)):
        for c in range(o, o + n):
            assert c not in owner, "relations overlap on the bit axis"
            owner[c] = i
    assert max(owner) < 32 * nw
    for c, i in owner.items():          # every column resolves to its relation through (rel_a, rel_b, split)
        w, b = divmod(c, 32)
        assert (rel_a[w] if b < split[w] else rel_b[w]) == i
    assert all(1 <= s <= 32 for s in split)


# ---- kernels ---------------------------------------------------------------------------------------------------------------
def _random_hub(G, m, sizes, density, dev, seed=0):
    """random bipartite edges hub x type_i, as (dense 0/1 matrices, PatientBits-like object built through the library)"""
    gen = torch.Generator().manual_seed(seed)
    dense = [(torch.rand(m, n, generator=gen) < p).float() for n, p in zip(sizes, density)]
    dense[0][min(3, m - 1)] = 0           # an isolated patient (mean of nothing = 0, PyG clamp(count, 1))
    data = importlib.import_module(PKG + ".heterodata").HeteroGraph()
    data["patient"].num_nodes = m
    names = [f"t{i}" for i in range(len(sizes))]
    for nm, n, a in zip(names, sizes, dense):
        data[nm].num_nodes = n
        idx = a.nonzero().t().contiguous()                         # [2, E]: (patient, type)
        data["patient", "has_" + nm, nm].edge_index = idx.to(dev)
        data[nm, "has_" + nm + "_rev", "patient"].edge_index = idx.flip(0).contiguous().to(dev)
    gi = G.GraphIndex(data)
    pb = gi.hub_bits("patient")
    assert pb is not None and pb.types == names
    assert pb.bits_out is pb.bits_in, "flip(0) reverse relations must share one bit matrix"
    return dense, pb


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True], ids=["tf32adj", "f16adj"])
@pytest.mark.parametrize("m,sizes,density", [(128, [50, 200, 100], [0.2, 0.01, 0.03]), (1000, [20, 30, 25], [0.5, 0.2, 0.3]),
                                             (4099, [160, 200, 100], [0.67, 0.05, 0.28]), (70001, [50, 200, 100], [0.4, 0.01, 0.05]),
                                             (777, [5, 5, 5, 40], [0.5, 0.5, 0.5, 0.1]), (300, [300], [0.1])])
def test_layer_fwd_tc(m, sizes, density, half, pkg):
    G, ops, _, _, L = _mods()
    dev = torch.device("cuda:0")
    dense, pb = _random_hub(G, m, sizes, density, dev, seed=m)
    d = 128
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(m, d, generator=gen)
    w = torch.randn(d, d, generator=gen) / d ** 0.5
    w2 = torch.randn(d, d, generator=gen) / d ** 0.5
    b1, b2 = torch.randn(d, generator=gen), torch.randn(d, generator=gen)
    ys = [torch.randn(n, d, generator=gen) for n in sizes]
    if half:
        ys[-1] = ys[-1] * 3e-9                # rows of any magnitude (backward: gradients) must survive the fp16 B operand
        ys[0][:, 5] = 0.0
    rs = [1.0 / a.sum(1).clamp(min=1) for a in dense]
    # library result
    wcat, bias = ops.layer_cat_weights_([w.to(dev), w2.to(dev)], False, [y.to(dev) for y in ys], [None] * len(sizes), pb.offs, d,
                                        d + 32 * pb.nw, d, [b1.to(dev), b2.to(dev)])
    torch.testing.assert_close(bias.cpu(), b1 + b2)
    assert torch.equal(wcat[:, :d].cpu(), w + w2)
    for y, off, n in zip(ys, pb.offs, sizes):
        assert torch.equal(wcat[:, d + off:d + off + n].cpu(), y.t())
    rs_lib = pb.rscale_in()
    for a, b in zip(rs_lib, rs):
        torch.testing.assert_close(a.cpu(), b, rtol=0, atol=0)
    out = torch.full((m, d), float("nan"), device=dev)
    sums = torch.zeros(2 * d, dtype=torch.float64, device=dev)
    hv = None
    if half:
        # b2g_layer_cat_half: per output column j a power-of-two scale puts max_t |Y[t, j]| into [2^13, 2^14); fp16, round to nearest
        wcat0 = wcat.clone()
        hv = ops.layer_cat_half_(wcat, d)
        whalf, unscale = hv
        ycat = torch.cat(ys, 0)                                     # [sum n, d]
        mant, expo = torch.frexp(ycat.abs().max(0).values)
        e_j = torch.where(ycat.abs().max(0).values > 0, 14 - expo, torch.zeros_like(expo)).double()
        sc_j = torch.pow(torch.tensor(2.0, dtype=torch.float64), e_j)
        torch.testing.assert_close(unscale.cpu().double(), 1.0 / sc_j, rtol=0, atol=0)
        assert torch.equal(wcat[:, :d].cpu().double(), wcat0[:, :d].cpu().double() * sc_j[:, None])
        for y, off, n in zip(ys, pb.offs, sizes):
            assert torch.equal(whalf[:, off:off + n].cpu().double(), (y.double() * sc_j[None, :]).half().double().t())
        assert whalf.shape[1] == 64 * ((pb.nw + 1) // 2)
    ops.layer_fwd_tc_(x.to(dev), wcat, bias, pb.bits_in, pb, rs_lib, out, sums, hv)
    torch.cuda.synchronize()
    # references
    ref = x.double() @ (w + w2).double().t() + (b1 + b2).double()
    ref_t = tf32_trunc(x).double() @ tf32_trunc(w + w2).double().t() + (b1 + b2).double()
    for a, r, y in zip(dense, rs, ys):
        ref += (a.double() * r.double()[:, None]) @ y.double()
        if half:
            ref_t += (a.double() * r.half().double()[:, None]) @ ((y.double() * sc_j[None, :]).half().double() / sc_j[None, :])
        else:
            ref_t += (a.double() * tf32_trunc(r).double()[:, None]) @ tf32_trunc(y).double()
    assert torch.isfinite(out).all()
    assert relmax(out, ref_t) < 2e-5, "tcgen05 result differs from the float64 product of the operands as the tensor core reads them"
    assert relmax(out, ref) < 3e-3
    # BatchNorm column statistics from the epilogue
    o64 = out.double().cpu()
    torch.testing.assert_close(sums[:d].cpu(), o64.sum(0), rtol=1e-6, atol=1e-6 * m)
    torch.testing.assert_close(sums[d:].cpu(), (o64 * o64).sum(0), rtol=1e-6, atol=1e-6 * m)
    # deterministic
    out2 = torch.empty_like(out)
    ops.layer_fwd_tc_(x.to(dev), wcat, bias, pb.bits_in, pb, rs_lib, out2, None, hv)
    assert torch.equal(out, out2)
    if half:
        # the contribution of the tiny-magnitude relation alone (x part and the other relations zeroed): relative precision kept
        y_only = [torch.zeros_like(y) for y in ys[:-1]] + [ys[-1]]
        wz, _ = ops.layer_cat_weights_([torch.zeros(d, d, device=dev)], False, [y.to(dev) for y in y_only], [None] * len(sizes), pb.offs, d,
                                       d + 32 * pb.nw, d, [])
        hz = ops.layer_cat_half_(wz, d)
        out3 = torch.empty_like(out)
        ops.layer_fwd_tc_(x.to(dev), wz, None, pb.bits_in, pb, rs_lib, out3, None, hz)
        ref3 = (dense[-1].double() * rs[-1].double()[:, None]) @ ys[-1].double()
        assert relmax(out3, ref3) < 2e-3, "small-magnitude rows lost in the fp16 operand"


@pytest.mark.gpu
@pytest.mark.parametrize("m,sizes,density", [(32, [50, 200, 100], [0.2, 0.01, 0.03]), (1000, [20, 30, 25], [0.5, 0.2, 0.3]),
                                             (4099, [160, 200, 100], [0.67, 0.05, 0.28]), (70001, [50, 200, 100], [0.4, 0.01, 0.05]),
                                             (777, [5, 5, 5, 40], [0.5, 0.5, 0.5, 0.1]), (5000, [300], [0.1])])
def test_layer_adjT_tc(m, sizes, density, pkg):
    G, ops, _, _, L = _mods()
    dev = torch.device("cuda:0")
    dense, pb = _random_hub(G, m, sizes, density, dev, seed=m + 1)
    d = 128
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(m, d, generator=gen)
    xt = tf32_trunc(x).double()
    # forward use: column scale = 1/deg of the type node, no row scale
    out = ops.layer_adjT_tc_(x.to(dev), pb.bits_out, pb, [None] * len(sizes), pb.col_scale_out())
    torch.cuda.synchronize()
    assert out.shape == (32 * pb.nw, d)
    covered = torch.zeros(32 * pb.nw, dtype=torch.bool)
    for a, off, n in zip(dense, pb.offs, sizes):
        cs = 1.0 / a.sum(0).clamp(min=1)
        ref_t = (a.double().t() @ xt) * cs.double()[:, None]
        ref = (a.double().t() @ x.double()) * cs.double()[:, None]
        got = out[off:off + n]
        assert relmax(got, ref_t) < 2e-5
        assert relmax(got, ref) < 3e-3
        covered[off:off + n] = True
    assert float(out.cpu()[~covered].abs().max() if (~covered).any() else 0.0) == 0.0, "padding columns must stay zero"
    # backward use: row scale = 1/deg of the patient per relation, no column scale
    rs = pb.rscale_in()
    out = ops.layer_adjT_tc_(x.to(dev), pb.bits_in, pb, rs, None)
    for a, r, off, n in zip(dense, rs, pb.offs, sizes):
        ref_t = (a.double() * tf32_trunc(r.cpu()).double()[:, None]).t() @ xt
        assert relmax(out[off:off + n], ref_t) < 2e-5
    out2 = ops.layer_adjT_tc_(x.to(dev), pb.bits_in, pb, rs, None)
    assert torch.equal(out, out2)
    # column sums of x ride along as one more column of ones (the bias gradient when x = dout)
    out3 = ops.layer_adjT_tc_(x.to(dev), pb.bits_in, pb, rs, None, with_colsum=True)
    assert out3.shape == (32 * (pb.nw + 1), d)
    assert relmax(out3[:32 * pb.nw], out) < 1e-6
    assert relmax(out3[32 * pb.nw], xt.sum(0)) < 2e-5
    assert float(out3[32 * pb.nw + 1:].abs().max()) == 0.0
    # ... and x^T B for a dense B [m, 128] as 128 more columns (the weight gradient dout^T x_patient), when TMEM holds it all
    if ops.adjT_columns_fit(pb, True, True):
        bmat = torch.randn(m, d, generator=gen)
        out4, dw = ops.layer_adjT_tc_(x.to(dev), pb.bits_in, pb, rs, None, with_colsum=True, dense_b=bmat.to(dev))
        assert torch.equal(out4, out3), "the adjacency columns must not depend on the dense block"
        assert relmax(dw, xt.t() @ tf32_trunc(bmat).double()) < 2e-5
        assert relmax(dw, x.double().t() @ bmat.double()) < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("spec_name,n_p", [("tiny", 2000), ("C1", 1834), ("C2", 6000)])
def test_fused_layer_equals_per_relation_path(spec_name, n_p, pkg):
    """model._layer through PatientSideFn (bit adjacency, 2 + 3 launches) vs the per-relation path of round 1, forward and
    backward, and both against the exact-fp32 kernels (which are pinned to the reference's golden vectors)."""
    G, ops, M, S, L = _mods()
    dev = torch.device("cuda:0")
    base = S.SPECS[spec_name]
    scale = n_p / base.n_patient
    spec = S.GraphSpec("t", n_p, base.n_lab, base.n_dx, base.n_med, int(base.e_lab * scale), int(base.e_dx * scale), int(base.e_med * scale),
                       base.low_degree_frac)
    g = S.make_graph(spec, seed=5).to(dev)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True, "activation": "relu"}}
    torch.manual_seed(0)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
    model._init_embeddings(g)
    model.train()
    gi = model._graph_index(g)
    gen = torch.Generator(device=dev).manual_seed(3)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    x0 = {nt: torch.randn(n, 128, device=dev, generator=gen) for nt, n in counts.items()}
    gout = {nt: torch.randn(n, 128, device=dev, generator=gen) for nt, n in counts.items()}
    params = list(model.convs[0].parameters())

    def run(mode, fused):
        ops.set_precision(mode)
        saved = M.HeteroRGCN._layer_fused
        if not fused:
            M.HeteroRGCN._layer_fused = lambda self, *a, **k: None
        try:
            x = {nt: v.clone().requires_grad_(True) for nt, v in x0.items()}
            for p in params:
                p.grad = None
            out = model._layer(0, x, gi)
            torch.autograd.backward([out[nt] for nt in out], [gout[nt] for nt in out])
            res = {"out." + nt: out[nt].detach().clone() for nt in out}
            res.update({"dx." + nt: x[nt].grad.clone() for nt in x})
            res.update({"dp.%d" % i: p.grad.clone() for i, p in enumerate(params)})
            return res
        finally:
            M.HeteroRGCN._layer_fused = saved

    old = ops.PRECISION
    try:
        exact = run("fp32", False)
        launches0 = L.load().b2g_launch_count()
        fused = run("tf32", True)
        n_fused = L.load().b2g_launch_count() - launches0
        launches0 = L.load().b2g_launch_count()
        unfused = run("tf32", False)
        n_unfused = L.load().b2g_launch_count() - launches0
    finally:
        ops.set_precision(old)
    assert set(exact) == set(fused) == set(unfused)
    assert n_fused < n_unfused, f"the fused layer should need fewer launches ({n_fused} vs {n_unfused})"
    for k in exact:
        e_f, e_u = relmax(fused[k], exact[k]), relmax(unfused[k], exact[k])
        assert e_f < 5e-3, f"{k}: fused tf32 vs exact fp32 {e_f:.2e}"
        assert e_u < 5e-3, f"{k}: per-relation tf32 vs exact fp32 {e_u:.2e}"


@pytest.mark.gpu
def test_batchnorm_statistics_come_from_the_layer_epilogue(pkg):
    """The patient BatchNorm after a fused layer (model.py:259-261) takes {sum x, sum x^2} from k_layer_tf32's epilogue
    (b2g_bn_finalize_sums) instead of a pass of its own; outputs and running buffers equal the separate-pass result."""
    G, ops, M, S, L = _mods()
    dev = torch.device("cuda:0")
    g = S.make_graph(S.GraphSpec("t", 3000, 50, 114, 100, 90000, 8000, 24000, 0.05), seed=2).to(dev)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True, "activation": "relu"}}
    old = ops.PRECISION
    ops.set_precision("tf32")
    try:
        res = {}
        for fused in (True, False):
            torch.manual_seed(0)
            model = M.build_model(cfg, (g.node_types, g.edge_types), None).to(dev)
            model._init_embeddings(g)
            model.train()
            saved = M.HeteroRGCN._layer_fused
            if not fused:
                M.HeteroRGCN._layer_fused = lambda self, *a, **k: None
            try:
                ops.PROFILE = []
                out = model(g)
                names = [p[0] for p in ops.PROFILE]
            finally:
                ops.PROFILE = None
                M.HeteroRGCN._layer_fused = saved
            res[fused] = (out, {k: v.clone() for k, v in model.state_dict().items() if "running" in k}, names)
        # 2 patient-MLP BatchNorms (statistics from k_linear_tf32's epilogue, both paths) + 2 GNN-layer patient BatchNorms (fused only)
        assert res[True][2].count("b2g_bn_finalize_sums") == 4 and res[False][2].count("b2g_bn_finalize_sums") == 2
        assert res[True][2].count("b2g_bn_stats") == res[False][2].count("b2g_bn_stats") - 2
        for nt in res[True][0]:
            assert relmax(res[True][0][nt], res[False][0][nt]) < 5e-3, nt
        for k in res[True][1]:
            assert relmax(res[True][1][k], res[False][1][k]) < 2e-3, k
    finally:
        ops.set_precision(old)


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k", [(512, 128, 128), (4099, 128, 128), (70001, 128, 128), (5000, 64, 128), (3000, 128, 64)])
def test_linear_epilogue_statistics_and_l2norm(m, n, k, pkg):
    """k_linear_tf32's extended epilogue (b2g_linear_fwd_tc_ex): BatchNorm column statistics of y = x W^T + b, and
    F.normalize(y) with the reciprocal norms (model.py:93-105,232), against torch in float64 on TF32-truncated operands."""
    G, ops, _, _, L = _mods()
    lib = L.load()
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(m + n)
    x = torch.randn(m, k, generator=gen)
    w = torch.randn(n, k, generator=gen) / k ** 0.5
    b = torch.randn(n, generator=gen)
    ref = tf32_trunc(x).double() @ tf32_trunc(w).double().t() + b.double()
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    y = torch.empty(m, n, device=dev)
    sums = torch.zeros(2 * n, dtype=torch.float64, device=dev)
    ws = torch.empty(lib.b2g_linear_stats_ws_bytes(n), dtype=torch.uint8, device=dev)
    L.check(lib.b2g_linear_fwd_tc_ex(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), m, n, k, y.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                     None, 0.0, None))
    torch.cuda.synchronize()
    assert relmax(y, ref) < 2e-5
    y64 = y.double().cpu()
    torch.testing.assert_close(sums[:n].cpu(), y64.sum(0), rtol=2e-6, atol=2e-6 * m)
    torch.testing.assert_close(sums[n:].cpu(), (y64 * y64).sum(0), rtol=2e-6, atol=2e-6 * m)
    # fused row normalisation
    y2 = torch.empty(m, n, device=dev)
    inv = torch.empty(m, device=dev)
    L.check(lib.b2g_linear_fwd_tc_ex(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), m, n, k, y2.data_ptr(), None, None, 0, inv.data_ptr(), 1e-12, None))
    torch.cuda.synchronize()
    nrm = ref.norm(dim=1).clamp_min(1e-12)
    assert relmax(y2, ref / nrm[:, None]) < 2e-5
    assert relmax(inv, 1.0 / nrm) < 2e-5
    # autograd wrapper == separate linear + normalize
    ops.set_precision("tf32")
    xa, wa, ba = xd.clone().requires_grad_(True), wd.clone().requires_grad_(True), bd.clone().requires_grad_(True)
    xb, wb, bb = xd.clone().requires_grad_(True), wd.clone().requires_grad_(True), bd.clone().requires_grad_(True)
    go = torch.randn(m, n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    if ops.LinearL2NormFn.supported(xa, wa):
        ops.LinearL2NormFn.apply(xa, wa, ba, 1e-12).backward(go)
        ops.L2NormFn.apply(ops.linear(xb, wb, bb), 1e-12).backward(go)
        for a_, b_ in ((xa, xb), (wa, wb), (ba, bb)):
            assert relmax(a_.grad, b_.grad) < 5e-4          # (the two forward results differ in the last bits: fused vs separate normalisation)


@pytest.mark.gpu
@pytest.mark.parametrize("m,sizes,density,d", [(5000, [50, 200, 100], [0.2, 0.01, 0.03], 128), (70001, [160, 200, 100], [0.3, 0.02, 0.1], 128),
                                               (9000, [20, 30, 25], [0.5, 0.2, 0.3], 64), (6000, [300], [0.1], 256)])
def test_gather_reduce_staged_equals_unstaged(m, sizes, density, d, pkg):
    """b2g_gather_reduce_staged (source tables staged in shared memory by bulk copies, north_star (b)) computes the mean
    aggregation onto the patient rows (PyG SAGEConv(aggr='mean').propagate, model.py:256) with the same arithmetic and edge
    order as b2g_gather_reduce: bit-identical, and within 1e-5 of the float64 product."""
    G, ops, _, _, L = _mods()
    dev = torch.device("cuda:0")
    dense, pb = _random_hub(G, m, sizes, density, dev, seed=m + 7)
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(n, d, generator=gen) for n in sizes]
    gi_rel = [rel for rel in pb.in_rel]
    assert all(r is not None for r in gi_rel)
    csrs = [r.by_dst for r in gi_rel]
    rsc = [r.by_dst.inv_deg for r in gi_rel]
    xd = [x.to(dev) for x in xs]
    outs = {}
    old = ops.GATHER_STAGED
    try:
        # the warp-per-row kernel (k_gather_reduce) is the reference for both streaming forms (global rows / staged tables)
        ops.GATHER_STAGED = False
        saved_stream = (ops.GATHER_STREAM, ops.STREAM_MIN_ROWS, ops.STREAM_MAX_AVG_DEG)
        ops.GATHER_STREAM = False
        plain = torch.full((m, d), float("nan"), device=dev)
        ops.gather_reduce_(csrs, xd, rsc, [None] * len(sizes), plain, False)
        ops.GATHER_STREAM, ops.STREAM_MIN_ROWS, ops.STREAM_MAX_AVG_DEG = True, 1, 10 ** 9       # force the streaming form below
        for staged in (False, True):
            ops.GATHER_STAGED = staged
            ops.PROFILE = []
            out = torch.full((m, d), float("nan"), device=dev)
            ops.gather_reduce_(csrs, xd, rsc, [None] * len(sizes), out, False)
            names = [p[0] for p in ops.PROFILE]
            fits = sum(sizes) * d * 4 <= 200 * 1024          # [160, 200, 100] x 512 B does not: the call falls back to the L1 / L2 kernel
            assert ("b2g_gather_reduce_staged" in names) == (staged and fits) and ("b2g_gather_reduce" in names) == (not (staged and fits))
            out2 = out.clone()
            ops.gather_reduce_(csrs, xd, rsc, [None] * len(sizes), out2, True)       # accumulate
            outs[staged] = (out, out2)
    finally:
        ops.GATHER_STAGED = old
        ops.PROFILE = None
        ops.GATHER_STREAM, ops.STREAM_MIN_ROWS, ops.STREAM_MAX_AVG_DEG = saved_stream
    assert torch.equal(outs[True][0], outs[False][0]) and torch.equal(outs[True][1], outs[False][1])
    assert torch.equal(outs[False][0], plain), "k_gather_reduce_stream differs from k_gather_reduce"
    ref = sum((a.double() / a.sum(1).clamp(min=1).double()[:, None]) @ x.double() for a, x in zip(dense, xs))
    assert relmax(outs[True][0], ref) < 1e-5
    assert relmax(outs[True][1], 2 * ref) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True], ids=["tf32adj", "f16adj"])
@pytest.mark.parametrize("m,n,kx,sizes", [(1000, 64, 64, [50, 200, 100]), (3001, 256, 256, [20, 30, 25]), (20000, 128, 0, [160, 200, 100]),
                                          (129, 32, 32, [33, 31]), (128 * 148 * 2 + 5, 128, 128, [40]), (777, 128, 128, [256, 256, 250])])
def test_layer_fwd_tc_general_shapes(m, n, kx, sizes, half, pkg):
    """b2g_layer_fwd_tc for every shape b2g_layer_fwd_tc_supported admits (the model only uses n = kx = 128): output widths
    32..256, no dense part (kx = 0), one and 24 bit words per row, several tiles per CTA with a partial last tile."""
    G, ops, _, _, L = _mods()
    lib = L.load()
    dev = torch.device("cuda:0")
    dense, pb = _random_hub(G, m, sizes, [0.1] * len(sizes), dev, seed=m + n)
    assert lib.b2g_layer_fwd_tc_supported(m, n, kx, pb.nw)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(m, max(kx, 32), generator=gen)[:, :kx].contiguous() if kx else None
    w = torch.randn(n, max(kx, 1), generator=gen)[:, :kx].contiguous() / max(kx, 1) ** 0.5
    b = torch.randn(n, generator=gen)
    ys = [torch.randn(s, n, generator=gen) for s in sizes]
    rs = [1.0 / a.sum(1).clamp(min=1) for a in dense]
    wcat, bias = ops.layer_cat_weights_([w.to(dev)] if kx else [], False, [y.to(dev) for y in ys], [None] * len(sizes), pb.offs, kx, kx + 32 * pb.nw, n,
                                        [b.to(dev)])
    hv = ops.layer_cat_half_(wcat, kx) if half else None
    out = torch.full((m, n), float("nan"), device=dev)
    xd = x.to(dev) if kx else torch.empty((m, 0), device=dev)
    ops.layer_fwd_tc_(xd, wcat, bias, pb.bits_in, pb, pb.rscale_in(), out, None, hv)
    torch.cuda.synchronize()
    ref = b.double().expand(m, n).clone()
    if kx:
        ref = ref + x.double() @ w.double().t()
    for a, r, y in zip(dense, rs, ys):
        ref += (a.double() * r.double()[:, None]) @ y.double()
    assert torch.isfinite(out).all()
    assert relmax(out, ref) < 3e-3

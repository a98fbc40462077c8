"""CPU, world_size 2 over gloo: the host-side logic of the patient-partitioned multi-GPU path (dist.py).

  * partition_graph: contiguous balanced patient ranges, relabelled local ids, order-preserving edge subsets;
  * the three autograd building blocks (ReplicatedToLocal / PartialToReplicated / ScaleGrad) and the gradient bucket;
  * the ALGORITHM: a partitioned training step assembled from the oracle's CPU math + exactly the sync points the CUDA
    model uses (partial type sums, patient sync-BatchNorm, replicated->local gradients, 1/world on replicated
    parameters, loss re-weighting, one gradient all-reduce) reproduces the single-process oracle step.
The CUDA kernels themselves cannot run here; the same step on real GPUs is checked by tools/dist_check.py under torchrun.
"""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import hetero_rgcn_ref as R

PKG = "multi-modal-gnn_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, world=2, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q, args)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    results = []
    while not q.empty():
        results.append(q.get())
    errs = [r for r in results if r[1] is not None]
    assert not errs, errs[0][1]
    assert len(results) == world, "a rank died without reporting"
    return dict((r[0], r[2]) for r in results)


def _entry(fn, rank, world, port, q, args):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None, out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))


# ----------------------------------------------------------------------------------------------------------------------
def test_partition_graph_is_a_partition(pkg):
    D = importlib.import_module(PKG + ".dist")
    g = pkg.synth.make_graph("C1")
    world = 3
    bounds = D.partition_bounds(g, world)
    assert bounds[0] == 0 and bounds[-1] == 1834 and bool((bounds[1:] > bounds[:-1]).all())
    seen = {tuple(et): torch.zeros(g[et].edge_index.shape[1], dtype=torch.int32) for et in g.edge_types}
    loads = []
    for r in range(world):
        loc, info = D.partition_graph(g, world, r)
        p0, p1 = info["range"]
        assert loc["patient"].num_nodes == p1 - p0 and loc["lab"].num_nodes == 50
        load = 0
        for et in g.edge_types:
            ids = info["edge_ids"][tuple(et)]
            assert bool((ids[1:] > ids[:-1]).all()), "edge order must be preserved"
            seen[tuple(et)][ids] += 1
            ge, le = g[et].edge_index[:, ids], loc[et].edge_index
            prow = 0 if et[0] == "patient" else 1
            assert torch.equal(ge[prow] - p0, le[prow]) and torch.equal(ge[1 - prow], le[1 - prow])
            assert int(le[prow].min()) >= 0 and int(le[prow].max()) < p1 - p0
            load += ids.numel()
        assert torch.equal(loc["patient", "has_lab", "lab"].edge_attr, g["patient", "has_lab", "lab"].edge_attr[info["edge_ids"][("patient", "has_lab", "lab")]])
        loads.append(load)
    for et, cnt in seen.items():
        assert bool((cnt == 1).all()), et
    assert max(loads) <= 1.1 * min(loads), loads


def _blocks(rank, world):
    D = importlib.import_module(PKG + ".dist")
    dctx = D.DistContext()
    # ReplicatedToLocal: forward identity, backward sums the per-rank gradients
    x = torch.arange(4.0, requires_grad=True)
    y = D.replicated_to_local(x, dctx)
    (y * (rank + 1)).sum().backward()
    assert torch.equal(x.grad, torch.full((4,), 3.0))
    # PartialToReplicated: forward sums, backward identity
    a = torch.full((3,), float(rank + 1), requires_grad=True)
    b = D.partial_to_replicated(a, dctx)
    assert torch.equal(b.detach(), torch.full((3,), 3.0))
    (b * 2).sum().backward()
    assert torch.equal(a.grad, torch.full((3,), 2.0))
    # rep_param: gradient divided by world
    w = torch.ones(2, requires_grad=True)
    (D.rep_param(w, dctx) * 4).sum().backward()
    assert torch.equal(w.grad, torch.full((2,), 2.0))
    # gradient bucket: SUM, and None stays None when no rank has a gradient (N8)
    p1, p2, p3 = (torch.nn.Parameter(torch.zeros(3)) for _ in range(3))
    p1.grad = torch.full((3,), float(rank + 1))
    if rank == 0:
        p2.grad = torch.ones(3)
    D.allreduce_gradients([p1, p2, p3], dctx)
    assert torch.equal(p1.grad, torch.full((3,), 3.0)) and torch.equal(p2.grad, torch.ones(3)) and p3.grad is None
    assert dctx.global_row_count(10 + rank, torch.device("cpu")) == 21
    return True


def test_dist_building_blocks_gloo():
    _spawn(_blocks, 2)


# ----------------------------------------------------------------------------------------------------------------------
def _sync_bn(sd, prefix, x, dctx, m_total):
    """differentiable sync BatchNorm (training mode) on CPU tensors: statistics over the rows of all ranks"""
    import torch.distributed.nn.functional as dfn
    s0 = dfn.all_reduce(x.sum(0).double(), group=dctx.group)
    s1 = dfn.all_reduce((x.double() ** 2).sum(0), group=dctx.group)
    mean = s0 / m_total
    var = s1 / m_total - mean ** 2
    xh = (x - mean.float()) / torch.sqrt(var.float() + R.BN_EPS)
    return xh * sd[prefix + ".weight"] + sd[prefix + ".bias"]


def _partitioned_step(rank, world, spec_seed):
    pkg = importlib.import_module(PKG)
    D = importlib.import_module(PKG + ".dist")
    dctx = D.DistContext()
    g = pkg.synth.make_graph("tiny", seed=spec_seed)
    counts = {nt: int(g[nt].num_nodes) for nt in g.node_types}
    ets = [tuple(e) for e in g.edge_types]
    sd = R.init_state(counts, ets, seed=2)
    ei_g = g["patient", "has_lab", "lab"].edge_index
    attr_g = g["patient", "has_lab", "lab"].edge_attr.squeeze(-1)
    m_train = R.split_masks(ei_g.shape[1])[0]
    sup_g = torch.zeros(ei_g.shape[1], dtype=torch.bool)
    sup_g[m_train.nonzero().squeeze(1)] = R.supervision_mask(int(m_train.sum()), 0.2, 99)
    w = R.lab_weights(ei_g[1][m_train], attr_g[m_train], counts["lab"])

    # ---- single-process truth (every rank computes it; cheap at this size)
    sup_train = sup_g[m_train]
    loss_ref, pred_ref, grads_ref = R.train_step_grads({k: v.clone() for k, v in sd.items()}, counts, ets, g.edge_index_dict,
                                                       ei_g[0][m_train], ei_g[1][m_train], attr_g[m_train], sup_train, w, "mse", 0.0)

    # ---- partitioned step
    loc, info = D.partition_graph(g, world, rank)
    p0, p1 = info["range"]
    n_p_total = counts["patient"]
    ids = info["edge_ids"][("patient", "has_lab", "lab")]
    tr_l, sup_l = m_train[ids], sup_g[ids]
    ei_l = loc["patient", "has_lab", "lab"].edge_index
    pi, li, tgt, sup = ei_l[0][tr_l], ei_l[1][tr_l], attr_g[ids][tr_l], sup_l[tr_l]
    keys = R.trainable_keys(sd, include_tables=True)
    leaf = {k: sd[k].clone().requires_grad_(True) for k in keys}
    P = lambda k: leaf[k]                       # noqa: E731  local use of a parameter
    RP = lambda k: D.rep_param(leaf[k], dctx)    # noqa: E731  use inside replicated work
    # global degrees of the type nodes (one integer all-reduce per relation)
    inv_deg_t = {}
    for et in ets:
        if et[2] != "patient":
            d_ = torch.bincount(loc[et].edge_index[1], minlength=counts[et[2]])
            dist.all_reduce(d_)
            inv_deg_t[et] = 1.0 / d_.clamp(min=1).float()

    def encode():
        x = {nt: (P(f"embeddings.{nt}.weight")[p0:p1] if nt == "patient" else RP(f"embeddings.{nt}.weight")) for nt in counts}
        h = F.linear(x["patient"], P("patient_transform.0.weight"), P("patient_transform.0.bias"))
        h = F.relu(_sync_bn(leaf, "patient_transform.1", h, dctx, n_p_total))
        h = F.linear(h, P("patient_transform.4.weight"), P("patient_transform.4.bias"))
        h = F.relu(_sync_bn(leaf, "patient_transform.5", h, dctx, n_p_total))
        h = F.linear(h, P("patient_transform.8.weight"), P("patient_transform.8.bias"))
        x["patient"] = F.normalize(h, p=2.0, dim=1, eps=R.L2_EPS)
        return x

    def layer(l, x):
        out = {}
        for dst in counts:
            terms = []
            for et in ets:
                if et[2] != dst:
                    continue
                k = R.conv_key(l, et)
                e_l = loc[et].edge_index
                if dst == "patient":                                    # rank-local rows, replicated sources
                    y = F.linear(x[et[0]], RP(k + ".lin_l.weight"))     # replicated product on the type rows ...
                    y = D.replicated_to_local(y, dctx)                  # ... consumed by this rank's patients only
                    agg = R.mean_aggregate(y, e_l, p1 - p0)
                    terms.append(agg + P(k + ".lin_l.bias") + F.linear(x["patient"], P(k + ".lin_r.weight")))
                else:                                                   # replicated rows, sources spread over ranks
                    msg = x["patient"].index_select(0, e_l[0])
                    part = torch.zeros(counts[dst], msg.shape[1]).index_add_(0, e_l[1], msg) * inv_deg_t[et].unsqueeze(1)
                    agg = D.partial_to_replicated(part, dctx)
                    terms.append(F.linear(agg, RP(k + ".lin_l.weight"), RP(k + ".lin_l.bias")) + F.linear(x[dst], RP(k + ".lin_r.weight")))
            o = terms[0] if len(terms) == 1 else torch.stack(terms, 0).sum(0)
            pre = f"batch_norms.{l}.{dst}"
            if dst == "patient":
                o = _sync_bn(leaf, pre, o, dctx, n_p_total)
            else:
                o = F.batch_norm(o, None, None, RP(pre + ".weight"), RP(pre + ".bias"), True, 0.1, R.BN_EPS)
            out[dst] = F.relu(o)
        return out

    init = encode()
    x = init
    for l in range(2):
        x = layer(l, x)
    deg = torch.bincount(ei_l[0], minlength=p1 - p0)
    low = deg[pi] < 6

    def head(name, hp, hl, sel):
        w0 = leaf[f"{name}.mlp.0.weight"]
        u = F.linear(hp, w0[:, :128])
        v = D.replicated_to_local(F.linear(hl, D.rep_param(w0[:, 128:], dctx), RP(f"{name}.mlp.0.bias")), dctx)
        z = F.relu(u[pi[sel]] + v[li[sel]])
        z = F.relu(F.linear(z, P(f"{name}.mlp.3.weight"), P(f"{name}.mlp.3.bias")))
        return F.linear(z, P(f"{name}.mlp.6.weight"), P(f"{name}.mlp.6.bias")).squeeze(-1)

    pred = torch.zeros(pi.numel())
    if bool(low.any()):
        pred = pred.index_put((low.nonzero().squeeze(1),), head("tabular_mlp", init["patient"], init["lab"], low))
    if bool((~low).any()):
        pred = pred.index_put(((~low).nonzero().squeeze(1),), head("edge_predictor", x["patient"], x["lab"], ~low))
    n_loc = sup.sum().float()
    n_tot = n_loc.clone()
    dist.all_reduce(n_tot)
    loss_local = R.weighted_loss(pred, tgt, li, w, sup, "mse") if int(n_loc) > 0 else pred.sum() * 0
    loss = loss_local * (n_loc / n_tot)
    loss.backward()
    params = [leaf[k] for k in keys]
    D.allreduce_gradients(params, dctx)
    total = loss.detach().clone()
    dist.all_reduce(total)

    assert abs(float(total) - float(loss_ref)) <= 2e-5 * abs(float(loss_ref)), (float(total), float(loss_ref))
    tr_ids_local = ids[tr_l]                                     # positions of my train pairs in the global pair list
    pos_in_train = torch.cumsum(m_train.long(), 0)[tr_ids_local] - 1
    torch.testing.assert_close(pred.detach(), pred_ref[pos_in_train], rtol=1e-4, atol=1e-5)
    worst = 0.0
    for k in keys:
        gr = grads_ref[k]
        if gr is None:
            assert leaf[k].grad is None, k
            continue
        if k in ("patient_transform.0.bias", "patient_transform.4.bias") or k.endswith("lin_l.bias"):
            continue
        got = leaf[k].grad
        if k == "embeddings.patient.weight":                       # only my rows carry gradient before the bucket sum
            pass
        err = float((got - gr).abs().max() / gr.abs().max().clamp_min(1e-30))
        worst = max(worst, err)
        assert err <= 2e-3, (k, err)
    return worst


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_step_equals_single_process_gloo(world):
    """world 3: an odd split, so the partitions are of unequal size and the middle rank has two neighbours."""
    out = _spawn(_partitioned_step, world, 4)
    assert len(out) == world and max(out.values()) <= 2e-3


def _route(rank, world):
    D = importlib.import_module(PKG + ".dist")
    dctx = D.DistContext()
    bounds = torch.tensor([0, 10, 25, 40][:world + 1] if world == 3 else [0, 17, 40])
    gen = torch.Generator().manual_seed(100 + rank)
    n = 50 + 7 * rank
    pi = torch.randint(0, int(bounds[-1]), (n,), generator=gen)
    li = torch.randint(0, 9, (n,), generator=gen)
    route = D.PairRoute(pi, li, bounds, dctx)
    p0, p1 = int(bounds[rank]), int(bounds[rank + 1])
    assert bool(((route.patient_local >= 0) & (route.patient_local < p1 - p0)).all()), "every routed pair belongs to this rank's range"
    # the owner "predicts" f(global patient, lab); the asking rank gets the values back in ITS order
    pred_local = ((route.patient_local + p0) * 1000 + route.lab_local).double()
    back = route.gather_back(pred_local)
    assert torch.equal(back, (pi * 1000 + li).double())
    # totals are conserved
    tot = torch.tensor([int(route.patient_local.numel()), n])
    dist.all_reduce(tot)
    assert int(tot[0]) == int(tot[1])
    return True


@pytest.mark.parametrize("world", [2, 3])
def test_pair_routing_to_owner_ranks(world):
    """Bulk imputation (BASELINE config 5): arbitrary global (patient, lab) pairs reach the rank owning the patient and the
    predictions return in the caller's order (gloo all-to-all on the CPU)."""
    res = _spawn(_route, world)
    assert all(res.values())

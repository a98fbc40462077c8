"""Drop-in replacement for the reference's ``src/model.py`` call surface:

    build_model(config, metadata, patient_feature_dim) -> nn.Module
    HeteroRGCN._init_embeddings / encode_nodes / forward / predict_lab_values
    EdgeRegressionHead, compute_regression_loss

Same parameter / buffer names and shapes as the reference (108 ``state_dict`` entries for the shipped
config), so ``torch.optim.Adam``, ``torch.save`` and ``load_state_dict`` in the reference's ``train.py`` /
``evaluate.py`` / ``inference.py`` keep working unchanged; all arithmetic runs in libb2g's sm_100a kernels
(``ops.py``).  The ``nn.Linear`` / ``nn.BatchNorm1d`` / ``nn.Embedding`` objects below are parameter
containers only -- their ``forward`` is never called.

Reference behaviours kept on purpose (SURVEY.md section 3.1 notes):
  N2  tables are created lazily at the first forward, i.e. after train.py built its optimizer -> they are
      not optimised unless the caller creates them first (``_init_embeddings``);
  N3  ``predict_lab_values`` encodes twice in training mode (two dropout draws, two running-stat updates of
      the patient MLP's BatchNorms); with dropout == 0 the two passes are identical, so one pass is computed
      and the second running-stat update is applied analytically;
  N4  edge_attr is never a message weight;  N8  the last layer's diagnosis / medication outputs are computed
      (their BN running stats move) but receive no gradient.
Reference defect fixed on purpose: N1 -- lazily created tables are placed on the module's device.
"""
from __future__ import annotations

import logging
import re
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn as nn

from . import _lib, ops
from .dist import (DistContext, globalize_degrees, partial_to_replicated_many, rep_param, replicated_to_local,
                   replicated_to_local_many)
from .graph import GraphIndex, PairIndex, graph_index, _stream

EdgeType = Tuple[str, str, str]


class _DropoutStreams:
    """Per-call dropout bookkeeping: one 64-bit seed drawn from torch's CPU generator (so
    ``torch.manual_seed`` makes runs repeatable) and a running stream id, one per dropout site."""

    SEED_IS_POINTER = 1 << 63          # csrc/common.cuh: `seed` is a device pointer to the 64-bit seed

    def __init__(self, active: bool, seed_buffer: Optional[torch.Tensor] = None):
        self.next_id = 0
        self.log: List[Tuple[str, int]] = []
        self.flag = 0
        if active and seed_buffer is not None:
            # CUDA-graph mode: kernels read the seed from device memory; the owner of the graph rewrites that
            # buffer before every replay (trainer.py), so nothing here may touch it.
            self.seed = seed_buffer.data_ptr()
            self.flag = self.SEED_IS_POINTER
        else:
            self.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if active else 0

    def take(self, tag: str = "") -> int:
        sid = self.next_id
        self.next_id += 1
        self.log.append((tag, sid))
        return sid | self.flag


def _bn_act_drop(x, bn: Optional[nn.BatchNorm1d], training: bool, act: int, p: float, streams: _DropoutStreams, tag: str,
                 dctx: Optional[DistContext] = None, sharded_rows: bool = False):
    """dctx given: `sharded_rows` says whether x's rows are this rank's share of a partitioned node type (statistics are
    all-reduced) or a replicated tensor (every rank repeats the computation; parameters see the full gradient)."""
    sid = streams.take(tag) if (training and p > 0) else 0
    if bn is None:
        return ops.ActDropFn.apply(x, act, p, streams.seed, sid, training)
    if training and bn.track_running_stats:
        bn.num_batches_tracked += 1
    mom = bn.momentum if bn.momentum is not None else 0.1
    if dctx is not None and sharded_rows and training:
        m_total = dctx.global_row_count(x.shape[0], x.device)
        return ops.SyncBNActDropFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, act, p, streams.seed, sid, bn.eps, mom,
                                         dctx, m_total)
    gamma, beta = bn.weight, bn.bias
    if dctx is not None and not sharded_rows:
        gamma, beta = rep_param(gamma, dctx), rep_param(beta, dctx)
    return ops.BNActDropFn.apply(x, gamma, beta, bn.running_mean, bn.running_var, training, act, p, streams.seed, sid, bn.eps, mom)


class EdgeRegressionHead(nn.Module):
    """Linear(2d,64)-ReLU-Dropout-Linear(64,32)-ReLU-Dropout-Linear(32,1)  (model.py:342-396)."""

    def __init__(self, input_dim: int, hidden_dims: list = [64, 32], output_dim: int = 1, dropout: float = 0.2):
        super().__init__()
        layers = []
        prev = input_dim
        for h in hidden_dims:
            layers += [nn.Linear(prev, h), nn.ReLU(), nn.Dropout(dropout)]
            prev = h
        layers.append(nn.Linear(prev, output_dim))
        self.mlp = nn.Sequential(*layers)
        self.dropout_p = float(dropout)
        self.fused = True     # hidden_dims == [64, 32] (the reference's hard-coded shape) -> csrc/decoder.cu

    def _linears(self):
        return [m for m in self.mlp if isinstance(m, nn.Linear)]

    def _tail(self, h, linears, streams, tag, relu_done=False):
        """everything after the first Linear: ReLU-Dropout-(Linear-ReLU-Dropout)*-Linear; ``relu_done`` says
        the first ReLU was already fused into the producer of ``h``."""
        training = self.training
        drop = training and self.dropout_p > 0
        for i, lin in enumerate(linears):
            need_relu = not (relu_done and i == 0)
            if need_relu or drop:
                sid = streams.take(f"{tag}.drop{i}") if drop else 0
                h = ops.ReluDropoutFn.apply(h, int(need_relu), self.dropout_p, streams.seed, sid, training)
            h = ops.linear(h, lin.weight, lin.bias)
        return h

    def forward(self, edge_embeds: torch.Tensor, _streams: Optional[_DropoutStreams] = None) -> torch.Tensor:
        """edge_embeds [m, input_dim] -> [m, output_dim]  (model.py:388-396)."""
        streams = _streams or _DropoutStreams(self.training and self.dropout_p > 0)
        lins = self._linears()
        h = ops.linear(edge_embeds, lins[0].weight, lins[0].bias)
        return self._tail(h, lins[1:], streams, "head")

    def forward_pairs(self, h_p: torch.Tensor, h_l: torch.Tensor, pairs: PairIndex, streams: _DropoutStreams, tag: str,
                      dctx: Optional[DistContext] = None):
        """Same function of cat([h_p[pi], h_l[li]], 1) (model.py:319-333) with the first layer factorised per
        node: U = h_p W[:, :d]^T on patients, V = h_l W[:, d:]^T + b on labs, z = relu(U[pi] + V[li])."""
        lins = self._linears()
        if len(lins) < 2:
            raise _lib.B2GError("EdgeRegressionHead needs at least one hidden layer")
        d = h_p.shape[1]
        w0 = lins[0].weight
        u = ops.linear(h_p, w0[:, :d].contiguous(), None)
        # lab rows are replicated: every rank repeats this tiny product; its result is consumed by rank-local pairs
        v = ops.linear(h_l, rep_param(w0[:, d:].contiguous(), dctx), rep_param(lins[0].bias, dctx))
        v = replicated_to_local(v, dctx)
        if len(lins) == 3 and lins[1].weight.shape == (32, 64) and lins[2].weight.shape == (1, 32) and self.fused:
            drop = self.training and self.dropout_p > 0
            sid1 = streams.take(f"{tag}.drop0") if drop else 0
            sid2 = streams.take(f"{tag}.drop1") if drop else 0
            return ops.DecoderHeadFn.apply(u, v, lins[1].weight, lins[1].bias, lins[2].weight, lins[2].bias, pairs,
                                           self.dropout_p, streams.seed, sid1, sid2, self.training)
        z = ops.PairAddReluFn.apply(u, v, pairs)
        out = self._tail(z, lins[1:], streams, tag, relu_done=True)
        return out.squeeze(-1)


class _SAGEParams(nn.Module):
    """Parameter container with PyG SAGEConv's names: lin_l (bias) on the aggregated neighbours, lin_r (no
    bias) on the destination's own features (PyG sage_conv.py; param count pinned by SURVEY.md KA-1)."""

    def __init__(self, d_in: int, d_out: int):
        super().__init__()
        self.lin_l = nn.Linear(d_in, d_out, bias=True)
        self.lin_r = nn.Linear(d_in, d_out, bias=False)


class _HeteroConvParams(nn.Module):
    def __init__(self, edge_types, d: int):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(et): _SAGEParams(d, d) for et in edge_types})


_PYG24_KEY = re.compile(r"<([^<>]+?)___([^<>]+?)___([^<>]+?)>")


class HeteroRGCN(nn.Module):
    """Relational SAGE network of the reference (model.py:33-335) on libb2g kernels."""

    def __init__(self, metadata: Tuple, hidden_dim: int = 128, num_layers: int = 2, dropout: float = 0.2,
                 patient_feature_dim: int = None, use_batch_norm: bool = True, activation: str = "relu"):
        super().__init__()
        if activation not in ("relu", "elu", "leaky_relu"):
            raise ValueError(f"Unknown activation: {activation}")
        if hidden_dim not in (32, 64, 128, 256):
            raise ValueError(f"hidden_dim must be one of 32/64/128/256 for the sm_100a kernels, got {hidden_dim}")
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.dropout = dropout
        self.use_batch_norm = use_batch_norm
        self.activation_name = activation
        self._act = ops.ACT_CODES[activation]
        node_types, edge_types = metadata
        self._node_types = list(node_types)
        self._edge_types = [tuple(et) for et in edge_types]

        self.embeddings = nn.ModuleDict()
        self.embedding_dims = {}
        self.patient_transform = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim), nn.BatchNorm1d(hidden_dim), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(hidden_dim, hidden_dim))
        self.convs = nn.ModuleList([_HeteroConvParams(self._edge_types, hidden_dim) for _ in range(num_layers)])
        self.batch_norms = (nn.ModuleList([nn.ModuleDict({nt: nn.BatchNorm1d(hidden_dim) for nt in node_types})
                                           for _ in range(num_layers)]) if use_batch_norm else None)
        self.edge_predictor = EdgeRegressionHead(2 * hidden_dim, [64, 32], 1, dropout)
        self.tabular_mlp = EdgeRegressionHead(2 * hidden_dim, [64, 32], 1, dropout)
        self.degree_threshold = 6
        self._pair_plans: "OrderedDict[int, _PairPlan]" = OrderedDict()
        self._last_streams: Optional[_DropoutStreams] = None
        self._seed_buffer: Optional[torch.Tensor] = None      # set by Trainer.enable_cuda_graph()
        self._embedding_cache_on = False                      # enable_embedding_cache(): eval-mode (init, x) reuse
        self._embedding_cache = None
        self._generation = 0
        self.dist: Optional[DistContext] = None               # set by set_distributed(): patient-partitioned multi-GPU
        self._register_load_state_dict_pre_hook(self._rename_pyg24_keys)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.bump_generation())
        logging.info(f"Initialized HeteroRGCN with hidden_dim={hidden_dim}, num_layers={num_layers}")

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _rename_pyg24_keys(state_dict, prefix, *args):
        """Accept PyG >= 2.4 conv keys ('<src___rel___dst>') as well as the 2.3 form ('src__rel__dst')."""
        for k in list(state_dict.keys()):
            nk = _PYG24_KEY.sub(lambda m: "__".join(m.groups()), k)
            if nk != k:
                state_dict[nk] = state_dict.pop(k)

    def _device(self):
        return next(self.parameters()).device

    def set_distributed(self, dctx: Optional[DistContext]):
        """Patient-partitioned execution (dist.py): `data` passed to forward / predict_lab_values is then this rank's
        partition (its patients, all type nodes, its edges); results equal the single-GPU ones."""
        self.dist = dctx

    def _init_embeddings(self, data):
        """model.py:180-204 -- idempotent; tables live on the module's device (fixes reference note N1)."""
        dev = self._device()
        for nt in data.node_types:
            if nt not in self.embeddings:
                n = int(data[nt].num_nodes)
                emb = nn.Embedding(n, self.hidden_dim)
                nn.init.xavier_uniform_(emb.weight)
                self.embeddings[nt] = emb.to(dev)
                self.embedding_dims[nt] = n
                logging.info(f"Created embedding for {nt}: {n} nodes")

    # ------------------------------------------------------------------------------------------
    def _encode(self, node_types, streams: _DropoutStreams, tag: str) -> Dict[str, torch.Tensor]:
        """model.py:206-234.  embedding(arange(N)) is the table itself; patient rows go through
        Linear-BN-ReLU-Dropout x2, Linear, row L2 normalisation."""
        training = self.training
        dctx = self.dist
        sharded = dctx.sharded_type if dctx is not None else None
        x = {nt: (self.embeddings[nt].weight if (dctx is None or nt == sharded) else rep_param(self.embeddings[nt].weight, dctx))
             for nt in node_types}
        if "patient" in x:
            pt = self.patient_transform
            shard = sharded == "patient"
            # the two BatchNorms take their statistics from the preceding linear's epilogue, the row normalisation is the last
            # linear's epilogue (tf32 mode; the exact-fp32 mode keeps the separate kernels)
            h = ops.linear(x["patient"], pt[0].weight, pt[0].bias, want_stats=training)
            h = _bn_act_drop(h, pt[1], training, 1, self.dropout, streams, tag + ".drop0", dctx, shard)
            h = ops.linear(h, pt[4].weight, pt[4].bias, want_stats=training)
            h = _bn_act_drop(h, pt[5], training, 1, self.dropout, streams, tag + ".drop1", dctx, shard)
            if ops.LinearL2NormFn.supported(h, pt[8].weight):
                x["patient"] = ops.LinearL2NormFn.apply(h, pt[8].weight, pt[8].bias, 1e-12)
            else:
                h = ops.linear(h, pt[8].weight, pt[8].bias)
                x["patient"] = ops.L2NormFn.apply(h, 1e-12)
        return x

    def _layer(self, layer: int, x: Dict[str, torch.Tensor], gi: GraphIndex) -> Dict[str, torch.Tensor]:
        """HeteroConv({et: SAGEConv(mean)}, aggr='sum') (model.py:125-131,256): per destination type one fused
        SageDstFn (see ops.py)."""
        convs = self.convs[layer].convs
        by_dst: "OrderedDict[str, list]" = OrderedDict()
        for et in self._edge_types:
            if et in gi.relations and et[0] in x and et[2] in x:
                by_dst.setdefault(et[2], []).append(et)
        fused = self._layer_fused(convs, x, gi, by_dst)
        if fused is not None:
            return fused
        dctx = self.dist
        sharded = dctx.sharded_type if dctx is not None else None
        # pass 1: per-relation products; multi-GPU tensors that need an exchange are collected so that each kind costs
        # ONE all-reduce per layer (partial type sums in forward, replicated->local gradients in backward)
        plans, partial_aggs, local_ys = {}, [], []
        pending_y = []      # (ys list, position, x_src, W_l): the few-source products of ALL destinations, issued as one group
        for dst, ets in by_dst.items():
            # multi-GPU: rows of the sharded type are rank-local work; every other destination type is replicated work
            rep = dctx if (dctx is not None and dst != sharded) else None
            w_root = b_root = None
            small_rels, ys, aggs, wls = [], [], [], []
            for et in ets:
                conv = convs["__".join(et)]
                w_r, b_l, w_l = rep_param(conv.lin_r.weight, rep), rep_param(conv.lin_l.bias, rep), conv.lin_l.weight
                w_root = w_r if w_root is None else w_root + w_r
                b_root = b_l if b_root is None else b_root + b_l
                rel = gi.relations[et]
                src_replicated = dctx is not None and et[0] != sharded
                partial_sources = rep is not None and not src_replicated    # this rank holds only some of the sources
                if rel.n_src < rel.n_dst and not partial_sources:   # few sources: transform them first, then aggregate
                    small_rels.append(rel)
                    pending_y.append((ys, len(ys), x[et[0]], rep_param(w_l, dctx if src_replicated else None)))
                    if src_replicated and rep is None:     # replicated table consumed by this rank's rows only
                        local_ys.append((ys, len(ys)))
                    ys.append(None)
                else:                              # many sources: aggregate first, then transform
                    x_src = x[et[0]]
                    if src_replicated and rep is None:
                        # a replicated table consumed by this rank's rows only: every rank holds a PARTIAL gradient of it, which
                        # must be summed over ranks before it meets the 1/world of rep_param (the few-source branch does the same
                        # through `local_ys`)
                        x_src = replicated_to_local(x_src, dctx)
                    agg = ops.MeanAggFn.apply(x_src, rel)
                    if partial_sources:            # partial mean over this rank's sources -> sum over ranks
                        partial_aggs.append((aggs, len(aggs)))
                    aggs.append(agg)
                    wls.append(rep_param(w_l, rep))
            plans[dst] = (w_root, b_root, small_rels, ys, aggs, wls)
        if pending_y:
            for (lst, i, _, _), y in zip(pending_y, ops.linear_group([p[2] for p in pending_y], [p[3] for p in pending_y])):
                lst[i] = y
        if dctx is not None:
            for (lst, i), t in zip(partial_aggs, partial_to_replicated_many([lst[i] for lst, i in partial_aggs], dctx)):
                lst[i] = t
            for (lst, i), t in zip(local_ys, replicated_to_local_many([lst[i] for lst, i in local_ys], dctx)):
                lst[i] = t
        # pass 2: one fused SageDstFn per destination type; the few-row destinations with a single aggregated relation
        # (labs, diagnoses, medications) share ONE grouped launch each way
        out = {}
        grouped = [dst for dst, (w_root, b_root, small_rels, ys, aggs, wls) in plans.items()
                   if not small_rels and len(aggs) == 1 and x[dst].shape[0] <= ops.SMALL_ROWS and b_root is not None]
        if len(grouped) >= 2:
            n = len(grouped)
            res = ops.SageTypeDstGroupFn.apply(n, *[x[d] for d in grouped], *[plans[d][0] for d in grouped], *[plans[d][1] for d in grouped],
                                               *[plans[d][4][0] for d in grouped], *[plans[d][5][0] for d in grouped])
            for d, o in zip(grouped, res):
                out[d] = o
        for dst, (w_root, b_root, small_rels, ys, aggs, wls) in plans.items():
            if dst not in out:
                out[dst] = ops.SageDstFn.apply(x[dst], w_root, b_root, tuple(small_rels), len(aggs), *ys, *aggs, *wls)
        return {dst: out[dst] for dst in plans}

    def _layer_fused(self, convs, x, gi: GraphIndex, by_dst) -> Optional[Dict[str, torch.Tensor]]:
        """The same layer with the hub (patient) side in ops.PatientSideFn: one tcgen05 launch for out_patient (self term +
        the three type -> patient means, adjacency as bits), one for the three patient -> type neighbour sums; the few-row
        products on the type nodes stay in the grouped small-GEMM launches.  tf32 mode, hub-shaped graphs, d = 128; returns
        None when not applicable (the caller then runs the per-relation path)."""
        dctx = self.dist
        hub = dctx.sharded_type if dctx is not None else "patient"
        if hub not in x or hub not in by_dst or ops.PRECISION != "tf32" or os.environ.get("B2G_FUSED_LAYER", "1") == "0":
            return None
        pb = gi.hub_bits(hub)
        if not ops.patient_side_supported(pb, x[hub].shape[0], self.hidden_dim):
            return None
        live = {et for ets in by_dst.values() for et in ets}
        if live != set(gi.relations.keys()) or any(t not in x for t in pb.types):
            return None                                   # graph relations the model does not know (or vice versa)
        if any(x[t].shape[0] > ops.SMALL_ROWS for t in pb.types):
            return None
        # few-row products first: Y_t = x_t W_l^T for every relation t -> hub (replicated work in multi-GPU mode)
        w_roots, b_roots, y_x, y_w = [], [], [], []
        for t, rel in zip(pb.types, pb.in_rel):
            if rel is None:
                continue
            conv = convs["__".join(rel.edge_type)]
            w_roots.append(conv.lin_r.weight)
            b_roots.append(conv.lin_l.bias)
            y_x.append(x[t])
            y_w.append(rep_param(conv.lin_l.weight, dctx))
        if not w_roots:
            return None
        ys_live = ops.linear_group(y_x, y_w)
        ys_live = replicated_to_local_many(ys_live, dctx)  # consumed by rank-local rows: gradients are summed over ranks
        it = iter(ys_live)
        ys = [None if rel is None else next(it) for rel in pb.in_rel]
        want_stats = self.training and self.use_batch_norm          # BatchNorm follows (model.py:259-261): statistics in the epilogue
        res = ops.PatientSideFn.apply(pb, len(w_roots), want_stats, x[hub], *w_roots, *b_roots, *ys)
        out = {hub: res[0]}
        out_types = [i for i, rel in enumerate(pb.out_rel) if rel is not None and pb.types[i] in by_dst]
        aggs = partial_to_replicated_many([res[1 + i] for i in out_types], dctx)     # partial neighbour sums -> all ranks
        if out_types:
            dst_x, w_r, b_l, w_l = [], [], [], []
            for i in out_types:
                conv = convs["__".join(pb.out_rel[i].edge_type)]
                dst_x.append(x[pb.types[i]])
                w_r.append(rep_param(conv.lin_r.weight, dctx))
                b_l.append(rep_param(conv.lin_l.bias, dctx))
                w_l.append(rep_param(conv.lin_l.weight, dctx))
            outs = ops.SageTypeDstGroupFn.apply(len(out_types), *dst_x, *w_r, *b_l, *aggs, *w_l)
            for i, o in zip(out_types, outs):
                out[pb.types[i]] = o
        return {dst: out[dst] for dst in by_dst if dst in out}

    def _gnn(self, x: Dict[str, torch.Tensor], gi: GraphIndex, streams: _DropoutStreams) -> Dict[str, torch.Tensor]:
        training = self.training
        for layer in range(self.num_layers):
            x = self._layer(layer, x, gi)
            p = self.dropout if layer < self.num_layers - 1 else 0.0
            nx = {}
            for nt, v in x.items():
                bn = self.batch_norms[layer][nt] if self.use_batch_norm else None
                nx[nt] = _bn_act_drop(v, bn, training, self._act, p, streams, f"fwd.l{layer}.{nt}", self.dist,
                                      self.dist is not None and nt == self.dist.sharded_type)
            x = nx
        return x

    def _streams(self) -> _DropoutStreams:
        s = _DropoutStreams(self.training and self.dropout > 0, self._seed_buffer)
        self._last_streams = s
        return s

    def _graph_index(self, data) -> GraphIndex:
        gi = graph_index(data)
        if self.dist is not None:
            globalize_degrees(gi, self.dist)       # mean onto replicated types divides by the global neighbour count
        return gi

    # ------------------------------------------------------------------------------------------
    # inference: the reference's per-patient report calls predict_lab_values twice per patient and every call recomputes
    # the full-graph forward (inference.py:92-159).  In eval mode without autograd the node embeddings only depend on
    # the graph and the parameters, so they can be kept.
    def enable_embedding_cache(self, enabled: bool = True):
        """Opt-in: in eval mode under torch.no_grad(), keep the pre-GNN and post-GNN node embeddings of the last graph and
        reuse them while no parameter or buffer has been modified (checked through the tensors' version counters)."""
        self._embedding_cache_on = bool(enabled)
        self._embedding_cache = None

    def _state_versions(self):
        """Cache key: an explicit generation counter (bumped by train(), load_state_dict() and by this package's Trainer /
        FusedAdam after every optimizer step -- they write parameters through raw pointers, also inside replayed CUDA graphs,
        where tensor version counters do not move) plus the tensors' own version counters (torch.optim, manual edits)."""
        return (self._generation,) + tuple(t._version for t in self.parameters()) + tuple(t._version for t in self.buffers())

    def bump_generation(self):
        self._generation += 1
        self._embedding_cache = None

    def train(self, mode: bool = True):
        self.bump_generation()
        return super().train(mode)

    def _eval_embeddings(self, data, gi: GraphIndex, node_types, streams):
        """(init, x) for eval mode: model.py:294 and :301 (one encode serves both, as dropout is off)."""
        use = self._embedding_cache_on and not torch.is_grad_enabled()
        key = (id(gi), self._state_versions()) if use else None
        if use and self._embedding_cache is not None and self._embedding_cache[0] == key and self._embedding_cache[1] is gi:
            return self._embedding_cache[2], self._embedding_cache[3]
        init = self._encode(node_types, streams, "enc")
        x = self._gnn(init, gi, streams)
        if use:
            self._embedding_cache = (key, gi, init, x)
        return init, x

    @torch.no_grad()
    def impute_missing(self, data, patient_ids: Optional[torch.Tensor] = None, chunk_pairs: int = 1 << 24):
        """All never-measured (patient, lab) pairs of the given patients (default: every patient) and their predicted
        normalised values -- the bulk form of inference.py:140-159 ("truly missing labs"), which the reference runs one
        patient at a time.  Returns (patient_idx int64[K], lab_idx int64[K], prediction float32[K]), ordered by patient,
        then lab.  Eval mode only; the node embeddings are computed once."""
        if self.training:
            raise _lib.B2GError("impute_missing runs in eval mode (call model.eval() first)")
        self._check_device()
        if len(self.embeddings) == 0:
            self._init_embeddings(data)
        gi = self._graph_index(data)
        dev = self._device()
        n_p, n_l = gi.node_counts["patient"], gi.node_counts["lab"]
        rel = gi.relations.get(("patient", "has_lab", "lab"))
        if rel is None:
            raise _lib.B2GError("graph has no ('patient','has_lab','lab') relation")
        ids = torch.arange(n_p, device=dev) if patient_ids is None else patient_ids.to(dev).long().unique()
        was_on, self._embedding_cache_on = self._embedding_cache_on, True
        try:
            out_p, out_l, out_v = [], [], []
            rows_per_chunk = max(1, chunk_pairs // max(n_l, 1))
            rowptr, col = rel.by_src.rowptr.long(), rel.by_src.col.long()          # by patient: measured labs of each patient
            for s0 in range(0, int(ids.numel()), rows_per_chunk):
                sub = ids[s0:s0 + rows_per_chunk]
                measured = torch.zeros((sub.numel(), n_l), dtype=torch.bool, device=dev)
                beg, cnt = rowptr[sub], rowptr[sub + 1] - rowptr[sub]
                row_of = torch.repeat_interleave(torch.arange(sub.numel(), device=dev), cnt)
                pos = torch.arange(int(cnt.sum()), device=dev) - torch.repeat_interleave(cnt.cumsum(0) - cnt, cnt) + torch.repeat_interleave(beg, cnt)
                measured[row_of, col[pos]] = True
                miss = (~measured).nonzero()
                pi, li = sub[miss[:, 0]].contiguous(), miss[:, 1].contiguous()
                out_p.append(pi)
                out_l.append(li)
                out_v.append(self.predict_lab_values(data, pi, li) if pi.numel() else torch.zeros(0, device=dev))
            return torch.cat(out_p), torch.cat(out_l), torch.cat(out_v)
        finally:
            self._embedding_cache_on = was_on
            if not was_on:
                self._embedding_cache = None          # the cache was forced on for this call only

    def _check_device(self):
        if self._device().type != "cuda":
            raise _lib.B2GError("HeteroRGCN runs only on a CUDA device (B200); call .to('cuda') first -- there is no CPU path")

    # ------------------------------------------------------------------------------------------
    def encode_nodes(self, data) -> Dict[str, torch.Tensor]:
        """model.py:206-234 (public: advanced_visualizations.py:285)."""
        self._check_device()
        if len(self.embeddings) == 0:
            self._init_embeddings(data)
        return self._encode(list(data.node_types), self._streams(), "enc")

    def forward(self, data) -> Dict[str, torch.Tensor]:
        """model.py:236-271."""
        self._check_device()
        if len(self.embeddings) == 0:
            self._init_embeddings(data)
        gi = self._graph_index(data)
        streams = self._streams()
        x = self._encode(list(data.node_types), streams, "fwd.enc")
        return self._gnn(x, gi, streams)

    def predict_lab_values(self, data, patient_indices: torch.Tensor, lab_indices: torch.Tensor) -> torch.Tensor:
        """model.py:273-335: degree-gated edge regression for (patient, lab) pairs."""
        self._check_device()
        if len(self.embeddings) == 0:
            self._init_embeddings(data)
        if not patient_indices.is_cuda or not lab_indices.is_cuda:
            # the reference's EdgeMasker keeps its index tensors on the host (train.py:86,173 -- built before data.to(device)) and
            # hands them to the model as they are; PyTorch's fancy indexing accepts that, so does the drop-in: the INDEX lists are
            # staged onto the module's device (this is input staging, not a CPU compute path)
            dev = self._device()
            patient_indices = patient_indices.to(dev, non_blocking=True)
            lab_indices = lab_indices.to(dev, non_blocking=True)
        gi = self._graph_index(data)
        node_types = list(data.node_types)
        streams = self._streams()
        training = self.training

        if training and self.dropout > 0:
            init = self._encode(node_types, streams, "init.enc")         # model.py:294
            x0 = self._encode(node_types, streams, "fwd.enc")            # model.py:251 (fresh masks, N3)
        elif not training:
            init, x = self._eval_embeddings(data, gi, node_types, streams)
            x0 = None
        else:
            bns = [self.patient_transform[1], self.patient_transform[5]]
            before = [(b.running_mean.clone(), b.running_var.clone()) for b in bns] if training else None
            init = self._encode(node_types, streams, "enc")
            x0 = init
            if training:     # second, identical running-stat update of the reference's second encode (N3)
                with torch.no_grad():
                    for b, (rm0, rv0) in zip(bns, before):
                        mom = b.momentum if b.momentum is not None else 0.1
                        # r1 = (1-m) r0 + m s  ->  r2 = (1-m) r1 + m s = (2-m) r1 - (1-m) r0
                        b.running_mean.mul_(2 - mom).sub_(rm0, alpha=1 - mom)
                        b.running_var.mul_(2 - mom).sub_(rv0, alpha=1 - mom)
                        b.num_batches_tracked += 1
        if x0 is not None:
            x = self._gnn(x0, gi, streams)

        plan = self._pair_plan(gi, patient_indices, lab_indices)
        if self.dist is not None:
            return self._predict_heads_distributed(plan, init, x, streams, patient_indices.device)
        outs = []
        if plan.low is not None:
            outs.append(self.tabular_mlp.forward_pairs(init["patient"], init["lab"], plan.low, streams, "tabular_mlp"))
        if plan.high is not None:
            outs.append(self.edge_predictor.forward_pairs(x["patient"], x["lab"], plan.high, streams, "edge_predictor"))
        if plan.m == 0:
            return torch.zeros(0, dtype=torch.float32, device=patient_indices.device)
        if plan.idx_low is None:      # a single head covers every pair, already in caller order
            return outs[0]
        return _AssembleFn.apply(outs[0], outs[1], plan.idx_low, plan.idx_high, plan.m)

    def _predict_heads_distributed(self, plan: "_PairPlan", init, x, streams, device):
        """Multi-GPU: both heads run on EVERY rank, even when this rank has no pair for one of them, because their
        replicated lab-side products carry a gradient all-reduce that all ranks must enter (a rank-dependent branch
        around a collective deadlocks)."""
        empty = plan.empty_pairs()
        low = self.tabular_mlp.forward_pairs(init["patient"], init["lab"], plan.low if plan.low is not None else empty, streams,
                                             "tabular_mlp", self.dist)
        high = self.edge_predictor.forward_pairs(x["patient"], x["lab"], plan.high if plan.high is not None else empty, streams,
                                                 "edge_predictor", self.dist)
        if plan.idx_low is not None:
            return _AssembleFn.apply(low, high, plan.idx_low, plan.idx_high, plan.m)
        main, other = (high, low) if plan.high is not None else (low, high)
        if plan.m == 0:
            main = torch.zeros(0, dtype=torch.float32, device=device)
        return main + 0.0 * (low.sum() + high.sum())       # keeps the unused head in the autograd graph (zero gradient)

    # ------------------------------------------------------------------------------------------
    def _pair_plan(self, gi: GraphIndex, pi: torch.Tensor, li: torch.Tensor) -> "_PairPlan":
        for key, plan in self._pair_plans.items():
            if plan.gi is gi and plan.same_pairs(pi, li):
                self._pair_plans.move_to_end(key)
                return plan
        plan = _PairPlan(gi, pi, li, self.degree_threshold)
        self._pair_plans[id(plan)] = plan
        while len(self._pair_plans) > 4:
            self._pair_plans.popitem(last=False)
        return plan


class _PairPlan:
    """Gate + index structures for one list of (patient, lab) pairs (model.py:305-333).  Cached per model:
    train.py re-creates equal index tensors every epoch, so lookups compare content, not identity."""

    def __init__(self, gi: GraphIndex, pi: torch.Tensor, li: torch.Tensor, threshold: int):
        lib = _lib.load()
        self.gi = gi
        self.pi, self.li = pi, li
        self.versions = (pi._version, li._version)
        self.m = int(pi.numel())
        n_p, n_l = gi.node_counts["patient"], gi.node_counts["lab"]
        pi64 = pi.contiguous() if pi.dtype == torch.int64 else pi.long()
        li64 = li.contiguous() if li.dtype == torch.int64 else li.long()
        self.low = self.high = self.idx_low = self.idx_high = None
        self.low_mask = torch.zeros(self.m, dtype=torch.uint8, device=pi.device)
        if self.m == 0:
            return
        lo = torch.stack([pi64.min(), li64.min(), pi64.max(), li64.max()]).tolist()      # one host sync (another follows below)
        if lo[0] < 0 or lo[1] < 0 or lo[2] >= n_p or lo[3] >= n_l:
            raise IndexError(f"patient / lab indices out of range: patients in [{lo[0]}, {lo[2]}] of {n_p}, labs in [{lo[1]}, {lo[3]}] of {n_l}")
        if gi.patient_lab_degree is None:
            raise _lib.B2GError("graph has no ('patient','has_lab','lab') relation: cannot compute the degree gate")
        _lib.check(lib.b2g_degree_gate(gi.patient_lab_degree.data_ptr(), pi64.data_ptr(), self.m, int(threshold),
                                       self.low_mask.data_ptr(), _stream()), "b2g_degree_gate")
        n_low = int(self.low_mask.sum().item())
        if n_low == 0:
            self.high = PairIndex(pi64, li64, n_p, n_l)
        elif n_low == self.m:
            self.low = PairIndex(pi64, li64, n_p, n_l)
        else:
            mask = self.low_mask.bool()
            self.idx_low = mask.nonzero().squeeze(1)
            self.idx_high = (~mask).nonzero().squeeze(1)
            self.low = PairIndex(pi64[self.idx_low], li64[self.idx_low], n_p, n_l)
            self.high = PairIndex(pi64[self.idx_high], li64[self.idx_high], n_p, n_l)

    def empty_pairs(self) -> PairIndex:
        if getattr(self, "_empty", None) is None:
            e = torch.zeros(0, dtype=torch.int64, device=self.pi.device)
            self._empty = PairIndex(e, e, self.gi.node_counts["patient"], self.gi.node_counts["lab"])
        return self._empty

    def same_pairs(self, pi, li) -> bool:
        if pi is self.pi and li is self.li and (pi._version, li._version) == self.versions:
            return True
        if pi.numel() != self.m or pi.dtype != self.pi.dtype or li.dtype != self.li.dtype or pi.device != self.pi.device:
            return False
        if (self.pi._version, self.li._version) != self.versions:
            return False
        return bool(torch.equal(pi, self.pi)) and bool(torch.equal(li, self.li))


class _AssembleFn(torch.autograd.Function):
    """predictions[low] = tabular head, predictions[~low] = GNN head (model.py:317-333)."""

    @staticmethod
    def forward(ctx, out_low, out_high, idx_low, idx_high, m):
        lib = _lib.load()
        out_low, out_high = out_low.contiguous(), out_high.contiguous()
        pred = torch.empty(m, dtype=torch.float32, device=out_low.device)
        _lib.check(lib.b2g_scatter_values(out_low.data_ptr(), idx_low.data_ptr(), idx_low.numel(), pred.data_ptr(), _stream()),
                   "b2g_scatter_values")
        _lib.check(lib.b2g_scatter_values(out_high.data_ptr(), idx_high.data_ptr(), idx_high.numel(), pred.data_ptr(), _stream()),
                   "b2g_scatter_values")
        ctx.idx = (idx_low, idx_high)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        lib = _lib.load()
        idx_low, idx_high = ctx.idx
        dpred = dpred.contiguous()
        dl = torch.empty(idx_low.numel(), dtype=torch.float32, device=dpred.device)
        dh = torch.empty(idx_high.numel(), dtype=torch.float32, device=dpred.device)
        _lib.check(lib.b2g_gather_values(dpred.data_ptr(), idx_low.data_ptr(), idx_low.numel(), dl.data_ptr(), _stream()),
                   "b2g_gather_values")
        _lib.check(lib.b2g_gather_values(dpred.data_ptr(), idx_high.data_ptr(), idx_high.numel(), dh.data_ptr(), _stream()),
                   "b2g_gather_values")
        return dl, dh, None, None, None


# ----------------------------------------------------------------------------------------------------
def build_model(config: Dict, metadata: Tuple, patient_feature_dim: int):
    """model.py:523-572.  Reads exactly config['model'][architecture, hidden_dim, num_layers, dropout,
    use_batch_norm, activation]; constructible before the graph is seen."""
    mc = config["model"]
    arch = mc["architecture"]
    if arch == "RGCN":
        model = HeteroRGCN(metadata=metadata, hidden_dim=mc["hidden_dim"], num_layers=mc["num_layers"], dropout=mc["dropout"],
                           patient_feature_dim=patient_feature_dim, use_batch_norm=mc["use_batch_norm"],
                           activation=mc["activation"])
        logging.info("Built HeteroRGCN model")
    elif arch == "HGT":
        raise NotImplementedError("HGT is unreachable in the reference's shipped configuration (needs patient.x, which "
                                  "graph_build.py no longer sets) and is outside the accelerated hot path")
    else:
        raise ValueError(f"Unknown architecture: {arch}")
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    logging.info(f"Model has {n:,} trainable parameters")
    return model


def compute_regression_loss(predictions: torch.Tensor, targets: torch.Tensor, loss_type: str = "mae") -> torch.Tensor:
    """model.py:579-612: unweighted l1 / mse / huber mean, one fused deterministic reduction."""
    if loss_type not in ops.LOSS_KINDS:
        raise ValueError(f"Unknown loss type: {loss_type}")
    return ops.weighted_loss(predictions, targets, None, None, None, loss_type)

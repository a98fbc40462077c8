"""ORACLE / TEST INFRASTRUCTURE ONLY.  Pure-torch restatement of PyG's SAGEConv / HeteroConv.

Semantics restated (PyG 2.3.x `torch_geometric/nn/conv/sage_conv.py`, `hetero_conv.py`,
`torch_geometric/utils/scatter.py`; call sites /root/reference/src/model.py:125-131,256):

  SAGEConv(in, out, aggr='mean'): lin_l = Linear(in, out, bias=True) applied to the mean of the
      neighbour messages, lin_r = Linear(in, out, bias=False) applied to the destination's own
      features; messages are x_src rows selected by edge_index[0], summed into edge_index[1] with
      scatter_add, divided by the neighbour count clamped to >= 1; no normalisation/projection.
  HeteroConv(convs, aggr='sum'): one conv per edge type, visited in constructor (insertion)
      order; results that share a destination type are combined with stack(...).sum(0);
      ModuleDict key is '__'.join(edge_type) (PyG 2.3 naming).
"""
from collections import defaultdict

import torch
import torch.nn as tnn

Linear = tnn.Linear


def _scatter_mean(src, index, dim_size):
    total = src.new_zeros((dim_size, src.size(1))).index_add_(0, index, src)
    count = src.new_zeros(dim_size).index_add_(0, index, src.new_ones(index.numel()))
    return total / count.clamp(min=1).unsqueeze(-1)


class SAGEConv(tnn.Module):
    def __init__(self, in_channels, out_channels, aggr="mean", **kwargs):
        super().__init__()
        if aggr != "mean":
            raise NotImplementedError("shim restates aggr='mean' only (model.py:128)")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.lin_l = tnn.Linear(in_channels[0], out_channels, bias=True)
        self.lin_r = tnn.Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index):
        x_src, x_dst = x if isinstance(x, (tuple, list)) else (x, x)
        messages = x_src.index_select(0, edge_index[0])
        agg = _scatter_mean(messages, edge_index[1], x_dst.size(0))
        return self.lin_l(agg) + self.lin_r(x_dst)


class HeteroConv(tnn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        if aggr != "sum":
            raise NotImplementedError("shim restates aggr='sum' only (model.py:131)")
        self.aggr = aggr
        self._edge_types = list(convs.keys())
        self.convs = tnn.ModuleDict({"__".join(k): v for k, v in convs.items()})

    def forward(self, x_dict, edge_index_dict):
        outs = defaultdict(list)
        for edge_type in self._edge_types:
            if edge_type not in edge_index_dict:
                continue
            src, _, dst = edge_type
            if src not in x_dict or dst not in x_dict:
                continue
            conv = self.convs["__".join(edge_type)]
            outs[dst].append(conv((x_dict[src], x_dict[dst]), edge_index_dict[edge_type]))
        return {k: (v[0] if len(v) == 1 else torch.stack(v, 0).sum(0)) for k, v in outs.items()}


def _unavailable(name):
    class _Stub:  # imported by model.py:23-24 but never constructed on the RGCN path
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the restated hot path")
    _Stub.__name__ = name
    return _Stub


GCNConv = _unavailable("GCNConv")
GATConv = _unavailable("GATConv")
HGTConv = _unavailable("HGTConv")


def to_hetero(*a, **k):
    raise NotImplementedError("to_hetero is outside the restated hot path")

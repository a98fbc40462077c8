"""Minimal heterogeneous-graph container with the duck-typed surface the model consumes.

The reference stores its graph in a PyG ``HeteroData`` (/root/reference/src/graph_build.py:148-261);
the drop-in accepts *any* object with that surface (last row): ``node_types``,
``edge_types``, ``data[nt].num_nodes``, ``data[(s, r, d)].edge_index`` / ``.edge_attr``,
``edge_index_dict`` and ``.to(device)``.  This class is what ``synth.py`` / ``bench.py`` build when PyG
is not installed (it is not, on the GPU box); a real PyG ``HeteroData`` works equally well.
"""
from __future__ import annotations

import torch


class AttrStore(dict):
    """dict whose keys are also attributes (``store.num_nodes``, ``store.edge_index``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


class HeteroGraph:
    """Insertion-ordered node and edge stores (order fixes HeteroConv's summation order,
    /root/reference/src/graph_build.py:216-247)."""

    def __init__(self):
        self.__dict__["_nodes"] = {}
        self.__dict__["_edges"] = {}
        self.__dict__["_extra"] = {}

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 3:
                raise KeyError(f"edge type must be (src, rel, dst), got {key!r}")
            return self._edges.setdefault(tuple(key), AttrStore())
        return self._nodes.setdefault(key, AttrStore())

    def __setattr__(self, name, value):
        self._extra[name] = value

    def __getattr__(self, name):
        extra = self.__dict__["_extra"]
        if name in extra:
            return extra[name]
        raise AttributeError(name)

    @property
    def node_types(self):
        return list(self._nodes)

    @property
    def edge_types(self):
        return list(self._edges)

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def edge_index_dict(self):
        return {k: v["edge_index"] for k, v in self._edges.items() if "edge_index" in v}

    def to(self, device, non_blocking: bool = False):
        for group in (self._nodes, self._edges):
            for store in group.values():
                for k, v in list(store.items()):
                    if torch.is_tensor(v):
                        store[k] = v.to(device, non_blocking=non_blocking)
        return self

    def cpu(self):
        return self.to("cpu")

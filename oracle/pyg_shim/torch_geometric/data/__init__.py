"""ORACLE / TEST INFRASTRUCTURE ONLY.  Duck-typed stand-in for torch_geometric.data.HeteroData.

Only the surface the reference path uses is restated: item access by node-type string or
(src, rel, dst) tuple, attribute stores, `node_types` / `edge_types` in insertion order,
`edge_index_dict`, `metadata()`, and `.to(device)` (returns self, moves every tensor).
"""
import torch


class _Store(dict):
    """Attribute store: `data['patient'].num_nodes = 7`, `data[et].edge_index = t`."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:  # pragma: no cover - mirrors AttributeError of PyG storages
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_node_stores", {})
        object.__setattr__(self, "_edge_stores", {})
        object.__setattr__(self, "_globals", {})

    # -- item access -----------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 3:
                raise KeyError(key)
            return self._edge_stores.setdefault(tuple(key), _Store())
        return self._node_stores.setdefault(key, _Store())

    def __setattr__(self, name, value):
        self._globals[name] = value

    def __getattr__(self, name):
        g = object.__getattribute__(self, "_globals")
        if name in g:
            return g[name]
        raise AttributeError(name)

    # -- metadata --------------------------------------------------------------------
    @property
    def node_types(self):
        return list(self._node_stores.keys())

    @property
    def edge_types(self):
        return list(self._edge_stores.keys())

    def metadata(self):
        return self.node_types, self.edge_types

    @property
    def edge_index_dict(self):
        return {k: s["edge_index"] for k, s in self._edge_stores.items() if "edge_index" in s}

    @property
    def num_nodes(self):
        return sum(int(s["num_nodes"]) for s in self._node_stores.values() if "num_nodes" in s)

    # -- device movement ---------------------------------------------------------------
    def to(self, device, *args, **kwargs):
        for stores in (self._node_stores, self._edge_stores):
            for s in stores.values():
                for k, v in list(s.items()):
                    if torch.is_tensor(v):
                        s[k] = v.to(device, *args, **kwargs)
        return self

    def cpu(self):
        return self.to("cpu")

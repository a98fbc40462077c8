"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference modules on the PyG shim.

Only usable where /root/reference exists (the build container); nothing executed on the GPU box may
import this.  Patches applied *around* (never inside) the reference, SURVEY.md notes N7/N9:
  * optim.lr_scheduler.ReduceLROnPlateau swallows the removed ``verbose`` kwarg (train.py:279-285);
  * ``train.time.time`` is replaced by a deterministic counter so the wall-clock reseeding in
    EdgeMasker.get_masked_data (train.py:156) is reproducible.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_SRC = "/root/reference/src"
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "model.py"))


class FakeClock:
    """time.time() stand-in: returns base, base+1, ... on successive calls."""

    def __init__(self, base=1_700_000_000):
        self.now = base

    def time(self):
        self.now += 1
        return float(self.now)


def load_reference(clock: FakeClock | None = None):
    """Returns (model_module, train_module) = the reference's src/model.py, src/train.py."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    for p in (REFERENCE_SRC, SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.optim.lr_scheduler as lrs
    if not getattr(lrs.ReduceLROnPlateau, "_b2g_patched", False):
        base = lrs.ReduceLROnPlateau

        class ReduceLROnPlateau(base):  # noqa: D401 - N9
            _b2g_patched = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)
        lrs.ReduceLROnPlateau = ReduceLROnPlateau
        torch.optim.lr_scheduler.ReduceLROnPlateau = ReduceLROnPlateau
    import importlib
    ref_model = importlib.import_module("model")
    ref_train = importlib.import_module("train")
    assert ref_model.__file__.startswith(REFERENCE_SRC), ref_model.__file__
    if clock is not None:
        import time as _time
        fake = types.ModuleType("time")
        fake.__dict__.update(_time.__dict__)
        fake.time = clock.time
        ref_train.time = fake
    return ref_model, ref_train


def to_shim_data(graph):
    """Copy any HeteroData-like object into the shim's HeteroData (tensors shared)."""
    from torch_geometric.data import HeteroData
    d = HeteroData()
    for nt in graph.node_types:
        d[nt].num_nodes = int(graph[nt].num_nodes)
    for et in graph.edge_types:
        d[et].edge_index = graph[et].edge_index
        if "edge_attr" in graph[et]:
            d[et].edge_attr = graph[et].edge_attr
    return d


DEFAULT_CONFIG = {
    "model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2,
              "activation": "relu", "use_batch_norm": True},
    "train": {"mask_fraction": 0.2, "train_split": 0.7, "val_split": 0.15, "test_split": 0.15,
              "loss": "mae", "epochs": 100, "early_stopping_patience": 15,
              "optimizer": {"type": "adam", "lr": 0.001, "weight_decay": 0.00001},
              "lr_scheduler": {"enabled": True, "type": "reduce_on_plateau", "factor": 0.5, "patience": 10},
              "seed": 42},
}


def make_config(dropout=0.2, loss="mae", hidden_dim=128, num_layers=2):
    import copy
    c = copy.deepcopy(DEFAULT_CONFIG)
    c["model"].update(dropout=dropout, hidden_dim=hidden_dim, num_layers=num_layers)
    c["train"]["loss"] = loss
    return c

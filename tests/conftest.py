import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    mod = importlib.import_module("multi-modal-gnn_b200")
    sys.modules.setdefault("mmgnn_b200", mod)
    return mod


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def golden_graph(blob):
    """Rebuild (node_counts, edge_types, edge_index_dict[int64], edge_attr) from a fixture."""
    g = blob["graph"]
    ets = [tuple(e) for e in g["edge_types"]]
    eid = {et: g["edge_index"]["__".join(et)].long() for et in ets}
    return dict(g["num_nodes"]), ets, eid, g["edge_attr"]


@pytest.fixture(scope="session")
def golden_tiny_mae():
    return load_golden("tiny_mae")


@pytest.fixture(scope="session")
def golden_tiny_mse():
    return load_golden("tiny_mse")


@pytest.fixture(scope="session")
def golden_c1():
    return load_golden("c1_mae")

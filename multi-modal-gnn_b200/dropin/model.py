"""Drop-in for the reference's ``src/model.py``: put this directory first on ``sys.path`` / ``PYTHONPATH`` and
``from model import build_model, compute_regression_loss`` (train.py:29, evaluate.py, inference.py) resolves to
the B200 implementation.  See INTEGRATION.md."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_impl = importlib.import_module("multi-modal-gnn_b200.model")

HeteroRGCN = _impl.HeteroRGCN
EdgeRegressionHead = _impl.EdgeRegressionHead
build_model = _impl.build_model
compute_regression_loss = _impl.compute_regression_loss

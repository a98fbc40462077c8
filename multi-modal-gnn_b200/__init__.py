"""B200-native replacement for the hetero-GNN training / imputation hot path of
AdalineL/Multi-Modal-GNN (src/model.py + the train_epoch/validate part of src/train.py).

The directory name contains a hyphen, so import it with
``importlib.import_module("multi-modal-gnn_b200")`` (``__graft_entry__.load_package()`` does that and
aliases it as ``mmgnn_b200``), or put ``multi-modal-gnn_b200/dropin`` first on ``sys.path`` so that the
reference's ``from model import build_model, compute_regression_loss`` resolves to this package.
"""
from .heterodata import HeteroGraph  # noqa: F401
from . import synth  # noqa: F401

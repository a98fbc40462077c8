"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Minimal pure-torch restatement of the parts of PyTorch-Geometric (torch-geometric>=2.3.0,
un-vendored dependency named in /root/reference/requirements.txt:19, no lock file) that the
reference hot path touches:

  * torch_geometric.nn.SAGEConv / HeteroConv   (used at /root/reference/src/model.py:23,125-131,256)
  * torch_geometric.data.HeteroData            (used at /root/reference/src/train.py:27,
                                                /root/reference/src/graph_build.py:24,148)

With this package first on sys.path the reference's src/model.py and src/train.py import and
run UNMODIFIED (oracle/ref_harness.py).  PyG itself is not installed in the build container
nor on the GPU box, so these semantics are restated from the published PyG 2.3/2.4 algorithm
("parity unpinned" by any PyG golden vector -- see DESIGN.md section Oracle).
"""
__version__ = "2.3.0+shim"
from . import nn, data, transforms  # noqa: F401

"""Graph ingest (multi-modal-gnn_b200/ingest.py + csrc/ingest.cu) against what the UNMODIFIED reference graph builder produced
for the same tables (tests/golden/ingest_small.pt, written by oracle/make_golden_ingest.py from
/root/reference/src/graph_build.py::build_heterogeneous_graph): bit-exact node counts, type orders, COO edges, attributes and
data.indexers; then the ingested graph drives the model."""
import importlib
import os

import pytest
import torch

PKG = "multi-modal-gnn_b200"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ingest_small.pt")


def test_vocabulary_is_first_occurrence_order():
    I = importlib.import_module(PKG + ".ingest")
    import numpy as np
    v = I.Vocabulary(np.array([30.0, 10.0, 30.0, 20.0, 10.0]))            # floats are compared as ints (NodeIndexer.add)
    assert v.ids.tolist() == [30, 10, 20] and v.as_dicts()["id_to_index"] == {"30": 0, "10": 1, "20": 2}
    s = I.Vocabulary(np.array(["V58", "250", "V58", "E88"], dtype=object))
    assert s.ids.tolist() == ["V58", "250", "E88"] and s.as_dicts()["index_to_id"] == {0: "V58", 1: "250", 2: "E88"}


@pytest.mark.gpu
def test_ingest_matches_reference_graph_builder():
    I = importlib.import_module(PKG + ".ingest")
    M = importlib.import_module(PKG + ".model")
    blob = torch.load(GOLDEN, weights_only=False)
    t = blob["tables"]
    g = I.build_graph_from_tables(t["cohort"], t["labs"], t["diagnoses"], t["medications"], device="cuda:0")
    assert g.node_types == blob["node_types"] and [tuple(e) for e in g.edge_types] == blob["edge_types"]
    for nt, n in blob["num_nodes"].items():
        assert int(g[nt].num_nodes) == n
    for et in blob["edge_types"]:
        key = "__".join(et)
        ei = g[et].edge_index
        assert ei.dtype == torch.int64 and ei.is_cuda and torch.equal(ei.cpu(), blob["edge_index"][key]), key     # bit-exact, row order
        if key in blob["edge_attr"]:
            ea = g[et].edge_attr
            assert ea.dtype == torch.float32 and torch.equal(ea.cpu(), blob["edge_attr"][key]), key
        else:
            assert "edge_attr" not in g[et]
    for nt, ref in blob["indexers"].items():                      # graph_build.py:254-260 (read by inference.py:330-331)
        assert g.indexers[nt]["id_to_index"] == ref["id_to_index"] and g.indexers[nt]["index_to_id"] == ref["index_to_id"], nt
    # the ingested graph is a valid input of the model (CSR build validates the index ranges)
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.0, "use_batch_norm": True, "activation": "relu"}}
    torch.manual_seed(0)
    model = M.build_model(cfg, (g.node_types, g.edge_types), None).to("cuda:0")
    model.eval()
    ei = g["patient", "has_lab", "lab"].edge_index
    with torch.no_grad():
        pred = model.predict_lab_values(g, ei[0], ei[1])
    assert pred.shape == (ei.shape[1],) and bool(torch.isfinite(pred).all())


@pytest.mark.gpu
def test_edges_from_rows_is_a_stable_filter():
    I = importlib.import_module(PKG + ".ingest")
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(3)
    for m in (0, 1, 2047, 2048, 2049, 1_000_003):
        a = torch.randint(-1, 50, (m,), generator=gen, dtype=torch.int32)
        b = torch.randint(-1, 7, (m,), generator=gen, dtype=torch.int32)
        v = torch.randn(m, generator=gen)
        ei, ea = I.edges_from_rows(a.to(dev), b.to(dev), v.to(dev))
        keep = (a >= 0) & (b >= 0)
        assert torch.equal(ei.cpu(), torch.stack([a[keep].long(), b[keep].long()])) and torch.equal(ea.cpu().squeeze(1), v[keep]), m

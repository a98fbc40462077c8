"""CPU: host-side logic of the product that needs no GPU -- synthetic graph contract, EdgeMasker bit-exactness
with the reference's split arithmetic, lab weights, model construction / state_dict layout."""
import importlib

import pytest
import torch

from oracle import hetero_rgcn_ref as R

PKG = "multi-modal-gnn_b200"


def test_synthetic_graph_contract(pkg):
    g = pkg.synth.make_graph("C1")
    assert g.node_types == ["patient", "lab", "diagnosis", "medication"]
    assert [et[1] for et in g.edge_types] == ["has_lab", "has_lab_rev", "has_diagnosis", "has_diagnosis_rev",
                                              "has_medication", "has_medication_rev"]
    for fwd, rev in zip(g.edge_types[::2], g.edge_types[1::2]):
        assert torch.equal(g[fwd].edge_index.flip(0), g[rev].edge_index)
        ei = g[fwd].edge_index
        key = ei[0] * 100000 + ei[1]
        assert key.unique().numel() == key.numel(), "pairs must be unique"
        assert ei.dtype == torch.int64
    ei = g["patient", "has_lab", "lab"].edge_index
    k = ei[1] * 1834 + ei[0]
    assert bool((k[1:] > k[:-1]).all()), "has_lab is grouped by lab, then patient"
    ea = g["patient", "has_lab", "lab"].edge_attr
    assert ea.shape == (61484, 1) and ea.dtype == torch.float32 and float(ea.abs().max()) <= 5.0
    assert "edge_attr" not in g["patient", "has_diagnosis", "diagnosis"]
    g2 = pkg.synth.make_graph("C1")
    assert torch.equal(g2["patient", "has_lab", "lab"].edge_index, ei), "generator must be deterministic"
    deg = torch.bincount(ei[0], minlength=1834)
    assert int((deg < 6).sum()) > 0 and int(deg.min()) >= 1


def test_edge_masker_matches_reference_split_arithmetic(pkg, golden_c1):
    T = importlib.import_module(PKG + ".trainer")
    g = pkg.synth.make_graph("C1")
    masker = T.EdgeMasker(g, 0.7, 0.15, 0.15, 0.2, 42)
    sp = golden_c1["split"]                      # produced by the unmodified reference EdgeMasker
    assert torch.equal(masker.train_mask, sp["train"]) and torch.equal(masker.val_mask, sp["val"])
    assert torch.equal(masker.test_mask, sp["test"])
    assert (int(masker.train_mask.sum()), int(masker.val_mask.sum()), int(masker.test_mask.sum())) == (43038, 9222, 9224)
    sup = masker.supervision_mask("train", seed=golden_c1["sup_seed"])
    assert torch.equal(sup, golden_c1["sup_mask"])
    assert bool(masker.supervision_mask("val").all())
    with pytest.raises(ValueError):
        masker.split_mask("nope")
    ei, ev, mask, s = masker.get_masked_data("val")
    assert ei.shape == (2, 9222) and ev.shape == (9222,) and mask is masker.val_mask
    with pytest.raises(AssertionError):
        T.EdgeMasker(g, 0.7, 0.2, 0.2)


def test_lab_weights_match_reference(pkg, golden_c1):
    T = importlib.import_module(PKG + ".trainer")
    g = pkg.synth.make_graph("C1")
    ei = g["patient", "has_lab", "lab"].edge_index
    ea = g["patient", "has_lab", "lab"].edge_attr
    tr = golden_c1["split"]["train"]
    w = T.compute_lab_weights(ei[1][tr], ea[tr].squeeze(-1), 50)
    torch.testing.assert_close(w, golden_c1["lab_weights"], rtol=1e-5, atol=1e-7)
    assert abs(float(w.sum()) - 50.0) < 1e-3
    # a lab with a single sample gets variance 1 (train.py:313-319)
    w2 = T.compute_lab_weights(torch.tensor([0, 0, 0, 1]), torch.tensor([1.0, 2.0, 4.0, 9.0]), 3)
    ref = R.lab_weights(torch.tensor([0, 0, 0, 1]), torch.tensor([1.0, 2.0, 4.0, 9.0]), 3)
    torch.testing.assert_close(w2, ref, rtol=1e-6, atol=1e-7)


def test_model_state_dict_layout(pkg, golden_c1):
    M = importlib.import_module(PKG + ".model")
    cfg = {"model": {"architecture": "RGCN", "hidden_dim": 128, "num_layers": 2, "dropout": 0.2, "use_batch_norm": True,
                     "activation": "relu"}}
    model = M.build_model(cfg, (pkg.synth.NODE_TYPES, pkg.synth.EDGE_TYPES), None)
    assert sum(p.numel() for p in model.parameters()) == 483970               # KA-1, before the lazy tables exist
    assert model.degree_threshold == 6 and model.hidden_dim == 128 and model.num_layers == 2 and model.dropout == 0.2
    g = pkg.synth.make_graph("C1")
    model._init_embeddings(g)
    model._init_embeddings(g)                                                  # idempotent
    assert sum(p.numel() for p in model.parameters()) == 752514
    assert list(model.state_dict().keys()) == golden_c1["state_keys"]          # the reference's 108 keys, same order
    model.load_state_dict(golden_c1["state_before"])
    assert model.embeddings["lab"].weight.shape == (50, 128)
    with pytest.raises(NotImplementedError):
        M.build_model({"model": dict(cfg["model"], architecture="HGT")}, (pkg.synth.NODE_TYPES, pkg.synth.EDGE_TYPES), None)


def test_new_entry_points_have_no_cpu_path(pkg):
    """optim.FusedAdam, metrics.evaluate_predictions and model.impute_missing run only on the CUDA kernels: CPU tensors
    must raise (there is no fallback to torch / numpy), and the peer communicator is never built without a CUDA device."""
    L = importlib.import_module(PKG + "._lib")
    O = importlib.import_module(PKG + ".optim")
    MX = importlib.import_module(PKG + ".metrics")
    D = importlib.import_module(PKG + ".dist")
    p = torch.zeros(8, requires_grad=True)
    opt = O.FusedAdam([p], lr=1e-3, weight_decay=1e-5)
    assert set(torch.optim.Adam([torch.zeros(1)]).defaults) <= set(opt.defaults)      # state_dict layout of torch.optim.Adam
    opt.step()                                   # no gradient anywhere: nothing to do, nothing raised (note N8)
    assert len(opt.state[p]) == 0
    p.grad = torch.ones(8)
    with pytest.raises(L.B2GError):
        opt.step()
    with pytest.raises(L.B2GError):
        MX.evaluate_predictions(torch.zeros(4), torch.zeros(4), torch.zeros(4, dtype=torch.int64), 2)
    # _metrics_from_sums == evaluate.py:36-82 on sufficient statistics (sklearn's corner cases included)
    m = MX._metrics_from_sums(4, 2.0, 2.0, 10.0, 30.0, 0.5, 4)           # targets 1,2,3,4: SS_tot = 30 - 100/4 = 5
    assert abs(m["mae"] - 0.5) < 1e-12 and abs(m["rmse"] - 0.5 ** 0.5) < 1e-12 and abs(m["r2"] - 0.6) < 1e-12
    assert abs(m["mape"] - 12.5) < 1e-12
    assert MX._metrics_from_sums(3, 0.0, 0.0, 6.0, 12.0, 0.0, 3)["r2"] == 1.0      # constant targets, perfect predictions
    assert MX._metrics_from_sums(3, 1.0, 1.0, 6.0, 12.0, 0.0, 3)["r2"] == 0.0      # constant targets, imperfect
    import math
    assert math.isnan(MX._metrics_from_sums(0, 0, 0, 0, 0, 0, 0)["mae"])
    # flat gradient buffer: padded to the peer all-reduce's 16-byte unit
    flat = D._flat_padded([torch.ones(3), torch.ones(6)])
    assert flat.numel() == 12 and float(flat.sum()) == 9.0


def test_c4_shard_spec_is_one_eighth_of_c4(pkg):
    c4, sh = pkg.synth.SPECS["C4"], pkg.synth.SPECS["C4s8"]
    assert sh.n_patient * 8 == c4.n_patient and sh.e_lab * 8 == c4.e_lab and sh.e_dx * 8 == c4.e_dx and sh.e_med * 8 == c4.e_med
    assert (sh.n_lab, sh.n_dx, sh.n_med, sh.low_degree_frac) == (c4.n_lab, c4.n_dx, c4.n_med, c4.low_degree_frac)
    assert sh.directed_edges_per_layer * 8 == c4.directed_edges_per_layer


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` is CPU-only (the oracle port on a bounded sample): run it here and check the JSON contract."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_edges_per_sec_fwd_bwd_per_heteroconv_step" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "C4s8" and "sample" in d["config"]


def test_documented_switches_match_the_source():
    """Every B2G_* environment variable read by the package is listed in INTEGRATION.md section 6 (and vice versa)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    found = set()
    pkg = os.path.join(root, "multi-modal-gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                found |= set(re.findall(r'(?:environ\.get\(|getenv\()"(B2G_[A-Z0-9_]+)"', txt))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    listed = set(re.findall(r"`(B2G_[A-Z0-9_]+)`", doc.split("## 6. Switches", 1)[1]))
    assert found == listed, (sorted(found - listed), sorted(listed - found))


def test_layer_traffic_tool_reproduces_the_committed_summary():
    """tools/layer_traffic.py over the committed ncu metrics pass gives the numbers bench.py reports as roofline.traffic."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r2_final_layer_traffic_c4s8.csv")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "layer_traffic.py"), csv_path, "C4s8"], capture_output=True, text=True, check=True)
    got = json.loads(out.stdout)["hetero_layer_fwd_bwd:C4s8"]
    want = json.load(open(os.path.join(root, "profiles", "r2_traffic.json")))["hetero_layer_fwd_bwd:C4s8"]
    assert got["dram_bytes_per_layer"] == want["dram_bytes_per_layer"] and got["n_kernels"] == want["n_kernels"]
    assert got["bytes_min"] == 5 * 1250000 * 128 * 4 + 2 * (12500000 + 1102500 + 3238750) * 4 + 6 * 1250001 * 4
    assert 1.0 < got["ratio_to_bytes_min"] < 1.5          # "ncu dram__bytes per layer <= 1.5x bytes_min"

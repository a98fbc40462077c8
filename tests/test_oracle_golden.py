"""CPU: pins the oracle restatement (oracle/hetero_rgcn_ref.py) to vectors produced by the UNMODIFIED
reference (tests/golden/, made by oracle/make_golden.py), and checks the reference's known-answer
facts KA-1..KA-5."""
import pytest
import torch

from oracle import hetero_rgcn_ref as R
from conftest import golden_graph

RTOL = 2e-5   # fp32 CPU vs fp32 CPU, different summation order only


def _pre_bn_bias(k):
    return k in ("patient_transform.0.bias", "patient_transform.4.bias") or k.endswith("lin_l.bias")


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize("fixture", ["golden_tiny_mae", "golden_tiny_mse", "golden_c1"])
def test_oracle_train_step_matches_reference(fixture, request):
    blob = request.getfixturevalue(fixture)
    counts, ets, eid, attr = golden_graph(blob)
    sd = _clone(blob["state_before"])
    ei = eid[("patient", "has_lab", "lab")]
    tr = blob["split"]["train"]
    pi, li, tgt = ei[0][tr], ei[1][tr], attr[tr].squeeze(-1)
    # lab weights (train.py:295-330)
    w = R.lab_weights(li, tgt, counts["lab"])
    torch.testing.assert_close(w, blob["lab_weights"], rtol=1e-6, atol=1e-7)
    assert abs(float(w.sum()) - counts["lab"]) < 1e-3                       # KA-4
    sup = R.supervision_mask(int(tr.sum()), 0.2, blob["sup_seed"])
    assert torch.equal(sup, blob["sup_mask"])
    loss, pred, grads = R.train_step_grads(sd, counts, ets, eid, pi, li, tgt, sup, w, blob["loss_fn"], p_drop=0.0)
    torch.testing.assert_close(pred, blob["pred_train"], rtol=RTOL, atol=2e-6)
    assert abs(float(loss) - blob["loss_train"]) <= RTOL * abs(blob["loss_train"])
    # gradient structure: N8 dead branches have no gradient at all
    none_keys = sorted(k for k, v in grads.items() if v is None)
    assert none_keys == blob["grad_is_none"]
    for k, gn in blob["grad_norm"].items():
        if _pre_bn_bias(k):     # true gradient is exactly 0 (BatchNorm removes the mean): rounding noise only
            assert gn < 1e-6 and float(grads[k].double().norm()) < 1e-6, k
            continue
        assert abs(float(grads[k].double().norm()) - gn) <= 1e-4 * gn + 1e-9, k
    for k, gref in blob["grads"].items():
        torch.testing.assert_close(grads[k], gref, rtol=1e-4, atol=1e-7 + 1e-4 * float(gref.abs().max()), msg=k)
    # BN running stats after one train step: patient MLP BNs updated twice (N3)
    for k, v in blob["after_buffers"].items():
        if v.dtype == torch.long:
            assert int(sd[k]) == int(v), k
        else:
            torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-7, msg=k)
    assert int(sd["patient_transform.1.num_batches_tracked"]) == 2
    assert int(sd["batch_norms.0.patient.num_batches_tracked"]) == 1


def test_oracle_eval_products(golden_tiny_mae):
    blob = golden_tiny_mae
    counts, ets, eid, attr = golden_graph(blob)
    sd = _clone(blob["state_before"])
    ei = eid[("patient", "has_lab", "lab")]
    # eval-mode encode / forward at the *before* parameters differ from the fixture (post-step state);
    # rebuild the post-step state from the stored pieces we have: only buffers + 4 params are stored,
    # so check the eval path on the train-mode replay instead: pred_train above, and degree/gate here.
    deg = R.patient_lab_degree(ei, counts["patient"])
    assert deg.dtype == torch.int64 and torch.equal(deg, blob["degree"])
    assert int((deg < R.DEGREE_THRESHOLD).sum()) > 0          # the gate is exercised by this graph


def test_known_answers_c1(golden_c1):
    blob = golden_c1
    counts, ets, eid, attr = golden_graph(blob)
    # KA-5 graph shape
    assert counts == {"patient": 1834, "lab": 50, "diagnosis": 114, "medication": 100}
    assert [eid[e].shape[1] for e in ets[::2]] == [61484, 5421, 15933]
    # KA-2 split sizes
    sp = blob["split"]
    assert (int(sp["train"].sum()), int(sp["val"].sum()), int(sp["test"].sum())) == (43038, 9222, 9224)
    masks = R.split_masks(61484, 0.7, 0.15, 42)
    for m, k in zip(masks, ("train", "val", "test")):
        assert torch.equal(m, sp[k])
    # KA-1 parameter counts
    assert blob["n_params_before_tables"] == 483970 and blob["optimizer_param_count"] == 483970
    sd = R.init_state(counts, ets)
    assert len(sd) == 108 and list(sd.keys()) == blob["state_keys"]
    n_train = sum(sd[k].numel() for k in R.trainable_keys(sd))
    assert n_train == 483970
    assert sum(sd[k].numel() for k in R.trainable_keys(sd, include_tables=True)) == 752514
    # N8: dead last-layer branches
    assert len(blob["grad_is_none"]) == 10


def test_regression_loss_variants():
    p, t = torch.tensor([0.0, 2.0, -3.0]), torch.tensor([0.5, 0.0, 0.0])
    assert abs(float(R.regression_loss(p, t, "mae")) - (0.5 + 2 + 3) / 3) < 1e-6
    assert abs(float(R.regression_loss(p, t, "mse")) - (0.25 + 4 + 9) / 3) < 1e-6
    assert abs(float(R.regression_loss(p, t, "huber")) - float(torch.nn.functional.huber_loss(p, t))) < 1e-6
    with pytest.raises(ValueError):
        R.regression_loss(p, t, "nope")


def test_eval_metrics_restatement_matches_reference_functions():
    """oracle/eval_metrics_ref.py against what the unmodified reference's compute_regression_metrics /
    compute_per_lab_metrics (evaluate.py:36-139) returned on the seeded case (tests/golden/eval_metrics.pt)."""
    import math
    import os
    import numpy as np
    import torch
    from oracle import eval_metrics_ref as E
    blob = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.pt"), weights_only=False)
    p, t, lab = blob["pred"].numpy(), blob["target"].numpy(), blob["lab"].numpy()
    p2, t2, lab2 = E.synthetic_case()
    assert np.array_equal(p, p2) and np.array_equal(t, t2) and np.array_equal(lab, lab2)      # the case is reproducible
    pw, n_cap = E.winsorize(p, t, lab)
    assert n_cap == blob["num_capped"] and np.array_equal(pw, blob["pred_winsorized"].numpy())

    def close(a, b):
        return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 2e-6 * max(1.0, abs(b))

    for tag, pp in (("raw", p), ("winsorized", pw)):
        got, want = E.regression_metrics(pp, t), blob[f"overall_{tag}"]
        assert all(close(got[k], want[k]) for k in ("mae", "rmse", "r2", "mape")), (tag, got, want)
        rows, wrows = E.per_lab_metrics(pp, t, lab), blob[f"per_lab_{tag}"]
        assert [r["lab_index"] for r in rows] == [r["lab_index"] for r in wrows]               # same labs, same MAE order
        for r, w in zip(rows, wrows):
            assert r["num_samples"] == w["num_samples"]
            assert all(close(r[k], w[k]) for k in ("mae", "rmse", "r2", "mape")), (tag, r, w)
    assert 48 not in [r["lab_index"] for r in rows] and 49 not in [r["lab_index"] for r in rows]   # 1-sample / empty labs


def test_eval_strata_restatement_matches_reference_golden():
    """oracle/eval_metrics_ref.stratify_by_* (evaluate.py:237-341) against what the unmodified reference functions returned."""
    import os
    import numpy as np
    from oracle import eval_metrics_ref as E
    blob = torch.load(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.pt"), weights_only=False)
    pw, t, lab = blob["pred_winsorized"].numpy(), blob["target"].numpy(), blob["lab"].numpy()
    ei, patient = blob["has_lab_edge_index"].numpy(), blob["patient"].numpy()
    for mine, ref in ((E.stratify_by_patient_degree(pw, t, patient, ei, blob["n_patients"]), blob["by_patient_degree"]),
                      (E.stratify_by_lab_frequency(pw, t, lab, ei, blob["n_labs"]), blob["by_lab_frequency"])):
        assert list(mine) == list(ref)
        for k in ref:
            assert mine[k]["num_samples"] == ref[k]["num_samples"]
            for m in ("mae", "rmse", "r2", "mape"):
                assert abs(mine[k][m] - ref[k][m]) <= 1e-5 * max(1.0, abs(ref[k][m])), (k, m)
    assert set(blob["by_patient_degree"]) == {"low (1-5 labs)", "medium (6-15 labs)", "high (16+ labs)"}

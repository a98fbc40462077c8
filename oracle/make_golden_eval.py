"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Writes tests/golden/eval_metrics.pt: what the UNMODIFIED reference functions
``evaluate.compute_regression_metrics`` / ``evaluate.compute_per_lab_metrics`` (/root/reference/src/evaluate.py:36-139)
return on the seeded case of oracle/eval_metrics_ref.synthetic_case, before and after the (restated) winsorisation.
Run from the repo root in the build container:  python -m oracle.make_golden_eval"""
import importlib
import os

import numpy as np
import torch

from . import eval_metrics_ref as E
from . import ref_harness as H


def main():
    H.load_reference()
    ev = importlib.import_module("evaluate")
    assert ev.__file__.startswith("/root/reference/src"), ev.__file__
    p, t, lab = E.synthetic_case()
    pw, n_cap = E.winsorize(p, t, lab)
    blob = {"pred": torch.from_numpy(p), "target": torch.from_numpy(t), "lab": torch.from_numpy(lab), "n_labs": 50,
            "pred_winsorized": torch.from_numpy(pw), "num_capped": n_cap,
            "overall_raw": ev.compute_regression_metrics(p, t), "overall_winsorized": ev.compute_regression_metrics(pw, t)}
    for tag, pp in (("raw", p), ("winsorized", pw)):
        df = ev.compute_per_lab_metrics(pp, t, lab, {})
        blob[f"per_lab_{tag}"] = [{k: (float(v) if k not in ("lab_index", "num_samples", "lab_name") else v) for k, v in row.items()
                                   if k != "lab_name"} for row in df.to_dict("records")]
    # stratified analysis (evaluate.py:237-341) by the UNMODIFIED reference functions, on the winsorised predictions like
    # evaluate_model does (evaluate.py:521-545); the graph argument is a HeteroData of the PyG shim
    from torch_geometric.data import HeteroData
    patient, ei, n_pat = E.synthetic_strata_case()
    g = HeteroData()
    g["patient"].num_nodes = n_pat
    g["lab"].num_nodes = 50
    g["patient", "has_lab", "lab"].edge_index = torch.from_numpy(ei)
    blob["patient"] = torch.from_numpy(patient)
    blob["has_lab_edge_index"] = torch.from_numpy(ei)
    blob["n_patients"] = n_pat
    to_f = lambda d: {k: {kk: (float(vv) if kk != "num_samples" else int(vv)) for kk, vv in v.items()} for k, v in d.items()}
    blob["by_patient_degree"] = to_f(ev.stratify_by_patient_degree(pw, t, patient, g))
    blob["by_lab_frequency"] = to_f(ev.stratify_by_lab_frequency(pw, t, lab, g))
    # the restatement agrees with the reference functions
    for mine, ref in ((E.stratify_by_patient_degree(pw, t, patient, ei, n_pat), blob["by_patient_degree"]),
                      (E.stratify_by_lab_frequency(pw, t, lab, ei, 50), blob["by_lab_frequency"])):
        assert set(mine) == set(ref), (set(mine), set(ref))
        for k in ref:
            assert mine[k]["num_samples"] == ref[k]["num_samples"]
            assert all(abs(mine[k][m] - ref[k][m]) <= 1e-5 * max(1.0, abs(ref[k][m])) for m in ("mae", "rmse", "r2", "mape")), (k, mine[k], ref[k])
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eval_metrics.pt")
    torch.save(blob, out)
    print("wrote", out, os.path.getsize(out), "bytes; capped", n_cap, "overall", blob["overall_winsorized"])


if __name__ == "__main__":
    main()

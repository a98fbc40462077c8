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
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eval_metrics.pt")
    torch.save(blob, out)
    print("wrote", out, os.path.getsize(out), "bytes; capped", n_cap, "overall", blob["overall_winsorized"])


if __name__ == "__main__":
    main()

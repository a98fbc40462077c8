"""On-device evaluation of imputed lab values: the arithmetic of the reference's ``src/evaluate.py`` right after the hot
path -- per-lab +-3 sigma winsorisation of the residuals (evaluate.py:417-440), then
MAE / RMSE / R^2 / MAPE overall (evaluate.py:36-82) and per lab (evaluate.py:88-139).  One libb2g launch over the pair list
(csrc/metrics.cu); only the per-lab records (a few hundred rows) come back to the host."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .graph import CSR, _stream

FIELDS = ("n", "residual_mean", "residual_std", "n_capped", "sum_abs", "sum_sq", "sum_t", "sum_t2", "sum_ape", "n_nonzero")


def _metrics_from_sums(n, sum_abs, sum_sq, sum_t, sum_t2, sum_ape, n_nz) -> Dict[str, float]:
    """compute_regression_metrics (evaluate.py:36-82) from sufficient statistics."""
    nan = float("nan")
    if n <= 0:
        return {"mae": nan, "rmse": nan, "r2": nan, "mape": nan}
    ss_tot = sum_t2 - sum_t * sum_t / n
    if ss_tot > 0:
        r2 = 1.0 - sum_sq / ss_tot
    else:                                   # sklearn.r2_score: constant targets -> 1.0 if perfect else 0.0; < 2 samples -> nan
        r2 = nan if n < 2 else (1.0 if sum_sq == 0 else 0.0)
    return {"mae": sum_abs / n, "rmse": (sum_sq / n) ** 0.5, "r2": r2, "mape": (sum_ape / n_nz * 100.0) if n_nz > 0 else nan}


DEGREE_GROUPS = ("low (1-5 labs)", "medium (6-15 labs)", "high (16+ labs)")            # evaluate.py:268-272
FREQUENCY_GROUPS = ("rare (bottom 25%)", "common (middle 50%)", "very common (top 25%)")    # evaluate.py:322-326


def evaluate_predictions(predictions: torch.Tensor, targets: torch.Tensor, lab_indices: torch.Tensor, num_labs: int,
                         winsorize: bool = True, n_sigma: float = 3.0, return_winsorized: bool = False,
                         patient_indices: Optional[torch.Tensor] = None, patient_lab_degree: Optional[torch.Tensor] = None,
                         lab_counts: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """predictions / targets float32[M], lab_indices int64[M] (CUDA).  Returns
        overall   {'mae', 'rmse', 'r2', 'mape'}            evaluate.py:445 after the outlier guard
        per_lab   list of {'lab_index', 'num_samples', 'mae', 'rmse', 'r2', 'mape'} for labs with >= 2 samples, sorted by MAE
        num_capped                                         evaluate.py:440
        records   float64 [num_labs, 10] tensor (FIELDS)   raw per-lab sums, on the device
        predictions (optional) the winsorised predictions  evaluate.py:434-437
        stratified  {'by_patient_degree': {...}, 'by_lab_frequency': {...}}  evaluate.py:237-341 (stratify_by_patient_degree /
                    stratify_by_lab_frequency on the winsorised predictions, evaluate.py:521-545), when `patient_indices` +
                    `patient_lab_degree` (int64 bincount of has_lab sources: GraphIndex.patient_lab_degree) resp. `lab_counts`
                    (bincount of has_lab destinations) are given; each group {'mae','rmse','r2','mape','num_samples'}
    """
    if not (predictions.is_cuda and targets.is_cuda and lab_indices.is_cuda):
        raise _lib.B2GError("evaluate_predictions needs CUDA tensors (there is no CPU path)")
    lib = _lib.load()
    pred = predictions.detach().contiguous().float()
    tgt = targets.detach().contiguous().float()
    lab = lab_indices.contiguous().long()
    if not (pred.numel() == tgt.numel() == lab.numel()):
        raise ValueError("predictions, targets and lab_indices must have the same length")
    rec = torch.zeros((int(num_labs), len(FIELDS)), dtype=torch.float64, device=pred.device)
    pw: Optional[torch.Tensor] = torch.empty_like(pred) if return_winsorized else None
    want_degree = patient_indices is not None and patient_lab_degree is not None
    bins = torch.zeros((int(num_labs), 3, 7), dtype=torch.float64, device=pred.device) if want_degree else None
    if pred.numel() > 0:
        by_lab = CSR(lab, lab, int(num_labs), int(num_labs), col_is_eid=True)      # pairs of each lab, original order
        if want_degree:
            pidx = patient_indices.contiguous().long()
            deg = patient_lab_degree.contiguous().long()
            if pidx.numel() != pred.numel():
                raise ValueError("patient_indices must have one entry per prediction")
            _lib.check(lib.b2g_eval_per_lab_strata(pred.data_ptr(), tgt.data_ptr(), by_lab.rowptr.data_ptr(), by_lab.col.data_ptr(),
                                                   pidx.data_ptr(), deg.data_ptr(), int(num_labs), int(bool(winsorize)), float(n_sigma),
                                                   rec.data_ptr(), bins.data_ptr(), None if pw is None else pw.data_ptr(), _stream()),
                       "b2g_eval_per_lab_strata")
        else:
            _lib.check(lib.b2g_eval_per_lab(pred.data_ptr(), tgt.data_ptr(), by_lab.rowptr.data_ptr(), by_lab.col.data_ptr(), int(num_labs),
                                            int(bool(winsorize)), float(n_sigma), rec.data_ptr(), None if pw is None else pw.data_ptr(), _stream()),
                       "b2g_eval_per_lab")
    host = rec.cpu()
    tot = host.sum(0).tolist()
    out: Dict[str, object] = {"overall": _metrics_from_sums(tot[0], tot[4], tot[5], tot[6], tot[7], tot[8], tot[9]),
                              "num_capped": int(tot[3]), "records": rec}
    rows = []
    for i, r in enumerate(host.tolist()):
        if r[0] < 2:                        # evaluate.py:122 "Need at least 2 samples"
            continue
        m = _metrics_from_sums(r[0], r[4], r[5], r[6], r[7], r[8], r[9])
        m.update(lab_index=i, num_samples=int(r[0]))
        rows.append(m)
    rows.sort(key=lambda m: m["mae"])       # evaluate.py:137
    out["per_lab"] = rows
    if pw is not None:
        out["predictions"] = pw
    strat: Dict[str, Dict] = {}
    if want_degree:                          # evaluate.py:237-287
        tot_b = bins.sum(0).cpu().tolist()
        groups = {}
        for name, b in zip(DEGREE_GROUPS, tot_b):
            if b[0] > 0:
                m = _metrics_from_sums(b[0], b[1], b[2], b[3], b[4], b[5], b[6])
                m["num_samples"] = int(b[0])
                groups[name] = m
        strat["by_patient_degree"] = groups
    if lab_counts is not None:               # evaluate.py:290-341: quartiles of the per-lab edge counts (labs with > 0 edges)
        import numpy as np
        counts = lab_counts.detach().cpu().numpy().astype(np.int64)
        q25, q75 = np.percentile(counts[counts > 0], 25), np.percentile(counts[counts > 0], 75)
        masks = (counts < q25, (counts >= q25) & (counts <= q75), counts > q75)
        h = host.numpy()
        groups = {}
        for name, mk in zip(FREQUENCY_GROUPS, masks):
            r = h[mk].sum(0)
            if r[0] > 0:
                m = _metrics_from_sums(r[0], r[4], r[5], r[6], r[7], r[8], r[9])
                m["num_samples"] = int(r[0])
                groups[name] = m
        strat["by_lab_frequency"] = groups
    if strat:
        out["stratified"] = strat
    return out

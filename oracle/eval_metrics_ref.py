"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement, in numpy, of the evaluation arithmetic that follows the hot path in the reference's src/evaluate.py:
  * winsorize()            evaluate.py:417-440  per-lab +-3 sigma cap of the residuals, written back into the predictions
  * regression_metrics()   evaluate.py:36-82    MAE / RMSE / R^2 / MAPE (sklearn semantics restated without sklearn)
  * per_lab_metrics()      evaluate.py:88-139   the same per lab, labs with < 2 samples skipped, sorted by MAE
Pinned: regression_metrics / per_lab_metrics against the reference's own functions (imported unmodified where
/root/reference exists: tests/test_oracle_vs_reference.py) and against tests/golden/eval_metrics.pt, which
oracle/make_golden_eval.py produced by calling those reference functions.  The winsorisation is an inline block of
evaluate_model() in the reference, so it is restated only (same numpy calls, same float32 arithmetic).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def winsorize(predictions: np.ndarray, targets: np.ndarray, lab_indices: np.ndarray, n_sigma: float = 3.0):
    """evaluate.py:417-440.  Returns (winsorised predictions, number of capped residuals)."""
    predictions = predictions.copy()
    residuals = predictions - targets
    num_capped = 0
    for lab in np.unique(lab_indices):
        mask = lab_indices == lab
        r = residuals[mask]
        if len(r) > 1:
            sd, mu = np.std(r), np.mean(r)
            capped = np.clip(r, mu - n_sigma * sd, mu + n_sigma * sd)
            num_capped += int(np.sum(r != capped))
            predictions[mask] = targets[mask] + capped
    return predictions, num_capped


def regression_metrics(predictions: np.ndarray, targets: np.ndarray) -> Dict[str, float]:
    """evaluate.py:36-82 (sklearn.metrics.mean_absolute_error / mean_squared_error / r2_score restated)."""
    p, t = predictions.astype(np.float64), targets.astype(np.float64)
    mae = float(np.mean(np.abs(t - p)))
    mse = float(np.mean((t - p) ** 2))
    ss_res, ss_tot = float(np.sum((t - p) ** 2)), float(np.sum((t - np.mean(t)) ** 2))
    if len(t) < 2:
        r2 = float("nan")
    elif ss_tot == 0.0:
        r2 = 1.0 if ss_res == 0.0 else 0.0
    else:
        r2 = 1.0 - ss_res / ss_tot
    nz = targets != 0
    mape = float(np.mean(np.abs((t[nz] - p[nz]) / t[nz])) * 100) if nz.sum() > 0 else float("nan")
    return {"mae": mae, "rmse": float(np.sqrt(mse)), "r2": r2, "mape": mape}


def per_lab_metrics(predictions: np.ndarray, targets: np.ndarray, lab_indices: np.ndarray) -> List[Dict[str, float]]:
    """evaluate.py:88-139: rows sorted by MAE, labs with fewer than 2 samples skipped."""
    rows = []
    for lab in np.unique(lab_indices):
        mask = lab_indices == lab
        if mask.sum() < 2:
            continue
        m = regression_metrics(predictions[mask], targets[mask])
        m["lab_index"] = int(lab)
        m["num_samples"] = int(mask.sum())
        rows.append(m)
    rows.sort(key=lambda m: m["mae"])
    return rows


def synthetic_case(seed: int = 7, n_pairs: int = 20000, n_labs: int = 50):
    """Seeded (predictions, targets, lab_indices) with what the edge cases need: heavy-tailed residuals (so the cap
    triggers), a lab with a single sample, a lab that never occurs, exact-zero targets (MAPE mask), a constant-target lab."""
    rng = np.random.RandomState(seed)
    lab = rng.randint(0, n_labs - 2, size=n_pairs).astype(np.int64)        # lab n_labs-1 never occurs
    lab[0] = n_labs - 2                                                      # ... and lab n_labs-2 exactly once
    t = np.clip(rng.randn(n_pairs), -5, 5).astype(np.float32)
    t[rng.rand(n_pairs) < 0.01] = 0.0
    t[lab == 3] = 0.5                                                        # constant targets
    p = (0.6 * t + 0.5 * rng.standard_t(3, size=n_pairs)).astype(np.float32)
    return p, t, lab


def synthetic_strata_case(seed: int = 11, n_pairs: int = 20000, n_labs: int = 50, n_patients: int = 700):
    """Patient index per pair of synthetic_case + a has_lab edge list [2, E] (patient, lab) whose degrees hit every stratum of
    evaluate.py:268-272 (1-5 / 6-15 / 16+ labs per patient, plus patients with no lab at all) and whose lab counts span the
    quartile groups of evaluate.py:315-326."""
    rng = np.random.RandomState(seed)
    patient = rng.randint(0, n_patients, size=n_pairs).astype(np.int64)
    deg = np.concatenate([rng.randint(1, 6, n_patients // 3), rng.randint(6, 16, n_patients // 3),
                          rng.randint(16, n_labs - 1, n_patients - 2 * (n_patients // 3))])
    deg[:5] = 0
    weights = np.linspace(1.0, 8.0, n_labs - 1)
    weights /= weights.sum()
    src, dst = [], []
    for p_, d_ in enumerate(deg):
        if d_ > 0:
            labs = rng.choice(n_labs - 1, size=int(d_), replace=False, p=weights)
            src += [p_] * int(d_)
            dst += labs.tolist()
    return patient, np.stack([np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)]), n_patients


def stratify_by_patient_degree(predictions, targets, patient_indices, has_lab_edge_index, n_patients):
    """evaluate.py:237-287."""
    degrees = np.bincount(has_lab_edge_index[0], minlength=n_patients)
    d = degrees[patient_indices]
    groups = {"low (1-5 labs)": (d >= 1) & (d <= 5), "medium (6-15 labs)": (d >= 6) & (d <= 15), "high (16+ labs)": d >= 16}
    out = {}
    for name, mask in groups.items():
        if mask.sum() > 0:
            m = regression_metrics(predictions[mask], targets[mask])
            m["num_samples"] = int(mask.sum())
            out[name] = m
    return out


def stratify_by_lab_frequency(predictions, targets, lab_indices, has_lab_edge_index, n_labs):
    """evaluate.py:290-341."""
    counts = np.bincount(has_lab_edge_index[1], minlength=n_labs)
    f = counts[lab_indices]
    q25, q75 = np.percentile(counts[counts > 0], 25), np.percentile(counts[counts > 0], 75)
    groups = {"rare (bottom 25%)": f < q25, "common (middle 50%)": (f >= q25) & (f <= q75), "very common (top 25%)": f > q75}
    out = {}
    for name, mask in groups.items():
        if mask.sum() > 0:
            m = regression_metrics(predictions[mask], targets[mask])
            m["num_samples"] = int(mask.sum())
            out[name] = m
    return out

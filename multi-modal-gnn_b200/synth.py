"""Seeded synthetic EHR-shaped heterogeneous graphs (BASELINE.json configs).

Real eICU / MIMIC-III data is credentialed and absent; every test and benchmark uses graphs drawn
here with the *shape* the reference's graph builder emits (/root/reference/src/graph_build.py:148-261):

  node types  patient, lab, diagnosis, medication               (graph_build.py:186-201)
  edge types  has_lab, has_lab_rev, has_diagnosis, has_diagnosis_rev, has_medication,
              has_medication_rev, in that order                  (graph_build.py:216-247)
  * every (patient, x) pair is unique                            (preprocess.py:85-100,242,385)
  * has_lab arrives grouped by lab, then patient                 (preprocess.py:141-147)
  * reverse relations are edge_index.flip(0)                     (graph_build.py:222,235,247)
  * edge_attr [E,1] float32 (z-scored lab value) on has_lab and has_lab_rev only (graph_build.py:217,223)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .heterodata import HeteroGraph

NODE_TYPES = ["patient", "lab", "diagnosis", "medication"]
EDGE_TYPES = [
    ("patient", "has_lab", "lab"),
    ("lab", "has_lab_rev", "patient"),
    ("patient", "has_diagnosis", "diagnosis"),
    ("diagnosis", "has_diagnosis_rev", "patient"),
    ("patient", "has_medication", "medication"),
    ("medication", "has_medication_rev", "patient"),
]


@dataclass(frozen=True)
class GraphSpec:
    name: str
    n_patient: int
    n_lab: int
    n_dx: int
    n_med: int
    e_lab: int
    e_dx: int
    e_med: int
    low_degree_frac: float = 0.0   # fraction of patients forced to 1..5 labs (exercises the gate)
    hidden_dim: int = 128

    @property
    def directed_edges_per_layer(self) -> int:
        """Unit of BASELINE.json's metric: 2*(E_l+E_d+E_m)."""
        return 2 * (self.e_lab + self.e_dx + self.e_med)


# BASELINE.json configs (sizes from SURVEY.md section 8d).
SPECS = {
    "tiny": GraphSpec("tiny", 300, 20, 30, 25, 4200, 700, 1500, low_degree_frac=0.15),
    # more diagnosis / medication nodes than a rank's patient share (multi-GPU corner case: a replicated table that is larger
    # than the rank-local destination set; tools/dist_check.py)
    "wide": GraphSpec("wide", 160, 20, 114, 100, 1600, 480, 1400, low_degree_frac=0.1),
    "C1": GraphSpec("C1", 1834, 50, 114, 100, 61484, 5421, 15933, low_degree_frac=0.01),
    "C2": GraphSpec("C2", 46520, 160, 200, 100, 5_000_000, 441_000, 1_296_000, low_degree_frac=0.01),
    "C3": GraphSpec("C3", 1_000_000, 50, 200, 100, 20_000_000, 1_764_000, 5_182_000, low_degree_frac=0.12),
    "C4": GraphSpec("C4", 10_000_000, 50, 200, 100, 100_000_000, 8_820_000, 25_910_000, low_degree_frac=0.12),
    # one rank's share of C4 when its 10 M patients are partitioned over 8 GPUs (bench.py --workload C4s8 --gpus 8 == C4)
    "C4s8": GraphSpec("C4s8", 1_250_000, 50, 200, 100, 12_500_000, 1_102_500, 3_238_750, low_degree_frac=0.12),
    "C5": GraphSpec("C5", 10_000_000, 50, 200, 100, 100_000_000, 8_820_000, 25_910_000,
                    low_degree_frac=0.12, hidden_dim=256),
}


def _target_degrees(n_rows, n_cols, total, lo, gen, low_frac=0.0):
    """Integer degree per row in [lo, n_cols] summing exactly to ``total``."""
    dev = gen.device
    total = int(total)
    assert lo * n_rows <= total <= n_rows * n_cols, "edge count infeasible for unique pairs"
    deg = torch.empty(n_rows, dtype=torch.int64, device=dev)
    n_low = int(round(low_frac * n_rows)) if n_cols > 6 else 0
    is_low = torch.zeros(n_rows, dtype=torch.bool, device=dev)
    if n_low > 0:
        is_low[torch.randperm(n_rows, generator=gen, device=dev)[:n_low]] = True
        deg[is_low] = torch.randint(1, 6, (n_low,), generator=gen, device=dev)
    n_hi = n_rows - n_low
    rest = total - int(deg[is_low].sum()) if n_low else total
    mean_hi = rest / max(n_hi, 1)
    hi_lo = max(lo, 6 if n_low else lo)
    # lab degrees cluster around the mean; diagnosis / medication degrees (lo == 0) spread over [0, 2*mean] so that
    # some patients have none at all (87 of 1,834 eICU patients have no diagnosis edge)
    spread = max(1.0, 0.25 * mean_hi) if lo > 0 else max(1.0, mean_hi + 0.5)
    draw = mean_hi + spread * (2 * torch.rand(n_hi, generator=gen, device=dev) - 1)
    deg[~is_low] = draw.round().long().clamp(hi_lo, n_cols)
    # fix the sum exactly by +-1 nudges on rows with head-room
    diff = total - int(deg.sum())
    while diff != 0:
        step = 1 if diff > 0 else -1
        ok = (~is_low) & ((deg < n_cols) if step > 0 else (deg > hi_lo))
        cand = ok.nonzero().squeeze(1)
        assert cand.numel() > 0, "cannot reach requested edge count"
        take = cand[torch.randperm(cand.numel(), generator=gen, device=dev)[: abs(diff)]]
        deg[take] += step
        diff = total - int(deg.sum())
    return deg


def _sample_pairs(n_rows, n_cols, deg, gen, chunk_cells=1 << 25):
    """Unique (row, col) pairs: row r gets deg[r] distinct columns, chosen by Gumbel top-k with a
    per-column popularity so column in-degrees are uneven (eICU labs: 350..1850 of 1834 patients)."""
    dev = gen.device
    pop = torch.rand(n_cols, generator=gen, device=dev) * 2.0  # log-popularity in [0, 2)
    rows_out, cols_out = [], []
    rows_per_chunk = max(1, chunk_cells // n_cols)
    for r0 in range(0, n_rows, rows_per_chunk):
        r1 = min(n_rows, r0 + rows_per_chunk)
        u = torch.rand(r1 - r0, n_cols, generator=gen, device=dev).clamp_(1e-12, 1 - 1e-7)
        key = pop.unsqueeze(0) - torch.log(-torch.log(u))
        order = key.argsort(dim=1, descending=True)
        rank = torch.empty_like(order)
        rank.scatter_(1, order, torch.arange(n_cols, device=dev).expand_as(order))
        pick = rank < deg[r0:r1].unsqueeze(1)
        rr, cc = pick.nonzero(as_tuple=True)
        rows_out.append(rr + r0)
        cols_out.append(cc)
    return torch.cat(rows_out), torch.cat(cols_out)


def make_graph(spec: GraphSpec | str, seed: int = 42, device: str | torch.device = "cpu") -> HeteroGraph:
    """Draw a graph of the given shape.  ``device`` is where generation runs (CPU for tests; a CUDA
    device for the 100M-edge configs); the result lives on that device."""
    if isinstance(spec, str):
        spec = SPECS[spec]
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    g = HeteroGraph()
    g["patient"].num_nodes = spec.n_patient
    g["lab"].num_nodes = spec.n_lab
    g["diagnosis"].num_nodes = spec.n_dx
    g["medication"].num_nodes = spec.n_med

    # patient-lab: every patient has >= 1 lab
    deg_l = _target_degrees(spec.n_patient, spec.n_lab, spec.e_lab, 1, gen, spec.low_degree_frac)
    p, l = _sample_pairs(spec.n_patient, spec.n_lab, deg_l, gen)
    order = (l * spec.n_patient + p).argsort()           # grouped by lab, then patient
    p, l = p[order], l[order]
    zp = torch.randn(spec.n_patient, 8, generator=gen, device=dev)
    wl = torch.randn(spec.n_lab, 8, generator=gen, device=dev)
    eps = torch.randn(p.numel(), generator=gen, device=dev)
    val = 0.6 * (zp[p] * wl[l]).sum(1) / math.sqrt(8.0) + 0.8 * eps
    cnt = torch.zeros(spec.n_lab, device=dev).index_add_(0, l, torch.ones_like(val)).clamp_(min=1)
    mean = torch.zeros(spec.n_lab, device=dev).index_add_(0, l, val) / cnt
    var = torch.zeros(spec.n_lab, device=dev).index_add_(0, l, (val - mean[l]) ** 2) / cnt
    val = ((val - mean[l]) / var[l].sqrt().clamp_(min=1e-6)).clamp_(-5.0, 5.0).float()
    ei = torch.stack([p, l]).contiguous()
    g["patient", "has_lab", "lab"].edge_index = ei
    g["patient", "has_lab", "lab"].edge_attr = val.unsqueeze(1).contiguous()
    g["lab", "has_lab_rev", "patient"].edge_index = ei.flip(0).contiguous()
    g["lab", "has_lab_rev", "patient"].edge_attr = val.unsqueeze(1).contiguous()

    for rel, n_t, e_t, dst in (("has_diagnosis", spec.n_dx, spec.e_dx, "diagnosis"),
                               ("has_medication", spec.n_med, spec.e_med, "medication")):
        deg = _target_degrees(spec.n_patient, n_t, e_t, 0, gen)
        p, t = _sample_pairs(spec.n_patient, n_t, deg, gen)
        ei = torch.stack([p, t]).contiguous()               # patient-major order (drop_duplicates order)
        g["patient", rel, dst].edge_index = ei
        g[dst, rel + "_rev", "patient"].edge_index = ei.flip(0).contiguous()
    return g

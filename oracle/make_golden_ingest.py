"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Writes tests/golden/ingest_small.pt: seeded synthetic preprocessing tables and what the
UNMODIFIED reference graph builder (/root/reference/src/graph_build.py::build_heterogeneous_graph, :104-273; NodeIndexer :34-97;
edge creators :476-586) makes of them: node counts, node / edge type order, every edge_index / edge_attr, data.indexers.
Run from the repo root in the build container:  python -m oracle.make_golden_ingest"""
import importlib
import os

import numpy as np
import pandas as pd
import torch

from . import ref_harness as H


def synthetic_tables(seed=5, n_patients=400, n_labs=37, n_dx=60, n_drugs=45):
    rng = np.random.RandomState(seed)
    subject = rng.permutation(np.arange(10_000, 10_000 + 7 * n_patients, 7))[:n_patients].astype(np.int64)     # non-contiguous ids
    cohort = pd.DataFrame({"SUBJECT_ID": subject})
    itemids = rng.permutation(np.arange(50_800, 50_800 + 3 * n_labs, 3))[:n_labs]
    rows = []
    for iid in itemids:                                    # grouped by lab (preprocess.py:141-147), patients in random order
        pats = rng.choice(subject, size=rng.randint(5, n_patients // 2), replace=False)
        for p in pats:
            rows.append((float(p), int(iid), float(np.clip(rng.randn(), -5, 5))))          # SUBJECT_ID as float64 (pandas does that)
    rows += [(99_999_999.0, int(itemids[0]), 0.25), (float(subject[0]), int(itemids[1]), -0.5)]   # unknown patient; a duplicate pair
    labs = pd.DataFrame(rows, columns=["SUBJECT_ID", "ITEMID", "VALUE_NORMALIZED"])
    codes = [f"{c:03d}" for c in rng.choice(900, n_dx - 6, replace=False)] + ["V58", "V10", "E88", "E93", "V45", "250"]
    dx = pd.DataFrame({"SUBJECT_ID": rng.choice(np.append(subject, 123), size=1500), "ICD3_CODE": rng.choice(codes, size=1500)}).drop_duplicates()
    drugs = [f"drug {i} hcl" for i in range(n_drugs)]
    med = pd.DataFrame({"SUBJECT_ID": rng.choice(subject, size=2500), "DRUG": rng.choice(drugs, size=2500)}).drop_duplicates()
    labitems = pd.DataFrame({"ITEMID": itemids, "LABEL": [f"lab {i}" for i in itemids], "FLUID": "Blood", "CATEGORY": "Chemistry"})
    demographics = pd.DataFrame({"SUBJECT_ID": subject, "AGE": rng.randint(18, 90, n_patients)})
    return cohort, labs, dx.reset_index(drop=True), med.reset_index(drop=True), demographics, labitems


CONFIG = {"graph": {"edge_types": {"patient_lab": {"enabled": True, "bidirectional": True},
                                    "patient_diagnosis": {"enabled": True, "bidirectional": True},
                                    "patient_medication": {"enabled": True, "bidirectional": True}}}}


def main():
    H.load_reference()
    gb = importlib.import_module("graph_build")
    assert gb.__file__.startswith("/root/reference/src"), gb.__file__
    cohort, labs, dx, med, demo, labitems = synthetic_tables()
    data = gb.build_heterogeneous_graph(cohort, labs, dx, med, demo, labitems, CONFIG)
    blob = {"tables": {"cohort": {"SUBJECT_ID": cohort["SUBJECT_ID"].to_numpy()},
                       "labs": {c: labs[c].to_numpy() for c in labs.columns},
                       "diagnoses": {"SUBJECT_ID": dx["SUBJECT_ID"].to_numpy(), "ICD3_CODE": dx["ICD3_CODE"].to_numpy().astype(str)},
                       "medications": {"SUBJECT_ID": med["SUBJECT_ID"].to_numpy(), "DRUG": med["DRUG"].to_numpy().astype(str)}},
            "node_types": list(data.node_types), "edge_types": [tuple(e) for e in data.edge_types],
            "num_nodes": {nt: int(data[nt].num_nodes) for nt in data.node_types},
            "edge_index": {"__".join(et): data[et].edge_index.clone() for et in data.edge_types},
            "edge_attr": {"__".join(et): data[et].edge_attr.clone() for et in data.edge_types if "edge_attr" in data[et]},
            "indexers": data.indexers}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ingest_small.pt")
    torch.save(blob, out)
    print("wrote", out, os.path.getsize(out), "bytes;", blob["num_nodes"], {k: tuple(v.shape) for k, v in blob["edge_index"].items()})


if __name__ == "__main__":
    main()

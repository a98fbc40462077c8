"""Graph ingest without per-row Python loops: the tables the reference's preprocessing writes
(`preprocess.py` parquet files: cohort, labs, diagnoses, medications) -> the same heterogeneous graph
`/root/reference/src/graph_build.py::build_heterogeneous_graph` produces, with the per-row work on the GPU.

What the reference does (graph_build.py:156-173, 476-586) and what is kept bit-for-bit:
  * NodeIndexer: entity id -> contiguous index in FIRST-OCCURRENCE order over `cohort.SUBJECT_ID`, `labs.ITEMID.unique()`,
    `diagnoses.ICD3_CODE.unique()`, `medications.DRUG.unique()`; numeric ids are compared as `str(int(id))`;
  * one edge per table row, in row order, rows whose patient (or lab / code / drug) is unknown dropped;
    `edge_index [2, E] int64`, `edge_attr [E, 1] float32` (VALUE_NORMALIZED) on has_lab and has_lab_rev;
  * reverse relations are `edge_index.flip(0)`; node / edge type insertion order of graph_build.py:186-247;
  * `data.indexers[node_type] = {'id_to_index': {str: int}, 'index_to_id': {int: str}}` (graph_build.py:254-260), which
    `inference.py:330-331` reads.

Division of labour: vocabulary construction (a few hundred to a few million DISTINCT ids, first-occurrence order) is a
vectorised host pass (`pandas.unique`, hash based, order preserving); the per-row look-ups (binary search in the sorted
dictionary) and the stable filtering of the rows into COO edges run in libb2g kernels (csrc/ingest.cu); the CSR / bit-matrix
structures the model needs are then built on the device as for any other graph (graph.py).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .graph import _stream, workspace
from .heterodata import HeteroGraph


def _column(table, name):
    """a column of a pandas DataFrame / dict of arrays as a numpy array"""
    col = table[name]
    return col.to_numpy() if hasattr(col, "to_numpy") else np.asarray(col)


def _canonical_ids(values: np.ndarray):
    """NodeIndexer's key normalisation (graph_build.py:63-69): numeric -> int -> str, anything else -> str.  Returns
    (int64 array or None, object array of strings or None): integer-valued columns stay numeric (device path)."""
    if values.dtype.kind in "iu":
        return values.astype(np.int64), None
    if values.dtype.kind == "f":
        return values.astype(np.int64), None              # int(10006.0) == 10006, like NodeIndexer.add
    return None, np.asarray([str(int(v)) if isinstance(v, (int, float, np.integer, np.floating)) else str(v) for v in values], dtype=object)


class Vocabulary:
    """One NodeIndexer: distinct ids in first-occurrence order."""

    def __init__(self, values: np.ndarray):
        import pandas as pd
        ints, strs = _canonical_ids(np.asarray(values))
        self.numeric = ints is not None
        self.ids = pd.unique(ints if self.numeric else strs)       # order of appearance (pandas.unique keeps it)
        self._sorted = None
        self._dicts = None

    def __len__(self):
        return int(len(self.ids))

    def device_dictionary(self, device):
        """(sorted ids int64, node index of each int32) on the device -- numeric vocabularies only"""
        if self._sorted is None:
            order = np.argsort(self.ids, kind="stable")
            self._sorted = (torch.from_numpy(self.ids[order].astype(np.int64)).to(device), torch.from_numpy(order.astype(np.int32)).to(device))
        return self._sorted

    def codes(self, values: np.ndarray, device) -> torch.Tensor:
        """int32[rows] node index of every row's id, -1 when the id is not in the vocabulary (NodeIndexer.get_index)."""
        lib = _lib.load()
        ints, strs = _canonical_ids(np.asarray(values))
        if self.numeric and ints is not None:
            q = torch.from_numpy(ints).to(device)
            out = torch.empty(q.numel(), dtype=torch.int32, device=device)
            ids, idx = self.device_dictionary(device)
            _lib.check(lib.b2g_id_lookup(ids.data_ptr(), idx.data_ptr(), ids.numel(), q.data_ptr(), q.numel(), out.data_ptr(), _stream()),
                       "b2g_id_lookup")
            return out
        # string keys (ICD-9 groups, drug names): hash look-up on the host, vectorised (pandas Index.get_indexer)
        import pandas as pd
        if strs is None:
            strs = np.asarray([str(v) for v in ints], dtype=object)
        keys = self.ids if not self.numeric else np.asarray([str(v) for v in self.ids], dtype=object)
        return torch.from_numpy(pd.Index(keys).get_indexer(strs).astype(np.int32)).to(device)

    def as_dicts(self) -> Dict[str, dict]:
        if self._dicts is None:
            keys = [str(v) for v in self.ids.tolist()]
            self._dicts = {"id_to_index": dict(zip(keys, range(len(keys)))), "index_to_id": dict(enumerate(keys))}
        return self._dicts


class _LazyIndexers(dict):
    """data.indexers (graph_build.py:254-260): the string dictionaries are only materialised when somebody reads them
    (10 M patients = 20 M Python objects)."""

    def __init__(self, vocabs: Dict[str, Vocabulary]):
        super().__init__()
        self._vocabs = vocabs
        for k in vocabs:
            dict.__setitem__(self, k, None)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if v is None:
            v = self._vocabs[key].as_dicts()
            dict.__setitem__(self, key, v)
        return v

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


def edges_from_rows(src_idx: torch.Tensor, dst_idx: torch.Tensor, attr: Optional[torch.Tensor] = None):
    """Stable device filter of table rows into COO edges: (edge_index int64 [2, E], edge_attr float32 [E, 1] or None)."""
    lib = _lib.load()
    m = int(src_idx.numel())
    dev = src_idx.device
    if m == 0:
        return (torch.empty((2, 0), dtype=torch.int64, device=dev),
                None if attr is None else torch.empty((0, 1), dtype=torch.float32, device=dev))
    buf = torch.empty(2 * max(m, 1), dtype=torch.int64, device=dev)
    attr_out = torch.empty(max(m, 1), dtype=torch.float32, device=dev) if attr is not None else None
    n_edges = ctypes.c_int64(0)
    ws = workspace(lib.b2g_edges_from_rows_ws_bytes(m), dev)
    _lib.check(lib.b2g_edges_from_rows(src_idx.data_ptr(), dst_idx.data_ptr(), None if attr is None else attr.data_ptr(), m, buf.data_ptr(),
                                       None if attr_out is None else attr_out.data_ptr(), None, ctypes.byref(n_edges), ws.data_ptr(),
                                       ws.numel(), _stream()), "b2g_edges_from_rows")
    e = int(n_edges.value)
    edge_index = buf[:2 * e].view(2, e).clone() if e else torch.empty((2, 0), dtype=torch.int64, device=dev)
    edge_attr = None if attr is None else (attr_out[:e].clone().unsqueeze(1) if e else torch.empty((0, 1), dtype=torch.float32, device=dev))
    return edge_index, edge_attr


def build_graph_from_tables(cohort, labs, diagnoses, medications, device="cuda", bidirectional: bool = True) -> HeteroGraph:
    """graph_build.build_heterogeneous_graph (graph_build.py:104-273) for the shipped configuration (all three edge families
    enabled, bidirectional).  `cohort` needs SUBJECT_ID; `labs` SUBJECT_ID, ITEMID, VALUE_NORMALIZED; `diagnoses` SUBJECT_ID,
    ICD3_CODE; `medications` SUBJECT_ID, DRUG (pandas DataFrames or dicts of arrays).  Returns a HeteroGraph on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.B2GError("build_graph_from_tables runs its per-row work in CUDA kernels: pass a CUDA device (there is no CPU path)")
    vocabs = {"patient": Vocabulary(_column(cohort, "SUBJECT_ID")), "lab": Vocabulary(_column(labs, "ITEMID")),
              "diagnosis": Vocabulary(_column(diagnoses, "ICD3_CODE")), "medication": Vocabulary(_column(medications, "DRUG"))}
    g = HeteroGraph()
    for nt in ("patient", "lab", "diagnosis", "medication"):              # graph_build.py:186-201
        g[nt].num_nodes = len(vocabs[nt])
    specs = (("has_lab", "lab", labs, "ITEMID", "VALUE_NORMALIZED"), ("has_diagnosis", "diagnosis", diagnoses, "ICD3_CODE", None),
             ("has_medication", "medication", medications, "DRUG", None))
    for rel, nt, table, col, attr_col in specs:                           # graph_build.py:210-247
        p_idx = vocabs["patient"].codes(_column(table, "SUBJECT_ID"), device)
        t_idx = vocabs[nt].codes(_column(table, col), device)
        attr = torch.from_numpy(_column(table, attr_col).astype(np.float32)).to(device) if attr_col else None
        ei, ea = edges_from_rows(p_idx, t_idx, attr)
        g["patient", rel, nt].edge_index = ei
        if ea is not None:
            g["patient", rel, nt].edge_attr = ea
        if bidirectional:
            g[nt, rel + "_rev", "patient"].edge_index = ei.flip(0).contiguous()
            if ea is not None:
                g[nt, rel + "_rev", "patient"].edge_attr = ea
    g.indexers = _LazyIndexers(vocabs)
    return g

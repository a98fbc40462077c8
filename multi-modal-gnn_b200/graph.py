"""Device-resident graph structure: per-relation CSR in both directions, degrees, the gate's degree
vector, and CSR views of (patient, lab) prediction pairs.

Input contract = what /root/reference/src/graph_build.py:148-261 emits: ``edge_index [2,E] int64`` COO per
edge type.  The structure is built once per graph object (the graph is static for a whole run, SURVEY.md
note N5) by libb2g's radix-sort CSR builder and cached.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib

EdgeType = Tuple[str, str, str]


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.B2GError(f"{what} must live on a CUDA device (got {t.device}); this package has no CPU path")


def _stream():
    return torch.cuda.current_stream().cuda_stream


_WS: Dict[int, torch.Tensor] = {}
_WS_RETIRED: Dict[int, list] = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Grow-only scratch buffer per device (all library calls are stream-ordered on the current stream).  A buffer that is
    outgrown is RETIRED, not freed: a captured CUDA graph may have its address baked in (Trainer._capture), and replaying it
    after the allocator handed that memory to another tensor would corrupt it.  Growth is geometric, so the retired buffers
    sum to less than the live one."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    buf = _WS.get(idx)
    if buf is None or buf.numel() < nbytes:
        size = max(int(nbytes), 1 << 20, 2 * buf.numel() if buf is not None else 0)
        if buf is not None:
            _WS_RETIRED.setdefault(idx, []).append(buf)
        buf = torch.empty(size, dtype=torch.uint8, device=torch.device("cuda", idx))
        _WS[idx] = buf
    return buf


class CSR:
    """rows -> sorted (stable) list of neighbour ids.  ``long_rows`` switches the reducer to the chunked
    two-phase kernel (type-destination rows own up to millions of neighbours)."""

    def __init__(self, key: torch.Tensor, val: torch.Tensor, n_rows: int, n_vals: int, col_is_eid: bool = False):
        _require_cuda(key, "edge_index")
        lib = _lib.load()
        dev = key.device
        key = key.contiguous()
        val = val.contiguous()
        if key.dtype != torch.int64 or val.dtype != torch.int64:
            key, val = key.long(), val.long()
        e = int(key.numel())
        self.n_rows, self.n_vals, self.n_edges = int(n_rows), int(n_vals), e
        self.rowptr = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
        self.col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        self.eid = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        ws_bytes = lib.b2g_csr_build_ws_bytes(e, self.n_rows)
        ws = workspace(ws_bytes, dev)
        _lib.check(lib.b2g_csr_build(key.data_ptr(), val.data_ptr(), e, self.n_rows, self.n_vals, self.rowptr.data_ptr(),
                                     self.col.data_ptr(), self.eid.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "b2g_csr_build")
        if col_is_eid:
            self.col = self.eid
        self.deg = torch.empty(self.n_rows, dtype=torch.int64, device=dev)
        self.inv_deg = torch.empty(self.n_rows, dtype=torch.float32, device=dev)
        _lib.check(lib.b2g_csr_degrees(self.rowptr.data_ptr(), self.n_rows, self.deg.data_ptr(), self.inv_deg.data_ptr(),
                                       _stream()), "b2g_csr_degrees")
        # long-row decomposition
        max_deg = int(self.deg.max().item()) if self.n_rows > 0 and e > 0 else 0
        self.max_deg = max_deg
        self.long_rows = max_deg > 1024
        self.chunk = 0
        self.n_items = 0
        self.item_row = self.item_start = self.row_item_ptr = None
        if self.long_rows:
            chunk = 32
            while chunk < 512 and e // (chunk * 2) >= 16384:
                chunk *= 2
            self.chunk = chunk
            self.row_item_ptr = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
            n_items = ctypes.c_int64(0)
            wsb = lib.b2g_csr_chunk_ws_bytes(self.n_rows)
            ws = workspace(wsb, dev)
            _lib.check(lib.b2g_csr_chunk_count(self.rowptr.data_ptr(), self.n_rows, chunk, self.row_item_ptr.data_ptr(),
                                               ctypes.byref(n_items), ws.data_ptr(), ws.numel(), _stream()),
                       "b2g_csr_chunk_count")
            self.n_items = int(n_items.value)
            self.item_row = torch.empty(self.n_items, dtype=torch.int32, device=dev)
            self.item_start = torch.empty(self.n_items, dtype=torch.int32, device=dev)
            _lib.check(lib.b2g_csr_chunk_fill(self.rowptr.data_ptr(), self.n_rows, chunk, self.row_item_ptr.data_ptr(),
                                              self.item_row.data_ptr(), self.item_start.data_ptr(), _stream()),
                       "b2g_csr_chunk_fill")


class Relation:
    """One edge type (src, rel, dst): CSR keyed by destination (forward aggregation) and by source
    (transposed aggregation for the backward pass -- no float atomics)."""

    def __init__(self, edge_type: EdgeType, edge_index: torch.Tensor, n_src: int, n_dst: int):
        self.edge_type = edge_type
        self.n_src, self.n_dst = int(n_src), int(n_dst)
        self.n_edges = int(edge_index.shape[1])
        self.by_dst = CSR(edge_index[1], edge_index[0], n_dst, n_src)
        self.by_src = CSR(edge_index[0], edge_index[1], n_src, n_dst)
        self._dense = None
        self._dense_checked = False

    @property
    def inv_deg_dst(self):
        return self.by_dst.inv_deg

    def dense(self, d: int) -> Optional["DenseAdjacency"]:
        """Dense [n_big, pad] adjacency for the tensor-core formulation (None when the relation does not qualify:
        small side > 256 nodes, big side too small to matter, sparser than the break-even, or over the memory budget)."""
        if self._dense_checked:
            return self._dense
        self._dense_checked = True
        lib = _lib.load()
        small_is_src = self.n_src < self.n_dst
        n_small, n_big = (self.n_src, self.n_dst) if small_is_src else (self.n_dst, self.n_src)
        pad = (n_small + 31) // 32 * 32
        if n_small > 256 or n_big < 4096 or self.n_edges == 0:
            return None
        # dense row (pad * 4 B) vs gathered rows (deg * d * 4 B): only worth it when the relation is dense enough
        if self.n_edges / n_big * d < pad:
            return None
        if not (lib.b2g_linear_fwd_tc_supported(n_big, d, pad) and lib.b2g_linear_bwd_weight_tc_supported(n_big, pad, d)):
            return None
        nbytes = n_big * pad * 4
        if DenseAdjacency.bytes_in_use + nbytes > DenseAdjacency.budget_bytes:
            return None
        csr = self.by_dst if small_is_src else self.by_src          # rows = the big side
        mat = torch.empty((n_big, pad), dtype=torch.float32, device=csr.rowptr.device)
        row_val = csr.inv_deg if small_is_src else None              # mean onto the big side: rows pre-scaled by 1/deg
        _lib.check(lib.b2g_dense_adjacency(csr.rowptr.data_ptr(), csr.col.data_ptr(), None if row_val is None else row_val.data_ptr(),
                                           n_big, pad, mat.data_ptr(), _stream()), "b2g_dense_adjacency")
        DenseAdjacency.bytes_in_use += nbytes
        self._dense = DenseAdjacency(mat, pad, n_small, n_big, small_is_src)
        return self._dense


class DenseAdjacency:
    """mat [n_big, pad]: entries 1 (big side is the source: plain adjacency) or 1/deg_big (big side is the destination:
    row-normalised adjacency); columns >= n_small are zero."""
    budget_bytes = 8 << 30
    bytes_in_use = 0

    def __init__(self, mat, pad, n_small, n_big, big_is_dst):
        self.mat, self.pad, self.n_small, self.n_big, self.big_is_dst = mat, pad, n_small, n_big, big_is_dst
        self._nbytes = int(mat.numel()) * 4

    def __del__(self):                       # give the budget back when the relation (graph) dies
        try:
            DenseAdjacency.bytes_in_use -= self._nbytes
        except Exception:
            pass


def bit_layout(sizes):
    """Column offsets of consecutive relations (sizes[i] type nodes each) on the concatenated bit axis such that no 32-bit
    word contains more than one relation boundary, and the per-word description the kernels need.
    Returns (offsets, nw, rel_a, rel_b, split)."""
    offs, off, last_boundary_word = [], 0, -1
    for n in sizes:
        start = off
        if start % 32 != 0:
            w = start // 32
            if w == last_boundary_word:            # this word already holds a boundary: start the relation at the next word
                start = (w + 1) * 32
            else:
                last_boundary_word = w
        offs.append(start)
        off = start + int(n)
    nw = max(1, (off + 31) // 32)
    owner = [-1] * (32 * nw)
    for i, (o, n) in enumerate(zip(offs, sizes)):
        for c in range(o, o + int(n)):
            owner[c] = i
    rel_a, rel_b, split = [], [], []
    for w in range(nw):
        seen = []
        for b in range(32):
            r = owner[32 * w + b]
            if r >= 0 and (not seen or seen[-1][0] != r):
                seen.append((r, b))
        assert len(seen) <= 2, "bit layout: more than one relation boundary in a word"
        if not seen:
            rel_a.append(0); rel_b.append(0); split.append(32)
        elif len(seen) == 1:
            rel_a.append(seen[0][0]); rel_b.append(seen[0][0]); split.append(32)
        else:
            rel_a.append(seen[0][0]); rel_b.append(seen[1][0]); split.append(seen[1][1])
    return offs, nw, rel_a, rel_b, split


class PatientBits:
    """Bit-matrix form of every relation between the hub node type (patient) and the small vocabularies
    (graph_build.py:216-247: every edge is patient <-> lab / diagnosis / medication) for the single-launch HeteroConv
    kernels (csrc/layer_tc.cu).  `types[i]` owns the bit columns [offs[i], offs[i] + n_i):

      bits_in   rows = patients, bit (p, t) set when edge t -> p exists in the relation (t, *, patient)   (forward out_p, dY)
      bits_out  same for the relation (patient, *, t)                                                    (type sums, dx_p)

    The reference builds every reverse relation as edge_index.flip(0) (graph_build.py:222,235,247), so the two matrices
    are normally identical and share storage."""

    MAX_TYPES = 4

    def __init__(self, gi: "GraphIndex", hub: str):
        lib = _lib.load()
        self.hub = hub
        n_hub = gi.node_counts[hub]
        self.types, self.in_rel, self.out_rel = [], [], []
        self.ok = True
        for et, rel in gi.relations.items():
            src, _, dst = et
            if src == hub and dst == hub:
                self.ok = False
            other = dst if src == hub else (src if dst == hub else None)
            if other is None:
                self.ok = False          # a relation that does not touch the hub type: not this graph family
                continue
            if other not in self.types:
                self.types.append(other)
                self.in_rel.append(None)
                self.out_rel.append(None)
            i = self.types.index(other)
            slot = self.out_rel if src == hub else self.in_rel
            if slot[i] is not None:
                self.ok = False          # two relations between the same pair of types
            slot[i] = rel
        if not self.types or len(self.types) > self.MAX_TYPES:
            self.ok = False
        if not self.ok:
            return
        sizes = [gi.node_counts[t] for t in self.types]
        self.sizes = sizes
        self.offs, self.nw, rel_a, rel_b, split = bit_layout(sizes)
        if self.nw > 24:
            self.ok = False
            return
        self.layout = _lib.BitLayoutT()
        self.layout.nw = self.nw
        for k in range(self.nw):
            self.layout.rel_a[k], self.layout.rel_b[k], self.layout.split[k] = rel_a[k], rel_b[k], split[k]
        dev = next(iter(gi.relations.values())).by_dst.rowptr.device
        self.device = dev

        def build(rels, by):
            if all(r is None for r in rels):
                return None
            bits = torch.zeros((n_hub, self.nw), dtype=torch.int32, device=dev)
            for r, off in zip(rels, self.offs):
                if r is None or r.n_edges == 0:
                    continue
                csr = getattr(r, by)
                _lib.check(lib.b2g_adj_bits_build(csr.rowptr.data_ptr(), csr.col.data_ptr(), n_hub, self.nw, int(off), bits.data_ptr(),
                                                  _stream()), "b2g_adj_bits_build")
            return bits

        self.bits_in = build(self.in_rel, "by_dst")         # rows keyed by the destination (patient), col = source type node
        self.bits_out = build(self.out_rel, "by_src")       # rows keyed by the source (patient), col = destination type node
        if self.bits_in is not None and self.bits_out is not None and bool(torch.equal(self.bits_in, self.bits_out)):
            self.bits_out = self.bits_in
        self._col_scale = None

    def rscale_in(self):
        """per type: 1/deg of the patient in the relation type -> patient (PyG mean), or None"""
        return [None if r is None else r.by_dst.inv_deg for r in self.in_rel]

    def col_scale_out(self):
        """[32 nw] 1/deg of every type node in its relation patient -> type (global degrees in multi-GPU mode), 0 on padding"""
        if self._col_scale is None:
            cs = torch.zeros(32 * self.nw, dtype=torch.float32, device=self.device)
            for r, off, n in zip(self.out_rel, self.offs, self.sizes):
                if r is not None:
                    cs[off:off + n] = r.by_dst.inv_deg
            self._col_scale = cs
        return self._col_scale

    def inv_deg_out(self):
        return [None if r is None else r.by_dst.inv_deg for r in self.out_rel]


class GraphIndex:
    """All relations of one heterogeneous graph + node counts."""

    def __init__(self, data):
        self.node_counts = {nt: int(data[nt].num_nodes) for nt in data.node_types}
        self.relations: Dict[EdgeType, Relation] = {}
        self._ptrs = {}
        for et, ei in data.edge_index_dict.items():
            et = tuple(et)
            src, _, dst = et
            _require_cuda(ei, f"edge_index of {et}")
            if ei.dim() != 2 or ei.shape[0] != 2:
                raise ValueError(f"edge_index of {et} must have shape [2, E], got {tuple(ei.shape)}")
            try:
                self.relations[et] = Relation(et, ei, self.node_counts[src], self.node_counts[dst])
            except _lib.B2GError as exc:
                if "out of range" in str(exc):
                    raise ValueError(f"edge_index of {et} has endpoints outside the node range") from exc
                raise
            self._ptrs[et] = (ei.data_ptr(), int(ei.shape[1]))
        lab_rel = self.relations.get(("patient", "has_lab", "lab"))
        # model.py:297-298: torch.bincount(has_lab.edge_index[0], minlength=N_patient) == row degrees of the by-source CSR
        self.patient_lab_degree: Optional[torch.Tensor] = lab_rel.by_src.deg if lab_rel is not None else None
        self._hub_bits: Dict[str, PatientBits] = {}

    def hub_bits(self, hub: str = "patient") -> Optional["PatientBits"]:
        """Bit-matrix adjacency around `hub` for the fused layer kernels (None when the graph is not hub-shaped)."""
        if hub not in self._hub_bits:
            self._hub_bits[hub] = PatientBits(self, hub) if hub in self.node_counts else None
        pb = self._hub_bits[hub]
        return pb if (pb is not None and pb.ok) else None

    def matches(self, data) -> bool:
        eid = data.edge_index_dict
        if set(map(tuple, eid.keys())) != set(self._ptrs):
            return False
        return all(self._ptrs[tuple(k)] == (v.data_ptr(), int(v.shape[1])) for k, v in eid.items())


_GRAPH_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
_GRAPH_CACHE_STRONG: Dict[int, Tuple[object, GraphIndex]] = {}


def graph_index(data) -> GraphIndex:
    """Cached GraphIndex for a HeteroData-like object (rebuilt if its edge tensors were replaced)."""
    try:
        gi = _GRAPH_CACHE.get(data)
    except TypeError:
        ent = _GRAPH_CACHE_STRONG.get(id(data))
        gi = ent[1] if ent is not None and ent[0] is data else None
    if gi is not None and gi.matches(data):
        return gi
    gi = GraphIndex(data)
    try:
        _GRAPH_CACHE[data] = gi
    except TypeError:
        _GRAPH_CACHE_STRONG[id(data)] = (data, gi)
    return gi


class PairIndex:
    """CSR views of a list of (patient, lab) prediction pairs, used by the decoder's backward to turn the
    per-pair gradients into per-patient / per-lab sums without atomics (replaces autograd's
    ``index_put_(accumulate=True)`` behind model.py:305-309)."""

    def __init__(self, patient_idx: torch.Tensor, lab_idx: torch.Tensor, n_patient: int, n_lab: int):
        self.m = int(patient_idx.numel())
        self.patient_idx = patient_idx.contiguous()
        self.lab_idx = lab_idx.contiguous()
        self.n_patient, self.n_lab = int(n_patient), int(n_lab)
        self._by_patient = self._by_lab = None       # built on first use: only the decoder's BACKWARD needs them

    @property
    def by_patient(self) -> "CSR":
        if self._by_patient is None:
            self._by_patient = CSR(self.patient_idx, self.patient_idx, self.n_patient, self.n_patient, col_is_eid=True)
        return self._by_patient

    @property
    def by_lab(self) -> "CSR":
        if self._by_lab is None:
            self._by_lab = CSR(self.lab_idx, self.lab_idx, self.n_lab, self.n_lab, col_is_eid=True)
        return self._by_lab

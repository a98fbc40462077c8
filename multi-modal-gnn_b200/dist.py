"""Patient-partitioned multi-GPU execution.  One process per GPU, torch.distributed (NCCL over
NVLink on the B200 box, gloo in the CPU tests) for the plumbing.

Every edge of the graph is patient <-> {lab, diagnosis, medication}, so a contiguous patient range owns all of its
edges, embedding rows, activations, prediction pairs and gradients; the type tables (<= 135 KB), all dense weights and
BatchNorm parameters are replicated.  To reproduce the single-GPU (= reference, full-batch) numbers exactly the ranks
exchange, all as fp32/fp64 SUM all-reduces of tiny buffers:

  forward   per-layer partial neighbour sums onto the type nodes          [N_type, d]      (PartialToReplicatedFn)
            patient BatchNorm statistics                                  [2, d] fp64      (ops.SyncBNActDropFn)
  backward  gradients of replicated tensors consumed by rank-local work   [N_type, d|64]   (ReplicatedToLocalFn)
            patient BatchNorm backward statistics                         [2, d] fp64
  step      one flat gradient all-reduce                                  483,970 floats

Replicated computations (everything on type rows) are executed identically on every rank; their parameters see the
*full* gradient on every rank, so those uses are wrapped in ScaleGradFn(1/world) to make the final SUM all-reduce exact.
The functions in this file are device-agnostic (they run on CPU tensors with gloo in tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch.autograd import Function

from .heterodata import HeteroGraph


class PeerComm:
    """libb2g's peer-memory communicator (include/b2g.h section (f), csrc/peer.cuh): every rank's symmetric region is
    mapped into every other rank of the node through CUDA IPC, and the small all-reduces of the step become one-shot
    kernels that store / load over NVLink directly -- fused into the producing kernel for the BatchNorm statistics."""

    def __init__(self, group, device: torch.device):
        import ctypes
        from . import _lib
        self._lib_mod = _lib
        lib = _lib.load()
        self.lib = lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise _lib.B2GError("PeerComm supports at most 8 ranks (one NVSwitch node)")
        region = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)

        def agree(ok: bool, what: str):
            """Every rank learns whether the step succeeded everywhere (a rank that raised alone would leave the others
            hanging in the next collective)."""
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                msg = lib.b2g_last_error()
                raise _lib.B2GError(f"peer-memory communicator: {what} failed on at least one rank"
                                    + (f" (this rank: {msg.decode()})" if (not ok and msg) else ""))

        with torch.cuda.device(device):
            rc = lib.b2g_comm_local_alloc(ctypes.byref(region), handle)
            agree(rc == 0, "cudaMalloc / cudaIpcGetMemHandle of the symmetric region")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)    # also the "everybody allocated" rendezvous
            comm = ctypes.c_void_p()
            rc = lib.b2g_comm_create(self.rank, self.world, region, b"".join(handles), ctypes.byref(comm))
            agree(rc == 0, "cudaIpcOpenMemHandle of a peer's region")          # also: every rank has mapped every region
        self.handle = comm
        self.max_bytes = int(lib.b2g_comm_max_bytes())

    def usable(self, t: torch.Tensor) -> bool:
        nbytes = t.numel() * t.element_size()
        return (t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64) and nbytes > 0
                and nbytes % 16 == 0 and t.data_ptr() % 16 == 0 and nbytes <= self.max_bytes)

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        fn = self.lib.b2g_comm_allreduce_f32 if t.dtype == torch.float32 else self.lib.b2g_comm_allreduce_f64
        self._lib_mod.check(fn(self.handle, t.data_ptr(), t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream),
                            "b2g_comm_allreduce")
        return t

    def check(self):
        """SYNC.  Raises if a wait inside a kernel timed out (a rank did not arrive)."""
        if self.lib.b2g_comm_error(self.handle):
            raise self._lib_mod.B2GError("peer-memory exchange timed out: a rank did not reach the rendezvous")


class DistContext:
    def __init__(self, group=None, sharded_type: str = "patient", device: Optional[torch.device] = None, peer: Optional[bool] = None):
        """``device``: this rank's CUDA device -- enables the peer-memory communicator (NVLink loads / stores from our own
        kernels) for the step's small SUM all-reduces unless ``peer=False`` / B2G_PEER_COMM=0; torch.distributed stays the
        path for everything else (CPU tensors under gloo, integer and MAX reductions, oversized payloads)."""
        import os
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.sharded_type = sharded_type
        self.global_rows: Dict[int, int] = {}      # local row count of the sharded type -> global row count
        self.n_collectives = 0
        self.n_peer = 0                            # how many of them went through the peer-memory kernels
        self.n_collectives_per_step = None         # set by Trainer._capture: exchanges inside one captured step
        self.n_peer_per_step = None
        self.trace = [] if os.environ.get("B2G_TRACE_COLLECTIVES") else None   # debugging aid
        if peer is None:
            peer = os.environ.get("B2G_PEER_COMM", "1") != "0"
        self.peer: Optional[PeerComm] = None
        self.peer_unavailable: Optional[str] = None
        if peer and device is not None and torch.device(device).type == "cuda" and self.world > 1:
            try:
                self.peer = PeerComm(self.group, torch.device(device))
            except Exception as exc:      # e.g. CUDA IPC not permitted in this container, GPUs without peer access: every
                import sys                # rank lands here together (PeerComm agrees on failures), the exchanges stay on NCCL
                self.peer_unavailable = str(exc)
                print(f"[multi-modal-gnn_b200] peer-memory communicator unavailable, using NCCL for the step's exchanges: {exc}",
                      file=sys.stderr, flush=True)

    def all_reduce_(self, t: torch.Tensor, op=dist.ReduceOp.SUM) -> torch.Tensor:
        if self.trace is not None:
            self.trace.append((self.n_collectives, tuple(t.shape), str(t.dtype)))
        self.n_collectives += 1
        if self.peer is not None and op == dist.ReduceOp.SUM and self.peer.usable(t):
            self.n_peer += 1
            return self.peer.all_reduce_(t)
        dist.all_reduce(t, op=op, group=self.group)
        return t

    def count_fused(self):
        """An exchange that ran inside a compute kernel (ops.SyncBNActDropFn on the peer path)."""
        self.n_collectives += 1
        self.n_peer += 1

    def global_row_count(self, local_rows: int, device) -> int:
        """Number of rows of the sharded node type over all ranks (one tiny all-reduce, cached)."""
        if local_rows not in self.global_rows:
            t = torch.tensor([local_rows], dtype=torch.int64, device=device)
            self.all_reduce_(t)
            self.global_rows[local_rows] = int(t.item())
        return self.global_rows[local_rows]


class ReplicatedToLocalFn(Function):
    """Identity in forward; in backward the gradient (each rank holds only its share) is summed over ranks, so that
    the replicated producer sees the full gradient on every rank."""

    @staticmethod
    def forward(ctx, x, dctx: DistContext):
        ctx.dctx = dctx
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        ctx.dctx.all_reduce_(g)
        return g, None


class PartialToReplicatedFn(Function):
    """Sum of per-rank partial results in forward (every rank ends with the full tensor); identity in backward (the
    gradient of a replicated tensor is already complete and identical on every rank)."""

    @staticmethod
    def forward(ctx, x, dctx: DistContext):
        y = x.contiguous().clone()
        dctx.all_reduce_(y)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


class PartialToReplicatedManyFn(Function):
    """PartialToReplicatedFn for several tensors with ONE all-reduce (the three per-layer partial type sums)."""

    @staticmethod
    def forward(ctx, dctx: DistContext, *xs):
        flat = _flat_padded([x.reshape(-1) for x in xs])
        dctx.all_reduce_(flat)
        outs, off = [], 0
        for x in xs:
            outs.append(flat[off:off + x.numel()].view_as(x))
            off += x.numel()
        ctx.set_materialize_grads(False)     # an unused aggregate keeps "no gradient" (None), see ScaleGradFn
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        return (None, *gs)


class ReplicatedToLocalManyFn(Function):
    """ReplicatedToLocalFn for several tensors: their gradients are summed over ranks with ONE all-reduce."""

    @staticmethod
    def forward(ctx, dctx: DistContext, *xs):
        ctx.dctx = dctx
        return tuple(x.view_as(x) for x in xs)

    @staticmethod
    def backward(ctx, *gs):
        flat = _flat_padded([g.reshape(-1) for g in gs])
        ctx.dctx.all_reduce_(flat)
        outs, off = [], 0
        for g in gs:
            outs.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return (None, *outs)


def _flat_padded(chunks):
    """torch.cat of 1-D fp32 chunks, zero-padded to a multiple of 4 elements (16 bytes: the peer all-reduce's unit)."""
    n = sum(int(c.numel()) for c in chunks)
    pad = (-n) % 4
    if pad:
        chunks = list(chunks) + [torch.zeros(pad, dtype=chunks[0].dtype, device=chunks[0].device)]
    return torch.cat(chunks)


class ScaleGradFn(Function):
    @staticmethod
    def forward(ctx, x, s: float):
        ctx.s = s
        ctx.set_materialize_grads(False)     # "no gradient" must stay None (SURVEY.md note N8), never become zeros
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (None if g is None else g * ctx.s), None


def replicated_to_local(x, dctx: Optional[DistContext]):
    return x if dctx is None else ReplicatedToLocalFn.apply(x, dctx)


def partial_to_replicated(x, dctx: Optional[DistContext]):
    return x if dctx is None else PartialToReplicatedFn.apply(x, dctx)


def partial_to_replicated_many(xs, dctx: Optional[DistContext]):
    return list(xs) if (dctx is None or not xs) else list(PartialToReplicatedManyFn.apply(dctx, *xs))


def replicated_to_local_many(xs, dctx: Optional[DistContext]):
    return list(xs) if (dctx is None or not xs) else list(ReplicatedToLocalManyFn.apply(dctx, *xs))


def rep_param(p, dctx: Optional[DistContext]):
    """A parameter (or replicated leaf) used inside a computation that every rank repeats identically."""
    return p if (dctx is None or p is None) else ScaleGradFn.apply(p, 1.0 / dctx.world)


# ----------------------------------------------------------------------------------------------------------------------
def partition_bounds(data, world: int, sharded_type: str = "patient") -> torch.Tensor:
    """Contiguous ranges of the sharded node type, balanced by incident-edge count.
    Returns int64[world + 1] boundaries."""
    n = int(data[sharded_type].num_nodes)
    load = torch.ones(n, dtype=torch.float64)
    for et, ei in data.edge_index_dict.items():
        src, _, dst = et
        if src == sharded_type:
            load += torch.bincount(ei[0].cpu(), minlength=n).double()
        if dst == sharded_type:
            load += torch.bincount(ei[1].cpu(), minlength=n).double()
    csum = load.cumsum(0)
    targets = csum[-1] * torch.arange(1, world, dtype=torch.float64) / world
    cuts = torch.searchsorted(csum, targets).clamp(1, n - 1) if world > 1 else torch.empty(0, dtype=torch.int64)
    bounds = torch.cat([torch.zeros(1, dtype=torch.int64), cuts.long(), torch.tensor([n])])
    return torch.cummax(bounds, 0)[0]


def partition_graph(data, world: int, rank: int, sharded_type: str = "patient") -> Tuple[HeteroGraph, dict]:
    """Local sub-graph of `rank`: its range of the sharded node type (ids relabelled to start at 0), every other node
    type in full, and exactly the edges incident to its range, in their original relative order (so the CSR build and
    the edge-level split masks stay consistent).  Returns (local graph, info) with info['edge_ids'][edge_type] = positions
    of the kept edges in the global edge list and info['range'] = (p0, p1)."""
    bounds = partition_bounds(data, world, sharded_type)
    p0, p1 = int(bounds[rank]), int(bounds[rank + 1])
    g = HeteroGraph()
    for nt in data.node_types:
        g[nt].num_nodes = (p1 - p0) if nt == sharded_type else int(data[nt].num_nodes)
    info = {"range": (p0, p1), "bounds": bounds, "edge_ids": {}, "global_nodes": {nt: int(data[nt].num_nodes) for nt in data.node_types}}
    for et in data.edge_types:
        src, _, dst = et
        ei = data[et].edge_index
        keep = torch.ones(ei.shape[1], dtype=torch.bool, device=ei.device)
        if src == sharded_type:
            keep &= (ei[0] >= p0) & (ei[0] < p1)
        if dst == sharded_type:
            keep &= (ei[1] >= p0) & (ei[1] < p1)
        ids = keep.nonzero().squeeze(1)
        loc = ei[:, ids].clone()
        if src == sharded_type:
            loc[0] -= p0
        if dst == sharded_type:
            loc[1] -= p0
        g[et].edge_index = loc.contiguous()
        if "edge_attr" in data[et]:
            g[et].edge_attr = data[et].edge_attr[ids].contiguous()
        info["edge_ids"][tuple(et)] = ids
    return g, info


class PairRoute:
    """Routing of arbitrary GLOBAL (patient, lab) prediction pairs to the ranks that own the patients ("pairs are
    routed to the GPU owning the patient"; bulk imputation, BASELINE config 5) and of the predictions back to the asking rank.

        route = PairRoute(patient_idx_global, lab_idx, bounds, dctx)     # all-to-all of the index lists
        pred_local = model.predict_lab_values(local_graph, route.patient_local, route.lab_local)
        pred = route.gather_back(pred_local)                             # same order as the caller's pair list
    """

    def __init__(self, patient_idx: torch.Tensor, lab_idx: torch.Tensor, bounds: torch.Tensor, dctx: DistContext):
        self.dctx = dctx
        world, dev = dctx.world, patient_idx.device
        b = bounds.to(dev)
        owner = torch.bucketize(patient_idx, b[1:-1].contiguous(), right=True)            # rank whose [b[r], b[r+1]) holds the patient
        order = torch.argsort(owner, stable=True)
        self.order = order
        self.send_counts = torch.bincount(owner, minlength=world)
        recv_counts = torch.empty_like(self.send_counts)
        dist.all_to_all_single(recv_counts, self.send_counts, group=dctx.group)
        self.recv_counts = recv_counts
        sc, rc = self.send_counts.tolist(), recv_counts.tolist()
        self._sc, self._rc = sc, rc
        payload = torch.stack([patient_idx[order], lab_idx[order]], 1).contiguous()        # [n, 2] int64
        got = torch.empty((sum(rc), 2), dtype=payload.dtype, device=dev)
        dist.all_to_all_single(got, payload, output_split_sizes=rc, input_split_sizes=sc, group=dctx.group)
        self.patient_local = (got[:, 0] - b[dctx.rank]).contiguous()
        self.lab_local = got[:, 1].contiguous()

    def gather_back(self, pred_local: torch.Tensor) -> torch.Tensor:
        back = torch.empty(sum(self._sc), dtype=pred_local.dtype, device=pred_local.device)
        dist.all_to_all_single(back, pred_local.contiguous(), output_split_sizes=self._sc, input_split_sizes=self._rc, group=self.dctx.group)
        out = torch.empty_like(back)
        out[self.order] = back
        return out


def globalize_degrees(graph_index, dctx: DistContext):
    """Mean aggregation onto replicated node types divides by the GLOBAL neighbour count: all-reduce the per-type-node
    degrees once and overwrite the local CSR's deg / inv_deg (integer all-reduce: exact)."""
    if getattr(graph_index, "_globalized", False):
        return
    for et, rel in graph_index.relations.items():
        if et[2] != dctx.sharded_type:
            deg = rel.by_dst.deg.clone()
            dctx.all_reduce_(deg)
            rel.by_dst.deg = deg
            rel.by_dst.inv_deg = 1.0 / deg.clamp(min=1).to(torch.float32)
    graph_index._globalized = True


def gradient_pattern(params, dctx: DistContext):
    """SYNC (host round trip).  Which parameters have a gradient on at least one rank -- static for a given model and
    graph, so the trainer computes it once; parameters whose gradient is None on every rank (dead branches, SURVEY.md
    note N8) stay None so that Adam keeps skipping them like the reference does."""
    params = list(params)
    if not params:
        return []
    present = torch.tensor([0 if p.grad is None else 1 for p in params], dtype=torch.int32, device=params[0].device)
    dctx.all_reduce_(present, op=dist.ReduceOp.MAX)
    return [bool(v) for v in present.tolist()]


def allreduce_gradients(params, dctx: DistContext, present=None):
    """The step's gradient exchange: one flat SUM all-reduce (cat -> all-reduce -> one multi-tensor copy back).  With a
    precomputed ``present`` pattern there is no host synchronisation, so the call can be captured into the step's graph."""
    params = list(params)
    if not params:
        return
    if present is None:
        present = gradient_pattern(params, dctx)
    used = [p for p, has in zip(params, present) if has]
    if not used:
        return
    for p in used:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = _flat_padded([p.grad.reshape(-1) for p in used])
    dctx.all_reduce_(flat)
    views, off = [], 0
    for p in used:
        n = p.numel()
        views.append(flat[off:off + n].view_as(p))
        off += n
    torch._foreach_copy_([p.grad for p in used], views)

"""autograd.Function wrappers over libb2g's C ABI.  PyTorch here is plumbing only: device memory, the
current CUDA stream and the autograd tape; every arithmetic kernel is ours (csrc/*.cu).

Each Function cites the reference expression it replaces.  There is no CPU implementation: passing a
CPU tensor raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch
from torch.autograd import Function

from . import _lib
from .graph import CSR, PairIndex, Relation, workspace, _stream

ACT_CODES = {"none": 0, "relu": 1, "leaky_relu": 2, "elu": 3}

# Dense-layer arithmetic: "tf32" = large-M linears on the tcgen05 tensor cores (fp32 operands read as TF32, fp32
# accumulation in TMEM; ~1e-3 relative, north_star tolerance 1e-2), "fp32" = exact-fp32 SIMT kernels (parity mode).
PRECISION = os.environ.get("B2G_PRECISION", "tf32")
TC_MIN_ROWS = 512


def set_precision(mode: str):
    global PRECISION
    if mode not in ("tf32", "fp32"):
        raise ValueError(f"precision must be 'tf32' or 'fp32', got {mode!r}")
    PRECISION = mode


def _use_tc(lib, m, n, k) -> bool:
    return PRECISION == "tf32" and m >= TC_MIN_ROWS and bool(lib.b2g_linear_fwd_tc_supported(m, n, k))


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.B2GError(f"{name} must be a CUDA tensor (got {t.device}); this package has no CPU path")
    if t.dtype != torch.float32:
        raise _lib.B2GError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# Optional per-call profiler: when PROFILE is a list, every library call is bracketed by CUDA events on the
# launching (current) stream and appended as (name, start, end, algorithmic_bytes, flops).  bench.py uses it for
# one instrumented step to find the dominant kernel and its live duration.
PROFILE: Optional[list] = None
_NEXT_COST = [0, 0]


def cost(nbytes: int = 0, flops: int = 0):
    """Declare the algorithmic bytes / flops of the next library call (read only when profiling)."""
    _NEXT_COST[0], _NEXT_COST[1] = int(nbytes), int(flops)


def _run(name, fn, *args):
    prof = PROFILE
    if prof is None:
        rc = fn(*args)
    else:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        prof.append((name, s, e, _NEXT_COST[0], _NEXT_COST[1]))
        _NEXT_COST[0] = _NEXT_COST[1] = 0
    if rc != 0:
        _lib.check(rc, name)


# --------------------------------------------------------------------------------------------------
# Column sums of a gradient tensor that its PRODUCER already computed (the BatchNorm backward pass writes dx and its
# column sums in one pass; the nn.Linear in front of the BatchNorm needs exactly those sums as its bias gradient).
# The hand-over is a tag on the tensor object, valid only while the tensor is unmodified (version counter) -- if autograd
# accumulated into it or handed over a different tensor, the consumer simply reduces the columns itself.
# --------------------------------------------------------------------------------------------------
def _tag_colsum(t: torch.Tensor, colsum: torch.Tensor):
    t._b2g_colsum = (colsum, t._version, t.data_ptr())


def _tagged_colsum(t: torch.Tensor):
    tag = getattr(t, "_b2g_colsum", None)
    if tag is None or tag[1] != t._version or tag[2] != t.data_ptr():
        return None
    return tag[0]


# BatchNorm column statistics {sum x, sum x^2} (fp64 [2 d]) that the PRODUCER of a tensor computed in its epilogue
# (csrc/layer_tc.cu): the BatchNorm that follows skips its own pass over the rows.  Same validity rule as above.
def _tag_bnsums(t: torch.Tensor, sums: torch.Tensor):
    t._b2g_bnsums = (sums, t._version, t.data_ptr())


def _tagged_bnsums(t: torch.Tensor):
    tag = getattr(t, "_b2g_bnsums", None)
    if tag is None or tag[1] != t._version or tag[2] != t.data_ptr():
        return None
    return tag[0]


# --------------------------------------------------------------------------------------------------
# raw (non-differentiable) kernel calls
# --------------------------------------------------------------------------------------------------
def linear_fwd_(x, w, b, y, accumulate=False):
    lib = _lib.load()
    m, k = x.shape
    n = w.shape[0]
    cost(4 * (m * k + n * k + m * n * (2 if accumulate else 1)), 2 * m * n * k)
    if _use_tc(lib, m, n, k) and (b is None or b.data_ptr() % 16 == 0):
        _run("b2g_linear_fwd_tc", lib.b2g_linear_fwd_tc, x.data_ptr(), w.data_ptr(), _ptr(b), m, n, k, y.data_ptr(), int(accumulate),
             _stream())
    else:
        _run("b2g_linear_fwd", lib.b2g_linear_fwd, x.data_ptr(), w.data_ptr(), _ptr(b), m, n, k, y.data_ptr(), int(accumulate), _stream())
    return y


def linear_bwd_input_(dy, w, dx, accumulate=False):
    lib = _lib.load()
    m, n = dy.shape
    k = w.shape[1]
    if _use_tc(lib, m, k, n):        # dx = dy W  ==  linear(dy, W^T): same tcgen05 kernel on the transposed weight
        wt = torch.empty((k, n), dtype=torch.float32, device=w.device)
        _run("b2g_transpose", lib.b2g_transpose, w.data_ptr(), n, k, wt.data_ptr(), _stream())
        cost(4 * (m * n + n * k + m * k * (2 if accumulate else 1)), 2 * m * n * k)
        _run("b2g_linear_bwd_input_tc", lib.b2g_linear_fwd_tc, dy.data_ptr(), wt.data_ptr(), None, m, k, n, dx.data_ptr(), int(accumulate),
             _stream())
        return dx
    cost(4 * (m * n + n * k + m * k * (2 if accumulate else 1)), 2 * m * n * k)
    _run("b2g_linear_bwd_input", lib.b2g_linear_bwd_input, dy.data_ptr(), w.data_ptr(), m, n, k, dx.data_ptr(), int(accumulate), _stream())
    return dx


def linear_bwd_weight_(dy, x, dw, db):
    lib = _lib.load()
    m, n = dy.shape
    k = x.shape[1]
    if db is not None:
        ready = _tagged_colsum(dy)           # the producer of dy already reduced its columns
        if ready is not None and ready.shape == db.shape:
            db.copy_(ready)
            db = None
    if PRECISION == "tf32" and m >= TC_MIN_ROWS and lib.b2g_linear_bwd_weight_tc_supported(m, n, k) and (db is None or n in (32, 64, 128, 256)):
        ws = workspace(lib.b2g_linear_bwd_weight_tc_ws_bytes(m, n, k), dy.device)
        cost(4 * (m * n + m * k + n * k), 2 * m * n * k)
        _run("b2g_linear_bwd_weight_tc", lib.b2g_linear_bwd_weight_tc, dy.data_ptr(), x.data_ptr(), m, n, k, dw.data_ptr(), ws.data_ptr(),
             ws.numel(), _stream())
        if db is not None:
            ws = workspace(lib.b2g_bn_ws_bytes(n), dy.device)
            cost(4 * m * n)
            _run("b2g_col_sums", lib.b2g_col_sums, dy.data_ptr(), m, n, db.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        return
    nb = lib.b2g_linear_bwd_weight_ws_bytes(m, n, k)
    ws = workspace(nb, dy.device)
    cost(4 * (m * n + m * k + n * k), 2 * m * n * k)
    _run("b2g_linear_bwd_weight", lib.b2g_linear_bwd_weight, dy.data_ptr(), x.data_ptr(), m, n, k, dw.data_ptr(), _ptr(db), ws.data_ptr(),
                                         ws.numel(), _stream())


# ---- grouped small GEMMs (type-node rows) -----------------------------------------------------------------------------------
GROUP_MAX = 12
SMALL_ROWS = 1024


def small_gemm_group_(problems):
    """problems: list of dicts with keys c, a, b (tensors), optional a2, b2, bias, and flags a_t, b_nk, acc; m/n/k are
    taken from c and the operands.  One launch per <= 12 problems."""
    lib = _lib.load()
    for s0 in range(0, len(problems), GROUP_MAX):
        grp = problems[s0:s0 + GROUP_MAX]
        arr = (_lib.GemmProblemT * len(grp))()
        nbytes = flops = 0
        for i, p in enumerate(grp):
            c, a, b = p["c"], p["a"], p["b"]
            m, n = c.shape
            a_t, b_nk = bool(p.get("a_t", False)), bool(p.get("b_nk", False))
            k = a.shape[0] if a_t else a.shape[1]
            a2, b2 = p.get("a2"), p.get("b2")
            k2 = 0 if a2 is None else (a2.shape[0] if a_t else a2.shape[1])
            arr[i].a, arr[i].b, arr[i].c = a.data_ptr(), b.data_ptr(), c.data_ptr()
            arr[i].a2, arr[i].b2 = _ptr(a2), _ptr(b2)
            arr[i].bias = _ptr(p.get("bias"))
            arr[i].m, arr[i].n, arr[i].k, arr[i].k2 = m, n, k, k2
            arr[i].a_transposed, arr[i].b_is_nk, arr[i].accumulate = int(a_t), int(b_nk), int(bool(p.get("acc", False)))
            nbytes += 4 * (m * n + (m + n) * (k + k2))
            flops += 2 * m * n * (k + k2)
        cost(nbytes, flops)
        _run("b2g_small_gemm_group", lib.b2g_small_gemm_group, arr, len(grp), _stream())


def small_colsum_group_(xs, outs):
    lib = _lib.load()
    for s0 in range(0, len(xs), GROUP_MAX):
        gx, go = xs[s0:s0 + GROUP_MAX], outs[s0:s0 + GROUP_MAX]
        n = len(gx)
        px = (ctypes.c_void_p * n)(*[t.data_ptr() for t in gx])
        po = (ctypes.c_void_p * n)(*[t.data_ptr() for t in go])
        pm = (ctypes.c_int * n)(*[t.shape[0] for t in gx])
        pn = (ctypes.c_int * n)(*[t.shape[1] for t in gx])
        _run("b2g_small_colsum_group", lib.b2g_small_colsum_group, px, po, pm, pn, n, _stream())


# b2g_gather_reduce_staged (source tables staged in shared memory) is bit-identical to b2g_gather_reduce and measured equal in
# speed at the C4 shard (0.86 vs 0.81 ms: the 175 KB of tables are L1-resident for the plain kernel too, and both are bound by
# the per-row dependency chain, profiles/r2_gather_bench_c4s8.json) -- opt-in with B2G_GATHER_STAGED=1
GATHER_STAGED = os.environ.get("B2G_GATHER_STAGED", "0") == "1"
STAGED_MIN_ROWS = 4096          # below this the one-off table copy per CTA costs more than the gathers save
# b2g_gather_reduce_stream serialises the 32 rows of a warp: it wins when there are enough rows to fill the GPU with such warps
# and the rows are short (C4 shard, decoder gradients by patient, 1.25 M rows x 7 entries: 0.37 -> 0.19 ms), and loses otherwise
# (C2, 46 k rows x 75 entries: 0.04 -> 0.12 ms)
GATHER_STREAM = os.environ.get("B2G_GATHER_STREAM", "1") != "0"
STREAM_MIN_ROWS = 262144
STREAM_MAX_AVG_DEG = 10


def gather_reduce_(csrs: Sequence[CSR], xs: Sequence[torch.Tensor], row_scales, col_scales, out: torch.Tensor,
                   accumulate: bool):
    """out[r] (+)= sum_k row_scale_k[r] * sum_{j in row r} col_scale_k[col_j] * x_k[col_j].  Short-row CSRs are
    fused into one launch (<= 4 relations); long-row CSRs go through the chunked two-phase reducer."""
    lib = _lib.load()
    d = out.shape[1]
    n_rows = out.shape[0]
    short = [i for i, c in enumerate(csrs) if not c.long_rows]
    long_ = [i for i, c in enumerate(csrs) if c.long_rows]
    acc = bool(accumulate)
    for s0 in range(0, len(short), 4):
        grp = short[s0:s0 + 4]
        arr = (_lib.RelT * len(grp))()
        for j, i in enumerate(grp):
            c = csrs[i]
            assert c.n_rows == n_rows
            arr[j].rowptr, arr[j].col, arr[j].x = c.rowptr.data_ptr(), c.col.data_ptr(), xs[i].data_ptr()
            arr[j].row_scale = _ptr(row_scales[i])
            arr[j].col_scale = _ptr(col_scales[i])
        # algorithmic bytes: CSR indices once + output rows once (+ read when accumulating) + source tables once
        cost(sum(4 * (csrs[i].n_edges + csrs[i].n_rows + 1) + 4 * d * min(csrs[i].n_vals, csrs[i].n_edges) for i in grp)
             + 4 * n_rows * d * (2 if acc else 1), 2 * d * sum(csrs[i].n_edges for i in grp))
        n_src = (ctypes.c_int * len(grp))(*[int(xs[i].shape[0]) for i in grp])
        if GATHER_STAGED and n_rows >= STAGED_MIN_ROWS and lib.b2g_gather_reduce_staged_supported(n_src, len(grp), d):
            # few-row source tables (lab / diagnosis / medication -> patient): staged in shared memory once per CTA
            _run("b2g_gather_reduce_staged", lib.b2g_gather_reduce_staged, arr, n_src, len(grp), n_rows, d, out.data_ptr(), int(acc), _stream())
        elif GATHER_STREAM and n_rows >= STREAM_MIN_ROWS and sum(csrs[i].n_edges for i in grp) <= STREAM_MAX_AVG_DEG * n_rows:
            # many short rows (the per-patient reduction of the decoder's pair gradients): CSR streamed through registers
            _run("b2g_gather_reduce", lib.b2g_gather_reduce_stream, arr, len(grp), n_rows, d, out.data_ptr(), int(acc), _stream())
        else:
            _run("b2g_gather_reduce", lib.b2g_gather_reduce, arr, len(grp), n_rows, d, out.data_ptr(), int(acc), _stream())
        acc = True
    for i in long_:
        c = csrs[i]
        assert c.n_rows == n_rows
        rel = _lib.RelT()
        rel.rowptr, rel.col, rel.x = c.rowptr.data_ptr(), c.col.data_ptr(), xs[i].data_ptr()
        rel.row_scale = _ptr(row_scales[i])
        rel.col_scale = _ptr(col_scales[i])
        ws = workspace(c.n_items * d * 4, out.device)
        cost(4 * (c.n_edges + c.n_rows + 1) + 4 * d * min(c.n_vals, c.n_edges) + 4 * n_rows * d * (2 if acc else 1),
             2 * d * c.n_edges)
        _run("b2g_gather_reduce_chunked", lib.b2g_gather_reduce_chunked, ctypes.byref(rel), c.item_row.data_ptr(), c.item_start.data_ptr(),
                                                 c.row_item_ptr.data_ptr(), c.n_items, c.chunk, n_rows, d, out.data_ptr(),
                                                 int(acc), ws.data_ptr(), ws.numel(), _stream())
        acc = True
    if not acc:          # no relation at all
        out.zero_()
    return out


# ---- tensor-core (dense adjacency) formulation of the aggregation, tf32 mode only ---------------------------------------
def _dense_of(rel: Relation, d: int):
    return rel.dense(d) if PRECISION == "tf32" else None


def _adj_times_table(dn, table, scale, out, accumulate):
    """out[n_big, d] (+)= D @ (table * scale[:, None])   -- D [n_big, pad], table [n_small, d]"""
    lib = _lib.load()
    d = table.shape[1]
    wt = torch.empty((d, dn.pad), dtype=torch.float32, device=table.device)
    _run("b2g_transpose_pad", lib.b2g_transpose_pad, table.data_ptr(), _ptr(scale), table.shape[0], d, dn.pad, wt.data_ptr(), _stream())
    cost(4 * (dn.n_big * dn.pad + d * dn.pad + dn.n_big * d * (2 if accumulate else 1)), 2 * dn.n_big * d * dn.pad)
    _run("b2g_adjacency_mma_fwd", lib.b2g_linear_fwd_tc, dn.mat.data_ptr(), wt.data_ptr(), None, dn.n_big, d, dn.pad, out.data_ptr(),
         int(accumulate), _stream())
    return out


def _adjT_times_rows(dn, x_big):
    """(D^T @ x_big)[:n_small]  -> [n_small, d] view of a [pad, d] buffer"""
    lib = _lib.load()
    d = x_big.shape[1]
    tmp = torch.empty((dn.pad, d), dtype=torch.float32, device=x_big.device)
    ws = workspace(lib.b2g_linear_bwd_weight_tc_ws_bytes(dn.n_big, dn.pad, d), x_big.device)
    cost(4 * (dn.n_big * dn.pad + dn.n_big * d + dn.pad * d), 2 * dn.n_big * d * dn.pad)
    _run("b2g_adjacency_mma_bwd", lib.b2g_linear_bwd_weight_tc, dn.mat.data_ptr(), x_big.data_ptr(), dn.n_big, dn.pad, d, tmp.data_ptr(),
         ws.data_ptr(), ws.numel(), _stream())
    return tmp[:dn.n_small]


def _row_scale(x, scale):
    lib = _lib.load()
    out = torch.empty_like(x)
    _run("b2g_row_scale", lib.b2g_row_scale, x.data_ptr(), scale.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), _stream())
    return out


def dropout_mask(n: int, p: float, seed: int, stream_id: int, device) -> torch.Tensor:
    """The keep mask (0 or 1/(1-p)) the kernels apply for (seed, stream_id) -- for replay on the CPU oracle."""
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.float32, device=device)
    _run("b2g_dropout_mask", lib.b2g_dropout_mask, n, float(p), int(seed), int(stream_id), out.data_ptr(), _stream())
    return out


# --------------------------------------------------------------------------------------------------
# differentiable ops
# --------------------------------------------------------------------------------------------------
class LinearFn(Function):
    """y = x W^T + b   (nn.Linear; model.py:93-103,373-386 and SAGEConv.lin_l / lin_r)."""

    @staticmethod
    def forward(ctx, x, w, b, want_stats=False):
        x, w = _f32(x, "x"), _f32(w, "weight")
        b = None if b is None else _f32(b, "bias")
        y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        m, k = x.shape
        n = w.shape[0]
        if want_stats and n <= 128 and _use_tc(lib, m, n, k) and (b is None or b.data_ptr() % 16 == 0):
            # the BatchNorm that follows (model.py:93-101) gets its column statistics from this launch's epilogue
            sums = torch.empty(2 * n, dtype=torch.float64, device=x.device)
            ws = workspace(lib.b2g_linear_stats_ws_bytes(n), x.device)
            cost(4 * (m * k + n * k + m * n), 2 * m * n * k)
            _run("b2g_linear_fwd_tc", lib.b2g_linear_fwd_tc_ex, x.data_ptr(), w.data_ptr(), _ptr(b), m, n, k, y.data_ptr(), sums.data_ptr(),
                 ws.data_ptr(), ws.numel(), None, 0.0, _stream())
            _tag_bnsums(y, sums)
        else:
            linear_fwd_(x, w, b, y)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        ctx.set_materialize_grads(False)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        if dy is None:
            return None, None, None, None
        dy = _f32(dy, "grad")
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            linear_bwd_input_(dy, w, dx)
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or need_b:
            dw = torch.empty_like(w)
            db = torch.empty(w.shape[0], dtype=torch.float32, device=w.device) if need_b else None
            linear_bwd_weight_(dy, x, dw, db)
        return dx, dw, db, None


def linear(x, w, b=None, want_stats=False):
    return LinearFn.apply(x, w, b, want_stats)


class LinearL2NormFn(Function):
    """F.normalize(x W^T + b, p=2, dim=1, eps)  -- the last patient-MLP linear and the row normalisation that follows it
    (model.py:101-105,232) in ONE launch: the tcgen05 kernel's epilogue holds whole rows (N = 128), computes their norms from
    the TMEM accumulator and stores the normalised rows; the un-normalised product never goes to HBM.  Backward = the
    normalisation's backward kernel followed by the linear's."""

    @staticmethod
    def supported(x, w) -> bool:
        lib = _lib.load()
        return w.shape[0] <= 128 and _use_tc(lib, x.shape[0], w.shape[0], x.shape[1])

    @staticmethod
    def forward(ctx, x, w, b, eps):
        lib = _lib.load()
        x, w = _f32(x, "x"), _f32(w, "weight")
        b = None if b is None else _f32(b, "bias")
        m, k = x.shape
        n = w.shape[0]
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        inv = torch.empty(m, dtype=torch.float32, device=x.device)
        cost(4 * (m * k + n * k + m * n), 2 * m * n * k)
        _run("b2g_linear_fwd_tc", lib.b2g_linear_fwd_tc_ex, x.data_ptr(), w.data_ptr(), _ptr(b), m, n, k, y.data_ptr(), None, None, 0,
             inv.data_ptr(), float(eps), _stream())
        ctx.save_for_backward(x, w, y, inv)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, y, inv = ctx.saved_tensors
        dy = _f32(dy, "grad")
        m, n = y.shape
        g = torch.empty_like(y)
        cost(12 * m * n)
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if need_b:                          # the bias gradient (column sums of g) from the same pass
            gsum = torch.empty(n, dtype=torch.float32, device=y.device)
            ws = workspace(lib.b2g_l2norm_bwd_cs_ws_bytes(n), y.device)
            _run("b2g_l2norm_bwd", lib.b2g_l2norm_bwd_cs, y.data_ptr(), dy.data_ptr(), inv.data_ptr(), m, n, g.data_ptr(), gsum.data_ptr(),
                 ws.data_ptr(), ws.numel(), _stream())
            _tag_colsum(g, gsum)
        else:
            _run("b2g_l2norm_bwd", lib.b2g_l2norm_bwd, y.data_ptr(), dy.data_ptr(), inv.data_ptr(), m, n, g.data_ptr(), _stream())
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            linear_bwd_input_(g, w, dx)
        if ctx.needs_input_grad[1] or need_b:
            dw = torch.empty_like(w)
            db = torch.empty(w.shape[0], dtype=torch.float32, device=w.device) if need_b else None
            linear_bwd_weight_(g, x, dw, db)
        return dx, dw, db, None


class GroupedLinearFn(Function):
    """[x_i W_i^T + b_i for i in range(n)] for n independent small problems (type-node rows) in one launch each way
    (forward; input gradients; weight gradients; bias column sums).  Same results as n LinearFn calls."""

    @staticmethod
    def forward(ctx, n, *tensors):
        xs = [_f32(t, "x") for t in tensors[:n]]
        ws = [_f32(t, "weight") for t in tensors[n:2 * n]]
        bs = [None if t is None else _f32(t, "bias") for t in tensors[2 * n:3 * n]]
        ys = [torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device) for x, w in zip(xs, ws)]
        small_gemm_group_([dict(c=y, a=x, b=w, b_nk=True, bias=b) for y, x, w, b in zip(ys, xs, ws, bs)])
        ctx.save_for_backward(*xs, *ws)
        ctx.n, ctx.has_bias = n, [b is not None for b in bs]
        ctx.set_materialize_grads(False)       # an unused output must stay "no gradient" (SURVEY.md note N8), not zeros
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        n = ctx.n
        saved = ctx.saved_tensors
        xs, ws = saved[:n], saved[n:2 * n]
        nig = ctx.needs_input_grad
        dxs, dws, dbs = [None] * n, [None] * n, [None] * n
        dgrad, wgrad, cs_x, cs_o = [], [], [], []
        for i in range(n):
            if dys[i] is None:
                continue
            dy = _f32(dys[i], "grad")
            if nig[1 + i]:
                dxs[i] = torch.empty_like(xs[i])
                dgrad.append(dict(c=dxs[i], a=dy, b=ws[i]))                       # dx = dy W
            if nig[1 + n + i]:
                dws[i] = torch.empty_like(ws[i])
                wgrad.append(dict(c=dws[i], a=dy, b=xs[i], a_t=True))             # dW = dy^T x
            if ctx.has_bias[i] and nig[1 + 2 * n + i]:
                ready = _tagged_colsum(dy)
                if ready is not None and ready.numel() == ws[i].shape[0]:
                    dbs[i] = ready
                else:
                    dbs[i] = torch.empty(ws[i].shape[0], dtype=torch.float32, device=dy.device)
                    cs_x.append(dy)
                    cs_o.append(dbs[i])
        if dgrad:
            small_gemm_group_(dgrad)
        if wgrad:
            small_gemm_group_(wgrad)
        if cs_x:
            small_colsum_group_(cs_x, cs_o)
        return (None, *dxs, *dws, *dbs)


def linear_group(xs, ws, bs=None):
    """n small linears at once; falls back to one call each when a problem is not small."""
    n = len(xs)
    bs = list(bs) if bs is not None else [None] * n
    if n == 0:
        return []
    if n == 1 or any(x.shape[0] > SMALL_ROWS for x in xs):
        return [linear(x, w, b) for x, w, b in zip(xs, ws, bs)]
    return list(GroupedLinearFn.apply(n, *xs, *ws, *bs))


class SageTypeDstGroupFn(Function):
    """SageDstFn for several destination node types with few rows (labs, diagnoses, medications), each with one aggregated
    relation:  out_t = x_t W_root,t^T + b_root,t + agg_t W_l,t^T  -- all destinations in ONE launch (the two products of a
    destination share its output tile), and three launches backward (input gradients, weight gradients, bias sums)."""

    @staticmethod
    def forward(ctx, n, *tensors):
        xs = [_f32(t, "x_dst") for t in tensors[:n]]
        wr = [_f32(t, "W_root") for t in tensors[n:2 * n]]
        br = [None if t is None else _f32(t, "b_root") for t in tensors[2 * n:3 * n]]
        ag = [_f32(t, "agg") for t in tensors[3 * n:4 * n]]
        wl = [_f32(t, "W_l") for t in tensors[4 * n:5 * n]]
        outs = [torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device) for x, w in zip(xs, wr)]
        small_gemm_group_([dict(c=o, a=x, b=w, b_nk=True, a2=a, b2=l, bias=b) for o, x, w, b, a, l in zip(outs, xs, wr, br, ag, wl)])
        ctx.save_for_backward(*xs, *wr, *ag, *wl)
        ctx.n, ctx.has_bias = n, [b is not None for b in br]
        ctx.set_materialize_grads(False)       # an unused destination (note N8) must yield None gradients, not zeros
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        n = ctx.n
        sv = ctx.saved_tensors
        xs, wr, ag, wl = sv[:n], sv[n:2 * n], sv[2 * n:3 * n], sv[3 * n:4 * n]
        nig = ctx.needs_input_grad
        dxs, dwr, dbr, dag, dwl = ([None] * n for _ in range(5))
        dgrad, wgrad, cs_x, cs_o = [], [], [], []
        for i in range(n):
            if douts[i] is None:          # this destination's output is unused (SURVEY.md note N8): no gradient at all
                continue
            do = _f32(douts[i], "grad")
            if nig[1 + i]:
                dxs[i] = torch.empty_like(xs[i])
                dgrad.append(dict(c=dxs[i], a=do, b=wr[i]))
            if nig[1 + 3 * n + i]:
                dag[i] = torch.empty_like(ag[i])
                dgrad.append(dict(c=dag[i], a=do, b=wl[i]))
            if nig[1 + n + i]:
                dwr[i] = torch.empty_like(wr[i])
                wgrad.append(dict(c=dwr[i], a=do, b=xs[i], a_t=True))
            if nig[1 + 4 * n + i]:
                dwl[i] = torch.empty_like(wl[i])
                wgrad.append(dict(c=dwl[i], a=do, b=ag[i], a_t=True))
            if ctx.has_bias[i] and nig[1 + 2 * n + i]:
                ready = _tagged_colsum(do)
                if ready is not None and ready.numel() == wr[i].shape[0]:
                    dbr[i] = ready
                else:
                    dbr[i] = torch.empty(wr[i].shape[0], dtype=torch.float32, device=do.device)
                    cs_x.append(do)
                    cs_o.append(dbr[i])
        if dgrad:
            small_gemm_group_(dgrad)
        if wgrad:
            small_gemm_group_(wgrad)
        if cs_x:
            small_colsum_group_(cs_x, cs_o)
        return (None, *dxs, *dwr, *dbr, *dag, *dwl)


class BNActDropFn(Function):
    """dropout(act(batch_norm(x)))  -- nn.BatchNorm1d -> activation -> F.dropout (model.py:93-101,259-269).
    Training mode uses batch statistics and updates the running buffers in place (momentum 0.1, unbiased
    variance); eval mode uses the running buffers.  The dropout mask is regenerated from (seed, sid)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, act, p_drop, seed, sid, eps, momentum):
        lib = _lib.load()
        x, gamma, beta = _f32(x, "x"), _f32(gamma, "bn.weight"), _f32(beta, "bn.bias")
        m, d = x.shape
        dev = x.device
        mean = torch.empty(d, dtype=torch.float32, device=dev)
        rstd = torch.empty(d, dtype=torch.float32, device=dev)
        sums = _tagged_bnsums(x) if training else None
        if training and sums is not None:        # the producing kernel's epilogue already reduced the columns
            _run("b2g_bn_finalize_sums", lib.b2g_bn_finalize_sums, sums.data_ptr(), m, d, float(eps), float(momentum), mean.data_ptr(),
                 rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), _stream())
        elif training:
            ws = workspace(lib.b2g_bn_ws_bytes(d), dev)
            cost(4 * m * d)
            _run("b2g_bn_stats", lib.b2g_bn_stats, x.data_ptr(), m, d, float(eps), float(momentum), mean.data_ptr(), rstd.data_ptr(),
                                        _ptr(running_mean), _ptr(running_var), ws.data_ptr(), ws.numel(), _stream())
        else:
            _run("b2g_bn_eval_stats", lib.b2g_bn_eval_stats, running_mean.data_ptr(), running_var.data_ptr(), d, float(eps), mean.data_ptr(),
                                             rstd.data_ptr(), _stream())
        p = float(p_drop) if training else 0.0
        y = torch.empty_like(x)
        cost(8 * m * d)
        _run("b2g_bn_apply", lib.b2g_bn_apply, x.data_ptr(), m, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                    int(act), p, int(seed), int(sid), y.data_ptr(), _stream())
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        ctx.cfg = (int(act), p, int(seed), int(sid), bool(training))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        act, p, seed, sid, training = ctx.cfg
        dy = _f32(dy, "grad")
        m, d = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(beta)
        colsum = torch.empty(d, dtype=torch.float32, device=x.device)
        ws = workspace(lib.b2g_bn_ws_bytes(d), x.device)
        cost(12 * m * d)
        _run("b2g_bn_bwd", lib.b2g_bn_bwd, x.data_ptr(), dy.data_ptr(), m, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                  beta.data_ptr(), act, p, seed, sid, int(training), dx.data_ptr(), dgamma.data_ptr(),
                                  dbeta.data_ptr(), colsum.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        _tag_colsum(dx, colsum)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class SyncBNActDropFn(Function):
    """Training-mode BNActDropFn for a node type whose rows are partitioned over ranks (patients): the fp64 column
    totals are all-reduced so that every rank normalises with the statistics of ALL rows (= the single-GPU result);
    the backward all-reduces {sum g, sum g*xhat} the same way.  gamma/beta gradients are the global totals on every
    rank, so they are returned divided by the world size (the step's gradient all-reduce sums them back)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, act, p_drop, seed, sid, eps, momentum, dctx, m_total):
        lib = _lib.load()
        x, gamma, beta = _f32(x, "x"), _f32(gamma, "bn.weight"), _f32(beta, "bn.bias")
        m, d = x.shape
        dev = x.device
        ws = workspace(lib.b2g_bn_ws_bytes(d), dev)
        mean = torch.empty(d, dtype=torch.float32, device=dev)
        rstd = torch.empty(d, dtype=torch.float32, device=dev)
        peer = getattr(dctx, "peer", None)
        cost(4 * m * d)
        tagged = _tagged_bnsums(x)
        if tagged is not None:    # local totals came out of the producing kernel's epilogue: exchange 2 KB, finalise
            sums = tagged.clone()
            dctx.all_reduce_(sums)
            _run("b2g_bn_finalize_sums", lib.b2g_bn_finalize_sums, sums.data_ptr(), int(m_total), d, float(eps), float(momentum),
                 mean.data_ptr(), rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), _stream())
        elif peer is not None:      # statistics + NVLink exchange + finalisation in ONE kernel (csrc/peer.cuh)
            _run("b2g_bn_stats_sync", lib.b2g_bn_stats_sync, peer.handle, x.data_ptr(), m, int(m_total), d, float(eps), float(momentum),
                 mean.data_ptr(), rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), ws.data_ptr(), ws.numel(), _stream())
            dctx.count_fused()
        else:
            sums = torch.empty(2 * d, dtype=torch.float64, device=dev)
            _run("b2g_bn_local_sums", lib.b2g_bn_local_sums, x.data_ptr(), m, d, sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            dctx.all_reduce_(sums)
            _run("b2g_bn_finalize_sums", lib.b2g_bn_finalize_sums, sums.data_ptr(), int(m_total), d, float(eps), float(momentum),
                 mean.data_ptr(), rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), _stream())
        y = torch.empty_like(x)
        cost(8 * m * d)
        _run("b2g_bn_apply", lib.b2g_bn_apply, x.data_ptr(), m, d, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
             int(act), float(p_drop), int(seed), int(sid), y.data_ptr(), _stream())
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        ctx.cfg = (int(act), float(p_drop), int(seed), int(sid), dctx, int(m_total))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        act, p, seed, sid, dctx, m_total = ctx.cfg
        dy = _f32(dy, "grad")
        m, d = x.shape
        dev = x.device
        ws = workspace(lib.b2g_bn_ws_bytes(d), dev)
        dx = torch.empty_like(x)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
        inv = 1.0 / dctx.world
        peer = getattr(dctx, "peer", None)
        if peer is not None:      # backward statistics + NVLink exchange fused into the reduction kernel, then dx
            colsum = torch.empty(d, dtype=torch.float32, device=dev)
            cost(12 * m * d)
            _run("b2g_bn_bwd_sync", lib.b2g_bn_bwd_sync, peer.handle, x.data_ptr(), dy.data_ptr(), m, m_total, d, mean.data_ptr(),
                 rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, p, seed, sid, dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                 colsum.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            _tag_colsum(dx, colsum)
            dctx.count_fused()
            return dx, dgamma * inv, dbeta * inv, None, None, None, None, None, None, None, None, None, None
        sums = torch.empty(2 * d, dtype=torch.float64, device=dev)
        cost(8 * m * d)
        _run("b2g_bn_bwd_local_sums", lib.b2g_bn_bwd_local_sums, x.data_ptr(), dy.data_ptr(), m, d, mean.data_ptr(), rstd.data_ptr(),
             gamma.data_ptr(), beta.data_ptr(), act, p, seed, sid, sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        dctx.all_reduce_(sums)
        cost(12 * m * d)
        _run("b2g_bn_bwd_from_sums", lib.b2g_bn_bwd_from_sums, x.data_ptr(), dy.data_ptr(), m, m_total, d, mean.data_ptr(), rstd.data_ptr(),
             gamma.data_ptr(), beta.data_ptr(), act, p, seed, sid, sums.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
             _stream())
        return dx, dgamma * inv, dbeta * inv, None, None, None, None, None, None, None, None, None, None


class ActDropFn(Function):
    """dropout(act(x)) without normalisation (use_batch_norm=False branch of model.py:259-269)."""

    @staticmethod
    def forward(ctx, x, act, p_drop, seed, sid, training):
        lib = _lib.load()
        x = _f32(x, "x")
        m, d = x.shape
        dev = x.device
        zeros = torch.zeros(d, dtype=torch.float32, device=dev)
        ones = torch.ones(d, dtype=torch.float32, device=dev)
        p = float(p_drop) if training else 0.0
        y = torch.empty_like(x)
        _run("b2g_bn_apply", lib.b2g_bn_apply, x.data_ptr(), m, d, zeros.data_ptr(), ones.data_ptr(), ones.data_ptr(), zeros.data_ptr(),
                                    int(act), p, int(seed), int(sid), y.data_ptr(), _stream())
        ctx.save_for_backward(x, zeros, ones)
        ctx.cfg = (int(act), p, int(seed), int(sid))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, zeros, ones = ctx.saved_tensors
        act, p, seed, sid = ctx.cfg
        dy = _f32(dy, "grad")
        m, d = x.shape
        dx = torch.empty_like(x)
        ws = workspace(lib.b2g_bn_ws_bytes(d), x.device)
        _run("b2g_bn_bwd", lib.b2g_bn_bwd, x.data_ptr(), dy.data_ptr(), m, d, zeros.data_ptr(), ones.data_ptr(), ones.data_ptr(),
                                  zeros.data_ptr(), act, p, seed, sid, 0, dx.data_ptr(), None, None, None, ws.data_ptr(),
                                  ws.numel(), _stream())
        return dx, None, None, None, None, None


class ReluDropoutFn(Function):
    """dropout(relu(x)) on a flat tensor (EdgeRegressionHead: nn.ReLU -> nn.Dropout, model.py:377-380)."""

    @staticmethod
    def forward(ctx, x, relu, p_drop, seed, sid, training):
        lib = _lib.load()
        x = _f32(x, "x")
        p = float(p_drop) if training else 0.0
        y = torch.empty_like(x)
        cost(8 * x.numel())
        _run("b2g_relu_dropout_fwd", lib.b2g_relu_dropout_fwd, x.data_ptr(), x.numel(), int(relu), p, int(seed), int(sid), y.data_ptr(), _stream())
        ctx.save_for_backward(y)
        ctx.cfg = (int(relu), p, int(seed), int(sid))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        relu, p, seed, sid = ctx.cfg
        dy = _f32(dy, "grad")
        dx = torch.empty_like(y)
        cost(12 * y.numel())
        _run("b2g_relu_dropout_bwd", lib.b2g_relu_dropout_bwd, y.data_ptr(), dy.data_ptr(), y.numel(), relu, p, seed, sid, dx.data_ptr(), _stream())
        return dx, None, None, None, None, None


class L2NormFn(Function):
    """F.normalize(x, p=2, dim=1, eps=1e-12)  (model.py:105,232)."""

    @staticmethod
    def forward(ctx, x, eps):
        lib = _lib.load()
        x = _f32(x, "x")
        m, d = x.shape
        y = torch.empty_like(x)
        inv = torch.empty(m, dtype=torch.float32, device=x.device)
        cost(8 * m * d)
        _run("b2g_l2norm_fwd", lib.b2g_l2norm_fwd, x.data_ptr(), m, d, float(eps), y.data_ptr(), inv.data_ptr(), _stream())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        y, inv = ctx.saved_tensors
        dy = _f32(dy, "grad")
        m, d = y.shape
        dx = torch.empty_like(y)
        cost(12 * m * d)
        _run("b2g_l2norm_bwd", lib.b2g_l2norm_bwd, y.data_ptr(), dy.data_ptr(), inv.data_ptr(), m, d, dx.data_ptr(), _stream())
        return dx, None


class MeanAggFn(Function):
    """agg[v] = mean_{u -> v} x_src[u]  with count clamped to >= 1  (PyG SAGEConv(aggr='mean').propagate,
    called from model.py:256).  Backward walks the transposed CSR: no float atomics.  In tf32 mode dense-enough
    relations use the tensor-core formulation on the dense adjacency (see b2g.h)."""

    @staticmethod
    def forward(ctx, x_src, rel: Relation):
        x_src = _f32(x_src, "x_src")
        dn = _dense_of(rel, x_src.shape[1])
        ctx.rel, ctx.dn = rel, dn
        ctx.set_materialize_grads(False)                  # an aggregate nobody differentiates (note N8) costs no backward work
        if dn is not None and not dn.big_is_dst:          # many sources -> few destinations: diag(1/deg_dst) D^T x_src
            return _row_scale(_adjT_times_rows(dn, x_src), rel.by_dst.inv_deg)
        out = torch.empty((rel.n_dst, x_src.shape[1]), dtype=torch.float32, device=x_src.device)
        if dn is not None:                                # few sources -> many destinations: D' x_src
            return _adj_times_table(dn, x_src, None, out, False)
        gather_reduce_([rel.by_dst], [x_src], [rel.by_dst.inv_deg], [None], out, False)
        return out

    @staticmethod
    def backward(ctx, dout):
        rel, dn = ctx.rel, ctx.dn
        if dout is None:
            return None, None
        dout = _f32(dout, "grad")
        if dn is not None and not dn.big_is_dst:          # dx_src = D (dout / deg_dst)
            dx = torch.empty((rel.n_src, dout.shape[1]), dtype=torch.float32, device=dout.device)
            return _adj_times_table(dn, dout, rel.by_dst.inv_deg, dx, False), None
        if dn is not None:                                # dx_src = D'^T dout
            return _adjT_times_rows(dn, dout), None
        dx = torch.empty((rel.n_src, dout.shape[1]), dtype=torch.float32, device=dout.device)
        gather_reduce_([rel.by_src], [dout], [None], [rel.by_dst.inv_deg], dx, False)
        return dx, None


class SageDstFn(Function):
    """All SAGEConv results that land on one destination node type, summed like HeteroConv(aggr='sum')
    (PyG hetero_conv.py group(); model.py:125-131,256):

        out = x_dst W_root^T + b_root                        (sum over relations of lin_r, and of lin_l's bias)
              + sum_{small-source rels} mean_rel(Y_rel)      (Y_rel = x_src W_l^T computed on the few source rows first)
              + sum_{big-source rels}  agg_rel W_l_rel^T     (agg_rel = mean_rel(x_src) computed first)

    The re-association (A X) W = A (X W) keeps the only large-M GEMM at K = d.
    """

    @staticmethod
    def forward(ctx, x_dst, w_root, b_root, small_rels, n_big, *tensors):
        n_small = len(small_rels)
        ys = [_f32(t, "Y") for t in tensors[:n_small]]
        aggs = [_f32(t, "agg") for t in tensors[n_small:n_small + n_big]]
        wls = [_f32(t, "W_l") for t in tensors[n_small + n_big:n_small + 2 * n_big]]
        x_dst, w_root = _f32(x_dst, "x_dst"), _f32(w_root, "W_root")
        b_root = None if b_root is None else _f32(b_root, "b_root")
        out = torch.empty((x_dst.shape[0], w_root.shape[0]), dtype=torch.float32, device=x_dst.device)
        linear_fwd_(x_dst, w_root, b_root, out)
        for agg, wl in zip(aggs, wls):
            linear_fwd_(agg, wl, None, out, accumulate=True)
        d = out.shape[1]
        dense = [_dense_of(r, d) for r in small_rels]
        dense = [dn if (dn is not None and dn.big_is_dst) else None for dn in dense]
        for r, y, dn in zip(small_rels, ys, dense):
            if dn is not None:
                _adj_times_table(dn, y, None, out, True)
        rest = [i for i, dn in enumerate(dense) if dn is None]
        if rest:
            gather_reduce_([small_rels[i].by_dst for i in rest], [ys[i] for i in rest], [small_rels[i].by_dst.inv_deg for i in rest],
                           [None] * len(rest), out, True)
        ctx.dense = dense
        ctx.save_for_backward(x_dst, w_root, *aggs, *wls)
        ctx.meta = (small_rels, n_big, b_root is not None, [y.shape for y in ys])
        return out

    @staticmethod
    def backward(ctx, dout):
        small_rels, n_big, has_bias, y_shapes = ctx.meta
        saved = ctx.saved_tensors
        x_dst, w_root = saved[0], saved[1]
        aggs, wls = saved[2:2 + n_big], saved[2 + n_big:2 + 2 * n_big]
        dout = _f32(dout, "grad")
        n_small = len(small_rels)
        nig = ctx.needs_input_grad
        dx = dw = db = None
        if nig[0]:
            dx = torch.empty_like(x_dst)
            linear_bwd_input_(dout, w_root, dx)
        need_b = has_bias and nig[2]
        if nig[1] or need_b:
            dw = torch.empty_like(w_root)
            db = torch.empty(w_root.shape[0], dtype=torch.float32, device=dout.device) if need_b else None
            linear_bwd_weight_(dout, x_dst, dw, db)
        d_ys: List[Optional[torch.Tensor]] = []
        for k, rel in enumerate(small_rels):
            if nig[5 + k]:
                if ctx.dense[k] is not None:
                    d_ys.append(_adjT_times_rows(ctx.dense[k], dout))
                    continue
                dy = torch.empty(y_shapes[k], dtype=torch.float32, device=dout.device)
                gather_reduce_([rel.by_src], [dout], [None], [rel.by_dst.inv_deg], dy, False)
                d_ys.append(dy)
            else:
                d_ys.append(None)
        d_aggs, d_wls = [], []
        for k in range(n_big):
            if nig[5 + n_small + k]:
                da = torch.empty_like(aggs[k])
                linear_bwd_input_(dout, wls[k], da)
                d_aggs.append(da)
            else:
                d_aggs.append(None)
            if nig[5 + n_small + n_big + k]:
                dwl = torch.empty_like(wls[k])
                linear_bwd_weight_(dout, aggs[k], dwl, None)
                d_wls.append(dwl)
            else:
                d_wls.append(None)
        return (dx, dw, db, None, None, *d_ys, *d_aggs, *d_wls)


# ---- single-launch patient side of a HeteroConv layer (csrc/layer_tc.cu), tf32 mode ----------------------------------------
def _ptr_array(ts):
    return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


def layer_cat_weights_(ws_list, transposed, tabs, scales, offs, kx, ktot, n, biases=()):
    """([sum of W (or its transpose) | tab_i^T * scale_i at column kx + offs[i]] -> [n, ktot],  sum of the biases or None)"""
    lib = _lib.load()
    dev = ws_list[0].device if ws_list else next(t for t in tabs if t is not None).device
    out = torch.empty((n, ktot), dtype=torch.float32, device=dev)
    biases = [b for b in biases if b is not None]
    bias_out = torch.empty(n, dtype=torch.float32, device=dev) if biases else None
    live = [(t, sc, o) for t, sc, o in zip(tabs, scales, offs) if t is not None]
    nr = len(live)
    pt = _ptr_array([t for t, _, _ in live] or [None])
    ps = _ptr_array([sc for _, sc, _ in live] or [None])
    rows = (ctypes.c_int * max(nr, 1))(*[int(t.shape[0]) for t, _, _ in live] or [0])
    off = (ctypes.c_int * max(nr, 1))(*[int(o) for _, _, o in live] or [0])
    _run("b2g_layer_cat_weights", lib.b2g_layer_cat_weights, _ptr_array(ws_list or [None]), len(ws_list), int(transposed),
         _ptr_array(biases or [None]), len(biases), _ptr(bias_out), n, kx, pt, ps, rows, off, nr, ktot, out.data_ptr(), _stream())
    return out, bias_out


LAYER_HALF = os.environ.get("B2G_LAYER_HALF", "1") != "0"      # adjacency part of k_layer_tf32 as fp16 {0 | 1/deg} (kind::f16)


def layer_cat_half_(wcat, kx):
    """(whalf [n, 64 ceil(nw/2)] fp16, unscale [n]) of b2g_layer_cat_half; scales wcat[:, :kx] in place"""
    lib = _lib.load()
    n, ktot = wcat.shape
    kh = 64 * (((ktot - kx) // 32 + 1) // 2)
    whalf = torch.empty((n, kh), dtype=torch.float16, device=wcat.device)
    unscale = torch.empty(n, dtype=torch.float32, device=wcat.device)
    _run("b2g_layer_cat_half", lib.b2g_layer_cat_half, wcat.data_ptr(), n, kx, ktot, whalf.data_ptr(), unscale.data_ptr(), _stream())
    return whalf, unscale


def layer_fwd_tc_(x, wcat, bias, bits, pb, rscales, y, stat_sums=None, half=None):
    """half: None -> TF32 adjacency tiles; (whalf, unscale) from layer_cat_half_ -> fp16 adjacency tiles"""
    lib = _lib.load()
    m, kx = x.shape
    n = wcat.shape[0]
    ws = None
    if stat_sums is not None:
        ws = workspace(lib.b2g_layer_stats_ws_bytes(n), x.device)
    whalf, unscale = half if half is not None else (None, None)
    cost(4 * (m * kx + m * n) + 4 * m * pb.nw + 4 * wcat.numel(), 2 * m * n * wcat.shape[1])
    _run("b2g_layer_fwd_tc", lib.b2g_layer_fwd_tc, x.data_ptr(), wcat.data_ptr(), _ptr(whalf), _ptr(unscale), _ptr(bias), bits.data_ptr(), ctypes.byref(pb.layout),
         _ptr_array(list(rscales) + [None] * (4 - len(rscales))), m, n, kx, y.data_ptr(), _ptr(stat_sums), _ptr(ws),
         0 if ws is None else ws.numel(), _stream())
    return y


def layer_adjT_tc_(x, bits, pb, rscales, col_scale, with_colsum=False, dense_b=None):
    """[32 nw, 128] = col_scale * (diag(rscale) A)^T x;  with_colsum: 32 more rows, row 32 nw = column sums of x;
    dense_b [m, 128]: additionally returns x^T dense_b [128, 128] from the same pass (-> (out, x^T dense_b))"""
    lib = _lib.load()
    m = x.shape[0]
    nsub = pb.nw + (1 if with_colsum else 0)
    out = torch.empty((32 * nsub, 128), dtype=torch.float32, device=x.device)
    dw = torch.empty((128, 128), dtype=torch.float32, device=x.device) if dense_b is not None else None
    ws = workspace(lib.b2g_layer_adjT_tc_ws_bytes(nsub), x.device)
    cost(4 * m * 128 * (2 if dense_b is not None else 1) + 4 * m * pb.nw + 4 * out.numel(), 2 * m * 128 * (32 * nsub + (128 if dense_b is not None else 0)))
    _run("b2g_layer_adjT_tc", lib.b2g_layer_adjT_tc, x.data_ptr(), bits.data_ptr(), ctypes.byref(pb.layout),
         _ptr_array(list(rscales) + [None] * (4 - len(rscales))), _ptr(col_scale), m, int(bool(with_colsum)), _ptr(dense_b), _ptr(dw),
         out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return out if dense_b is None else (out, dw)


def adjT_columns_fit(pb, with_colsum: bool, with_dense: bool) -> bool:
    return 32 * (pb.nw + (1 if with_colsum else 0)) + (128 if with_dense else 0) <= 512


def patient_side_supported(pb, n_rows: int, d: int) -> bool:
    if PRECISION != "tf32" or pb is None or d != 128 or n_rows < TC_MIN_ROWS:
        return False
    lib = _lib.load()
    return bool(lib.b2g_layer_fwd_tc_supported(n_rows, d, d, pb.nw)) and bool(lib.b2g_layer_adjT_tc_supported(n_rows, d, pb.nw))


class PatientSideFn(Function):
    """Everything a HeteroConv layer does on the PATIENT rows (model.py:125-131,256; PyG HeteroConv(aggr='sum') of
    SAGEConv(mean)), two launches forward and three backward, with the adjacency as a bit matrix (graph.PatientBits):

        out_p   = x_p (sum_r W_root,r)^T + sum_r b_r + sum_t diag(1/deg_t(p)) A_t Y_t      Y_t = x_t W_l^T, few rows, given
        agg_t   = diag(1/deg(t)) A_t^T x_p          for every type t with a relation patient -> t

    inputs: pb, n_w, n_t, x_p, W_root_1..n_w, b_1..n_w, Y_t for t in pb.types (None: no relation t -> patient)
    outputs: out_p, agg_t for t in pb.types (a zero-row tensor when there is no relation patient -> t)."""

    @staticmethod
    def forward(ctx, pb, n_w, want_stats, x_p, *tensors):
        nt = len(pb.types)
        w_roots = [_f32(t, "W_root") for t in tensors[:n_w]]
        biases = [None if t is None else _f32(t, "b_root") for t in tensors[n_w:2 * n_w]]
        ys = [None if t is None else _f32(t, "Y") for t in tensors[2 * n_w:2 * n_w + nt]]
        x_p = _f32(x_p, "x_patient")
        m, d = x_p.shape
        ktot = d + 32 * pb.nw
        ys_in = [y if r is not None else None for y, r in zip(ys, pb.in_rel)]
        wcat, bias = layer_cat_weights_(w_roots, False, ys_in, [None] * nt, pb.offs, d, ktot, d, biases)
        out = torch.empty((m, d), dtype=torch.float32, device=x_p.device)
        if pb.bits_in is not None and any(y is not None for y in ys_in):
            sums = torch.empty(2 * d, dtype=torch.float64, device=x_p.device) if (want_stats and d <= 128) else None
            half = layer_cat_half_(wcat, d) if LAYER_HALF else None
            layer_fwd_tc_(x_p, wcat, bias, pb.bits_in, pb, pb.rscale_in(), out, sums, half)
            if sums is not None:
                _tag_bnsums(out, sums)          # BatchNorm statistics of the layer output (model.py:259-261) for free
        else:
            linear_fwd_(x_p, wcat[:, :d].contiguous(), bias, out)
        aggs = []
        if pb.bits_out is not None:
            t_all = layer_adjT_tc_(x_p, pb.bits_out, pb, [None] * nt, pb.col_scale_out())
            for r, off, n in zip(pb.out_rel, pb.offs, pb.sizes):
                aggs.append(t_all[off:off + n] if r is not None else t_all[0:0])
        else:
            aggs = [out.new_zeros((0, d)) for _ in pb.types]
        ctx.save_for_backward(x_p, *w_roots)
        ctx.meta = (pb, n_w, [b is not None for b in biases], [None if y is None else tuple(y.shape) for y in ys])
        ctx.set_materialize_grads(False)
        return (out, *aggs)

    @staticmethod
    def backward(ctx, dout, *daggs):
        pb, n_w, has_bias, y_shapes = ctx.meta
        saved = ctx.saved_tensors
        x_p, w_roots = saved[0], list(saved[1:1 + n_w])
        nt = len(pb.types)
        m, d = x_p.shape
        nig = ctx.needs_input_grad[1:]        # (n_w, want_stats | x_p, W.., b.., Y..) -> indices below count from n_w = 0
        dev = x_p.device
        dx = None
        dws = [None] * n_w
        dbs = [None] * n_w
        dys = [None] * nt
        live_aggs = [(None if g is None or r is None else _f32(g, "grad")) for g, r in zip(daggs, pb.out_rel)]
        if dout is not None:
            dout = _f32(dout, "grad")
        if nig[2]:
            ktot = d + 32 * pb.nw
            if dout is None:
                dout_x = torch.zeros_like(x_p)
            else:
                dout_x = dout
            if pb.bits_out is not None and any(g is not None for g in live_aggs):
                wcat, _ = layer_cat_weights_(w_roots, True, live_aggs, pb.inv_deg_out(), pb.offs, d, ktot, d)
                dx = torch.empty_like(x_p)
                half = layer_cat_half_(wcat, d) if LAYER_HALF else None
                layer_fwd_tc_(dout_x, wcat, None, pb.bits_out, pb, [None] * nt, dx, None, half)
            elif dout is not None:
                w = w_roots[0]
                for extra in w_roots[1:]:
                    w = w + extra
                dx = torch.empty_like(x_p)
                linear_bwd_input_(dout, w, dx)
        if dout is not None:
            need_w = any(nig[3 + i] for i in range(n_w))
            need_b = any(has_bias[i] and nig[3 + n_w + i] for i in range(n_w))
            want_y = [y_shapes[i] is not None and pb.in_rel[i] is not None and nig[3 + 2 * n_w + i] for i in range(nt)]
            db = _tagged_colsum(dout) if need_b else None          # the producer of dout may have reduced its columns already
            dw = None
            if any(want_y) and pb.bits_in is not None:
                # ONE pass over dout: dY_t = (diag(1/deg_t(p)) A_t)^T dout, the bias gradient (column sums of dout) as one more
                # column of ones, and dW_root = dout^T x_p as 128 more columns (when the 512 TMEM columns hold all of it)
                fuse_b = need_b and db is None and adjT_columns_fit(pb, True, False)
                fuse_w = need_w and adjT_columns_fit(pb, fuse_b, True)
                res = layer_adjT_tc_(dout, pb.bits_in, pb, pb.rscale_in(), None, with_colsum=fuse_b, dense_b=x_p if fuse_w else None)
                dy_all, dw = res if fuse_w else (res, None)
                if fuse_b:
                    db = dy_all[32 * pb.nw]
                for i, (off, n) in enumerate(zip(pb.offs, pb.sizes)):
                    if want_y[i]:
                        dys[i] = dy_all[off:off + n]
            if dw is None and need_w:                   # separate weight-gradient launch (and the column sums, if still missing)
                dw = torch.empty((d, d), dtype=torch.float32, device=dev)
                db_new = torch.empty(d, dtype=torch.float32, device=dev) if (need_b and db is None) else None
                linear_bwd_weight_(dout, x_p, dw, db_new)
                db = db if db_new is None else db_new
            if need_b and db is None:
                lib = _lib.load()
                db = torch.empty(d, dtype=torch.float32, device=dev)
                ws = workspace(lib.b2g_bn_ws_bytes(d), dev)
                _run("b2g_col_sums", lib.b2g_col_sums, dout.data_ptr(), m, d, db.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
            if dw is not None:
                for i in range(n_w):
                    if nig[3 + i]:
                        dws[i] = dw
            if need_b:
                for i in range(n_w):
                    if has_bias[i] and nig[3 + n_w + i]:
                        dbs[i] = db
        return (None, None, None, dx, *dws, *dbs, *dys)


class PairAddReluFn(Function):
    """z[i] = relu(U[p_i] + V[l_i]) -- the first decoder layer after factorising
    Linear(2d, 64)(cat[h_p, h_l]) = h_p W[:, :d]^T + (h_l W[:, d:]^T + b)   (model.py:305-309,324-333,373-377):
    the per-pair [M, 2d] concatenation and its index_put backward are never materialised."""

    @staticmethod
    def forward(ctx, u, v, pairs: PairIndex):
        lib = _lib.load()
        u, v = _f32(u, "U"), _f32(v, "V")
        d = u.shape[1]
        z = torch.empty((pairs.m, d), dtype=torch.float32, device=u.device)
        cost(4 * pairs.m * d + 16 * pairs.m + 4 * (u.numel() + v.numel()))
        _run("b2g_gather_add_rows", lib.b2g_gather_add_rows, u.data_ptr(), pairs.patient_idx.data_ptr(), v.data_ptr(), pairs.lab_idx.data_ptr(),
                                           pairs.m, d, 1, z.data_ptr(), _stream())
        ctx.save_for_backward(z)
        ctx.meta = (pairs, u.shape, v.shape)
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = _lib.load()
        (z,) = ctx.saved_tensors
        pairs, u_shape, v_shape = ctx.meta
        dz = _f32(dz, "grad")
        g = torch.empty_like(z)
        cost(12 * z.numel())
        _run("b2g_relu_dropout_bwd", lib.b2g_relu_dropout_bwd, z.data_ptr(), dz.data_ptr(), z.numel(), 1, 0.0, 0, 0, g.data_ptr(), _stream())
        du = dv = None
        if ctx.needs_input_grad[0]:
            du = torch.empty(u_shape, dtype=torch.float32, device=dz.device)
            gather_reduce_([pairs.by_patient], [g], [None], [None], du, False)
        if ctx.needs_input_grad[1]:
            dv = torch.empty(v_shape, dtype=torch.float32, device=dz.device)
            gather_reduce_([pairs.by_lab], [g], [None], [None], dv, False)
        return du, dv, None


class DecoderHeadFn(Function):
    """EdgeRegressionHead([64, 32]) over a pair list in one kernel each way (csrc/decoder.cu):
    pred = w3 . drop(relu(W2 drop(relu(U[p] + V[l])) + b2)) + b3   (model.py:305-333,373-386)."""

    @staticmethod
    def forward(ctx, u, v, w2, b2, w3, b3, pairs: PairIndex, p_drop, seed, sid1, sid2, training):
        lib = _lib.load()
        u, v, w2, b2, w3, b3 = (_f32(t, n) for t, n in ((u, "U"), (v, "V"), (w2, "W2"), (b2, "b2"), (w3, "w3"), (b3, "b3")))
        if u.shape[1] != 64 or w2.shape != (32, 64) or w3.numel() != 32:
            raise _lib.B2GError("fused decoder supports hidden_dims [64, 32] only")
        p = float(p_drop) if training else 0.0
        pred = torch.empty(pairs.m, dtype=torch.float32, device=u.device)
        cost(pairs.m * (16 + 4 + 256) + 4 * v.numel(), 2 * pairs.m * (64 * 32 + 32 + 64))
        tc = PRECISION == "tf32" and pairs.m >= TC_MIN_ROWS
        _run("b2g_decoder_fwd_tc" if tc else "b2g_decoder_fwd", lib.b2g_decoder_fwd_tc if tc else lib.b2g_decoder_fwd, u.data_ptr(),
             v.data_ptr(), pairs.patient_idx.data_ptr(), pairs.lab_idx.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(),
             b3.data_ptr(), pairs.m, p, int(seed), int(sid1), int(sid2), pred.data_ptr(), _stream())
        ctx.save_for_backward(u, v, w2, b2, w3)
        ctx.meta = (pairs, p, int(seed), int(sid1), int(sid2))
        return pred

    @staticmethod
    def backward(ctx, dpred):
        lib = _lib.load()
        u, v, w2, b2, w3 = ctx.saved_tensors
        pairs, p, seed, sid1, sid2 = ctx.meta
        dpred = _f32(dpred, "grad")
        dev = u.device
        m = pairs.m
        if m == 0:      # a head without pairs on this rank: zero gradients (still returned, so collectives upstream stay matched)
            return (torch.zeros_like(u), torch.zeros_like(v), torch.zeros_like(w2), torch.zeros_like(b2), torch.zeros_like(w3),
                    torch.zeros(1, dtype=torch.float32, device=dev), None, None, None, None, None, None)
        g = torch.empty((m, 64), dtype=torch.float32, device=dev)
        flags = torch.empty(m, dtype=torch.float32, device=dev)
        dw2, db2, dw3 = torch.empty_like(w2), torch.empty_like(b2), torch.empty_like(w3)
        db3 = torch.empty(1, dtype=torch.float32, device=dev)
        ws = workspace(lib.b2g_decoder_bwd_ws_bytes(m), dev)
        cost(m * (16 + 4 + 4) + m // 5 * 512, 2 * (m // 5) * 3 * 64 * 32)
        tc = PRECISION == "tf32" and m >= TC_MIN_ROWS
        _run("b2g_decoder_bwd_tc" if tc else "b2g_decoder_bwd", lib.b2g_decoder_bwd_tc if tc else lib.b2g_decoder_bwd, u.data_ptr(),
             v.data_ptr(), pairs.patient_idx.data_ptr(), pairs.lab_idx.data_ptr(),
             w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), dpred.data_ptr(), m, p, seed, sid1, sid2, g.data_ptr(), flags.data_ptr(),
             dw2.data_ptr(), db2.data_ptr(), dw3.data_ptr(), db3.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        du = dv = None
        if ctx.needs_input_grad[0]:
            du = torch.empty_like(u)
            gather_reduce_([pairs.by_patient], [g], [None], [flags], du, False)
        if ctx.needs_input_grad[1]:
            dv = torch.empty_like(v)
            gather_reduce_([pairs.by_lab], [g], [None], [flags], dv, False)
        return du, dv, dw2, db2, dw3.view_as(w3), db3, None, None, None, None, None, None


class GatherRowsFn(Function):
    """table[idx]  (nn.Embedding lookup, model.py:225-226, for index sets other than arange)."""

    @staticmethod
    def forward(ctx, table, idx):
        lib = _lib.load()
        table = _f32(table, "table")
        idx = idx.contiguous()
        if idx.dtype != torch.int64:
            idx = idx.long()
        out = torch.empty((idx.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
        _run("b2g_gather_rows", lib.b2g_gather_rows, table.data_ptr(), idx.data_ptr(), idx.numel(), table.shape[0], table.shape[1],
                                       out.data_ptr(), _stream())
        ctx.idx = idx
        ctx.n = table.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _f32(dout, "grad")
        csr = CSR(ctx.idx, ctx.idx, ctx.n, ctx.n, col_is_eid=True)
        dt = torch.empty((ctx.n, dout.shape[1]), dtype=torch.float32, device=dout.device)
        gather_reduce_([csr], [dout], [None], [None], dt, False)
        return dt, None


class WeightedLossFn(Function):
    """mean over supervised pairs of w[lab] * (|p - t| | (p - t)^2 | huber)  (train.py:364-386,
    model.py:602-605).  One fused reduction; the gradient is produced in the same call."""

    @staticmethod
    def forward(ctx, pred, target, lab, w, sup, kind):
        lib = _lib.load()
        pred, target = _f32(pred, "pred"), _f32(target, "target")
        m = pred.numel()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        if sup is not None:
            sup = sup.contiguous()
            sup = sup.view(torch.uint8) if sup.dtype == torch.bool else sup.to(torch.uint8)
        ws = workspace(lib.b2g_loss_ws_bytes(m), pred.device)
        _run("b2g_weighted_loss", lib.b2g_weighted_loss, pred.data_ptr(), target.data_ptr(), _ptr(lab), _ptr(w), _ptr(sup), m, int(kind),
                                         loss.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None, None, None


LOSS_KINDS = {"mae": 0, "mse": 1, "huber": 2}


def weighted_loss(pred, target, lab_idx=None, lab_weights=None, sup_mask=None, loss_type="mae"):
    if loss_type not in LOSS_KINDS:
        raise ValueError(f"Unknown loss type: {loss_type}")
    return WeightedLossFn.apply(pred, target, lab_idx, lab_weights, sup_mask, LOSS_KINDS[loss_type])

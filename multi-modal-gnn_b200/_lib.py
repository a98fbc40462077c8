"""ctypes binding of libb2g.so (C ABI declared in include/b2g.h) + the in-tree nvcc build recipe.

There is deliberately no CPU fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError (BASELINE.json north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb2g.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["graph.cu", "spmm.cu", "dense.cu", "norm.cu", "decoder.cu", "dense_tc.cu", "layer_tc.cu", "ingest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


class B2GError(RuntimeError):
    pass


class RelT(ctypes.Structure):
    """b2g_rel_t"""
    _fields_ = [("rowptr", c_void_p), ("col", c_void_p), ("x", c_void_p), ("row_scale", c_void_p),
                ("col_scale", c_void_p)]


class GemmProblemT(ctypes.Structure):
    """b2g_gemm_problem_t"""
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("a2", c_void_p), ("b2", c_void_p), ("bias", c_void_p), ("c", c_void_p),
                ("m", ctypes.c_int32), ("n", ctypes.c_int32), ("k", ctypes.c_int32), ("k2", ctypes.c_int32),
                ("a_transposed", ctypes.c_int32), ("b_is_nk", ctypes.c_int32), ("accumulate", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class BitLayoutT(ctypes.Structure):
    """b2g_bit_layout_t"""
    _fields_ = [("nw", ctypes.c_int32), ("rel_a", ctypes.c_int8 * 24), ("rel_b", ctypes.c_int8 * 24), ("split", ctypes.c_int8 * 24)]


def _sources():
    extra = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f not in SOURCES)
    return SOURCES + extra


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "b2g.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into multi-modal-gnn_b200/libb2g.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise B2GError("nvcc not found: cannot build libb2g.so")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise B2GError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-o", LIB_PATH + ".tmp", *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise B2GError(f"link failed:\n{r.stdout}")
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


_P = c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    "b2g_last_error": (c_char_p, []),
    "b2g_version": (c_int, []),
    "b2g_launch_count": (c_ulonglong, []),
    "b2g_reset_launch_count": (None, []),
    "b2g_csr_build_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "b2g_csr_build": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "b2g_csr_degrees": (c_int, [_P, c_int64, _P, _P, _P]),
    "b2g_degree_gate": (c_int, [_P, _P, c_int64, c_int64, _P, _P]),
    "b2g_csr_chunk_ws_bytes": (c_size_t, [c_int64]),
    "b2g_csr_chunk_count": (c_int, [_P, c_int64, c_int32, _P, ctypes.POINTER(c_int64), _P, c_size_t, _P]),
    "b2g_csr_chunk_fill": (c_int, [_P, c_int64, c_int32, _P, _P, _P, _P]),
    "b2g_gather_reduce": (c_int, [ctypes.POINTER(RelT), c_int, c_int64, c_int, _P, c_int, _P]),
    "b2g_gather_reduce_stream": (c_int, [ctypes.POINTER(RelT), c_int, c_int64, c_int, _P, c_int, _P]),
    "b2g_gather_reduce_staged_supported": (c_int, [ctypes.POINTER(c_int), c_int, c_int]),
    "b2g_gather_reduce_staged": (c_int, [ctypes.POINTER(RelT), ctypes.POINTER(c_int), c_int, c_int64, c_int, _P, c_int, _P]),
    "b2g_gather_reduce_chunked": (c_int, [ctypes.POINTER(RelT), _P, _P, _P, c_int64, c_int32, c_int64, c_int, _P, c_int,
                                          _P, c_size_t, _P]),
    "b2g_dense_adjacency": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "b2g_transpose_pad": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "b2g_row_scale": (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    "b2g_gather_rows": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P]),
    "b2g_gather_add_rows": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, _P]),
    "b2g_scatter_values": (c_int, [_P, _P, c_int64, _P, _P]),
    "b2g_gather_values": (c_int, [_P, _P, c_int64, _P, _P]),
    "b2g_linear_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, c_int, _P]),
    "b2g_linear_fwd_tc_supported": (c_int, [c_int64, c_int, c_int]),
    "b2g_linear_fwd_tc": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, c_int, _P]),
    "b2g_linear_stats_ws_bytes": (c_size_t, [c_int]),
    "b2g_linear_fwd_tc_ex": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P, c_float, _P]),
    "b2g_transpose": (c_int, [_P, c_int, c_int, _P, _P]),
    "b2g_linear_bwd_weight_tc_supported": (c_int, [c_int64, c_int, c_int]),
    "b2g_linear_bwd_weight_tc_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "b2g_linear_bwd_weight_tc": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, c_size_t, _P]),
    "b2g_col_sums": (c_int, [_P, c_int64, c_int, _P, _P, c_size_t, _P]),
    "b2g_linear_bwd_input": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P]),
    "b2g_linear_bwd_weight_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "b2g_linear_bwd_weight": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b2g_bn_ws_bytes": (c_size_t, [c_int]),
    "b2g_bn_stats": (c_int, [_P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b2g_bn_local_sums": (c_int, [_P, c_int64, c_int, _P, _P, c_size_t, _P]),
    "b2g_bn_finalize_sums": (c_int, [_P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "b2g_bn_bwd_local_sums": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, c_int, c_float, c_uint64, c_uint64, _P, _P, c_size_t, _P]),
    "b2g_bn_bwd_from_sums": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P, _P, _P, c_int, c_float, c_uint64, c_uint64, _P, _P, _P,
                                     _P, _P]),
    "b2g_bn_eval_stats": (c_int, [_P, _P, c_int, c_float, _P, _P, _P]),
    "b2g_bn_apply": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "b2g_bn_bwd": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, c_int, c_float, c_uint64, c_uint64, c_int, _P, _P, _P, _P,
                           _P, c_size_t, _P]),
    "b2g_comm_region_bytes": (c_size_t, []),
    "b2g_comm_max_bytes": (c_size_t, []),
    "b2g_comm_local_alloc": (c_int, [ctypes.POINTER(c_void_p), ctypes.c_char_p]),
    "b2g_comm_create": (c_int, [c_int, c_int, _P, ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "b2g_comm_destroy": (c_int, [_P]),
    "b2g_comm_error": (c_int, [_P]),
    "b2g_comm_set_timeout": (c_int, [_P, ctypes.c_double]),
    "b2g_comm_allreduce_f32": (c_int, [_P, _P, _P, c_int64, _P]),
    "b2g_comm_allreduce_f64": (c_int, [_P, _P, _P, c_int64, _P]),
    "b2g_bn_stats_sync": (c_int, [_P, _P, c_int64, c_int64, c_int, c_float, c_float, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b2g_bn_bwd_sync": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, _P, _P, _P, _P, c_int, c_float, c_uint64, c_uint64, _P, _P, _P, _P,
                                _P, c_size_t, _P]),
    "b2g_adam_chunk_elems": (c_int, []),
    "b2g_adam_step": (c_int, [_P, c_int, _P, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                              ctypes.c_double, ctypes.c_double, _P]),
    "b2g_eval_fields": (c_int, []),
    "b2g_eval_per_lab": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, _P, _P, _P]),
    "b2g_eval_per_lab_strata": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_float, _P, _P, _P, _P]),
    "b2g_small_gemm_group": (c_int, [ctypes.POINTER(GemmProblemT), c_int, _P]),
    "b2g_small_colsum_group": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                       c_int, _P]),
    "b2g_relu_dropout_fwd": (c_int, [_P, c_int64, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "b2g_relu_dropout_bwd": (c_int, [_P, _P, c_int64, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "b2g_dropout_mask": (c_int, [c_int64, c_float, c_uint64, c_uint64, _P, _P]),
    "b2g_l2norm_fwd": (c_int, [_P, c_int64, c_int, c_float, _P, _P, _P]),
    "b2g_l2norm_bwd": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    "b2g_l2norm_bwd_cs_ws_bytes": (c_size_t, [c_int]),
    "b2g_l2norm_bwd_cs": (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P, c_size_t, _P]),
    "b2g_decoder_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, c_uint64, _P, _P]),
    "b2g_decoder_fwd_tc": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, c_uint64, _P, _P]),
    "b2g_decoder_bwd_ws_bytes": (c_size_t, [c_int64]),
    "b2g_decoder_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, c_uint64, _P, _P, _P, _P,
                                _P, _P, _P, c_size_t, _P]),
    "b2g_decoder_bwd_tc": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, c_uint64, _P, _P, _P, _P,
                                   _P, _P, _P, c_size_t, _P]),
    "b2g_id_lookup": (c_int, [_P, _P, c_int64, _P, c_int64, _P, _P]),
    "b2g_edges_from_rows_ws_bytes": (c_size_t, [c_int64]),
    "b2g_edges_from_rows": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, ctypes.POINTER(c_int64), _P, c_size_t, _P]),
    "b2g_adj_bits_build": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    "b2g_layer_cat_weights": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, ctypes.POINTER(c_void_p), c_int, _P, c_int, c_int,
                                      ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                      c_int, c_int, _P, _P]),
    "b2g_layer_fwd_tc_supported": (c_int, [c_int64, c_int, c_int, c_int]),
    "b2g_layer_stats_ws_bytes": (c_size_t, [c_int]),
    "b2g_layer_cat_half": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "b2g_layer_fwd_tc": (c_int, [_P, _P, _P, _P, _P, _P, ctypes.POINTER(BitLayoutT), ctypes.POINTER(c_void_p), c_int64, c_int, c_int, _P, _P, _P,
                                 c_size_t, _P]),
    "b2g_layer_adjT_tc_supported": (c_int, [c_int64, c_int, c_int]),
    "b2g_layer_adjT_tc_ws_bytes": (c_size_t, [c_int]),
    "b2g_layer_adjT_tc": (c_int, [_P, _P, ctypes.POINTER(BitLayoutT), ctypes.POINTER(c_void_p), _P, c_int64, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "b2g_loss_ws_bytes": (c_size_t, [c_int64]),
    "b2g_weighted_loss": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, c_size_t, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and bind every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B2GError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root. "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise B2GError(f"libb2g.so does not export {name}; rebuild it") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_PROTOS)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b2g_last_error()
        raise B2GError(f"{what or 'libb2g'} failed (code {rc}): {msg.decode() if msg else ''}")

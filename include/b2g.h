/*
 * b2g.h -- C ABI of libb2g.so: the B200 (sm_100a) kernels behind the hetero-GNN training / imputation
 * hot path of AdalineL/Multi-Modal-GNN (reference: /root/reference/src/model.py, src/train.py).
 *
 * The reference has no FFI of its own (pure Python; all arithmetic in the torch / torch-geometric wheels),
 * so the boundary is: the Python surface of src/model.py (kept by multi-modal-gnn_b200/model.py) on top of
 * THIS library.  Every entry point below names the reference expression it replaces.
 *
 * Conventions
 *   - plain pointers + sizes; no torch types; all pointers are DEVICE pointers unless named h_*;
 *   - every function returns 0 on success, a negative B2G_E* code otherwise; b2g_last_error() gives text;
 *   - nothing allocates or frees caller memory; scratch space is passed in (`ws`, size from *_ws_bytes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation unless
 *     stated ("SYNC");
 *   - feature matrices are row-major fp32 [rows, d], 16-byte aligned, d in {32, 64, 128, 256};
 *   - indices inside the library are int32 (E < 2^31, rows < 2^31); row*d offsets are 64-bit;
 *   - results are deterministic (run-to-run bit-identical): no floating-point atomics anywhere.
 */
#ifndef B2G_H_
#define B2G_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_OK 0
#define B2G_EINVAL (-1)   /* bad argument (shape, alignment, null pointer)                     */
#define B2G_ECUDA (-2)    /* a CUDA runtime call failed                                        */
#define B2G_ERANGE (-3)   /* an edge endpoint is outside [0, n) -- graph_build.py:611-633       */
#define B2G_EWS (-4)      /* workspace too small                                               */

const char* b2g_last_error(void);
int b2g_version(void);
/* number of kernels this library has launched since load / since reset (bench.py's gpu_launches) */
unsigned long long b2g_launch_count(void);
void b2g_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (a) graph structure -- replaces PyG's COO index_select/scatter bookkeeping
 * ---------------------------------------------------------------------------------------------- */

/* Stable CSR of one relation, keyed by `key` (the destination for a forward CSR, the source for the
 * transposed one).  Input is the reference's COO layout: edge_index[2,E] int64 (graph_build.py:515),
 * key = edge_index[1] or [0], val = the other row.
 *   rowptr[n_rows+1]  exclusive prefix of the per-row edge counts (bit-exact torch.bincount / cumsum)
 *   col[E]            val[] reordered so that row r owns col[rowptr[r] .. rowptr[r+1])
 *   eid[E]            original edge position of each slot; within a row, eid is increasing (stable)
 * ws: b2g_csr_build_ws_bytes(E, n_rows) bytes.  Returns B2G_ERANGE (after a SYNC on `stream`) if any
 * key is outside [0,n_rows) or any val outside [0,n_vals) -- mirrors validate_graph's ValueError. */
size_t b2g_csr_build_ws_bytes(int64_t n_edges, int64_t n_rows);
int b2g_csr_build(const int64_t* key, const int64_t* val, int64_t n_edges, int64_t n_rows, int64_t n_vals,
                  int32_t* rowptr, int32_t* col, int32_t* eid, void* ws, size_t ws_bytes, void* stream);

/* deg[r] = rowptr[r+1]-rowptr[r] as int64  == torch.bincount(edge_index[0], minlength=N) (model.py:297-298)
 * inv[r] = 1 / max(deg[r], 1) as fp32      == PyG mean aggregation's clamp(count, min=1) reciprocal */
int b2g_csr_degrees(const int32_t* rowptr, int64_t n_rows, int64_t* deg, float* inv, void* stream);

/* low[i] = deg[patient_idx[i]] < threshold (uint8 0/1)     model.py:312-315 (threshold 6, model.py:178) */
int b2g_degree_gate(const int64_t* deg, const int64_t* patient_idx, int64_t m, int64_t threshold,
                    uint8_t* low, void* stream);

/* Work decomposition for rows that are too long for one warp (type-destination rows own up to millions
 * of patient neighbours).  Row r is cut into max(1, ceil(deg/chunk)) items; item i of row item_row[i]
 * covers CSR slots [item_start[i], min(item_start[i]+chunk, rowptr[row+1])); row r owns items
 * [row_item_ptr[r], row_item_ptr[r+1]).  _count fills row_item_ptr[n_rows+1] and returns the number of
 * items in *h_n_items (SYNC); _fill then writes item_row / item_start (n_items entries each). */
size_t b2g_csr_chunk_ws_bytes(int64_t n_rows);
int b2g_csr_chunk_count(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int32_t* row_item_ptr,
                        int64_t* h_n_items, void* ws, size_t ws_bytes, void* stream);
int b2g_csr_chunk_fill(const int32_t* rowptr, int64_t n_rows, int32_t chunk, const int32_t* row_item_ptr,
                       int32_t* item_row, int32_t* item_start, void* stream);

/* (a') graph ingest -- replaces the per-row Python loops of graph_build.py:476-586.
 * out[i] = index[j] where sorted_ids[j] == query[i], or -1: NodeIndexer.get_index (graph_build.py:84-89) for every table row;
 * sorted_ids int64[n_dict] ascending and distinct, index int32[n_dict] the node index of each id. */
int b2g_id_lookup(const int64_t* sorted_ids, const int32_t* index, int64_t n_dict, const int64_t* query, int64_t m,
                  int32_t* out, void* stream);
/* SYNC.  Rows i with src_idx[i] >= 0 and dst_idx[i] >= 0 become edges IN ROW ORDER (graph_build.py:502-508: rows with an unknown
 * entity are dropped): edge_index[e] = src_idx[i], edge_index[*h_n_edges + e] = dst_idx[i] (int64; the caller's buffer holds
 * 2 m entries, the result is the [2, E] tensor of graph_build.py:515 in its first 2 E entries), attr_out[e] = attr[i]
 * (optional), row_of_edge[e] = i (optional).  ws: b2g_edges_from_rows_ws_bytes(m). */
size_t b2g_edges_from_rows_ws_bytes(int64_t m);
int b2g_edges_from_rows(const int32_t* src_idx, const int32_t* dst_idx, const float* attr, int64_t m, int64_t* edge_index,
                        float* attr_out, int64_t* row_of_edge, int64_t* h_n_edges, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (b) message passing -- replaces SAGEConv.propagate: index_select + scatter_add + /clamp(count,1)
 *     (PyG sage_conv.py / utils/scatter.py; call sites model.py:125-131,256)
 * ---------------------------------------------------------------------------------------------- */

/* One relation of a fused gather-reduce. */
typedef struct {
  const int32_t* rowptr;     /* [n_rows+1] CSR keyed by the OUTPUT row                          */
  const int32_t* col;        /* [E] neighbour row in `x`                                        */
  const float* x;            /* [n_x, d] neighbour features                                     */
  const float* row_scale;    /* [n_rows] or NULL: multiplies the reduced row (1/deg for a mean)  */
  const float* col_scale;    /* [n_x]    or NULL: multiplies each gathered row (transposed mean) */
} b2g_rel_t;

/* out[r,:] = (accumulate ? out[r,:] : 0) + sum_k row_scale_k[r] * sum_{j in row r of rel k}
 *            col_scale_k[col[j]] * x_k[col[j], :]
 * Warp-per-row segmented gather-reduce, 128-bit row loads, up to 4 relations fused (patient
 * destinations: lab + diagnosis + medication neighbours in one pass over the patient rows).
 * Intended for SHORT rows (degree <= a few hundred). */
int b2g_gather_reduce(const b2g_rel_t* h_rels, int n_rels, int64_t n_rows, int d, float* out,
                      int accumulate, void* stream);

/* The same contraction, bit-identical results, for MANY SHORT rows (hundreds of thousands of rows with a few entries each, e.g.
 * the per-patient reduction of the decoder's pair gradients): a warp owns 32 consecutive rows, their row pointers are one
 * coalesced load and their column indices one stream walked in register chunks, so a row costs one dependent global load (the
 * gathered rows) instead of four (row pointer -> column -> column scale -> row). */
int b2g_gather_reduce_stream(const b2g_rel_t* h_rels, int n_rels, int64_t n_rows, int d, float* out, int accumulate, void* stream);

/* The same contraction (identical arithmetic and edge order) with every SOURCE TABLE STAGED IN SHARED MEMORY by bulk copies
 * (cp.async.bulk + mbarrier), one resident CTA per SM: for patient destinations, whose sources are the few-hundred-row lab /
 * diagnosis / medication tables (north_star (b)).  h_n_src[k] = rows of rels[k].x; supported when they sum to <= 200 KB. */
int b2g_gather_reduce_staged_supported(const int* h_n_src, int n_rels, int d);
int b2g_gather_reduce_staged(const b2g_rel_t* h_rels, const int* h_n_src, int n_rels, int64_t n_rows, int d, float* out,
                             int accumulate, void* stream);

/* Same contraction for LONG rows (type destinations).  Two deterministic phases: one warp per chunk
 * item writes a partial row into ws, then one warp per output row adds its partials in item order.
 * ws: n_items * d * 4 bytes. */
int b2g_gather_reduce_chunked(const b2g_rel_t* h_rel, const int32_t* item_row, const int32_t* item_start,
                              const int32_t* row_item_ptr, int64_t n_items, int32_t chunk, int64_t n_rows,
                              int d, float* out, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* Tensor-core formulation of the same contraction for the dense-ish bipartite EHR relations (millions of patients x a
 * vocabulary of <= 256 labs / diagnoses / drugs): the adjacency is held as a dense fp32 [n_big, pad] matrix D
 * (b2g_dense_adjacency; entries 0/1, or 0 / (1/deg_row)) and the four products of a relation become tcgen05 GEMMs:
 *   mean onto the small side   agg_T = diag(1/deg_T) D^T x_P      -> b2g_linear_bwd_weight_tc(dy = D, x = x_P) + b2g_row_scale
 *   its backward               dx_P  = D (dagg_T / deg_T)         -> b2g_linear_fwd_tc(x = D, w = b2g_transpose_pad(dagg_T, 1/deg_T))
 *   mean onto the big side     out_P += D' Y_T, D' = diag(1/deg_P) D -> b2g_linear_fwd_tc(x = D', w = b2g_transpose_pad(Y_T), accumulate)
 *   its backward               dY_T  = D'^T dout_P                -> b2g_linear_bwd_weight_tc(dy = D', x = dout_P) */
int b2g_dense_adjacency(const int32_t* rowptr, const int32_t* col, const float* row_val, int64_t n_rows, int pad,
                        float* out, void* stream);
int b2g_transpose_pad(const float* in, const float* scale, int rows, int cols, int pad, float* out, void* stream);
int b2g_row_scale(const float* in, const float* scale, int64_t rows, int d, float* out, void* stream);

/* (b') The patient side of a whole HeteroConv layer in one launch per direction (csrc/layer_tc.cu) -- replaces, for the
 * three type -> patient relations and the three patient -> type relations of model.py:125-131,256 together, PyG's
 * SAGEConv.propagate + lin_l + lin_r + HeteroConv's stack().sum(0) on the patient rows.
 * The patient x type adjacency of all relations is ONE bit matrix bits[N_patient, nw] (32 type columns per word; relation r
 * owns the column range [bit_off_r, bit_off_r + n_type_r)); b2g_bit_layout_t says, per word, which relation its low bits
 * [0, split) and high bits [split, 32) belong to (split = 32: one relation; at most one boundary per word), so that a per
 * relation row scale (1/deg of the patient, PyG's mean) can be applied while the bits are expanded into TF32 tiles. */
typedef struct {
  int32_t nw;           /* words per row, <= 24 */
  int8_t rel_a[24];     /* relation (index into h_rscale) of bits [0, split) of word k   */
  int8_t rel_b[24];     /* relation of bits [split, 32)                                   */
  int8_t split[24];     /* 1..32                                                          */
} b2g_bit_layout_t;
/* bits[row, (bit_off + col[j]) / 32] |= 1 << ((bit_off + col[j]) % 32) for every neighbour j of `row` in a by-patient CSR.
 * `bits` must be zeroed by the caller before the first relation is added. */
int b2g_adj_bits_build(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int nw, int bit_off, uint32_t* bits,
                       void* stream);
/* out[n, ktot] = [ W (or W^T) | tab_0^T * scale_0 | tab_1^T * scale_1 | ... ] : the reduction operand of b2g_layer_fwd_tc.
 * W = sum of the n_w (<= 4) matrices h_ws[i] (HeteroConv sums the lin_r products of all relations that share a destination),
 * each [n, kx] (w_transposed = 0: nn.Linear weight, y = x W^T) or [kx, n] (w_transposed = 1: dx = dy W); tab_r is
 * [rows_r, n] and lands in columns kx + off_r .. ; scale_r [rows_r] or NULL; all other columns are zero.
 * bias_out[n] = sum of the n_b biases (n_b = 0: untouched).  h_* are HOST arrays of device pointers / ints. */
int b2g_layer_cat_weights(const float* const* h_ws, int n_w, int w_transposed, const float* const* h_biases, int n_b,
                          float* bias_out, int n, int kx, const float* const* h_tabs, const float* const* h_scales,
                          const int* h_rows, const int* h_offs, int n_rel, int ktot, float* out, void* stream);
/* Half mode of the adjacency columns (optional; what the default tf32 precision mode uses): for every row j of wcat (= output
 * column j of the layer), whalf[j, :] = fp16_rn(wcat[j, kx:] * 2^e_j), zero-padded to 64 * ceil(nw / 2) columns, with the power
 * of two that puts the row's largest |value| into [2^13, 2^14); wcat[j, :kx] *= 2^e_j in place (exact); unscale[j] = 2^-e_j.
 * The adjacency entries are {0, 1/deg} -- fp16 holds them with 11 significant bits, rounded to nearest (TF32 as the tensor core
 * reads it: 11 bits, truncated) -- so b2g_layer_fwd_tc can run the adjacency part of the reduction with kind::f16 MMAs on
 * operand tiles half as large as the TF32 ones (shared-memory bandwidth, not HBM, bounds that kernel). */
int b2g_layer_cat_half(float* wcat, int n, int kx, int ktot, uint16_t* whalf, float* unscale, void* stream);
/* y[m, n] = x[m, kx] . wcat[:, :kx]^T + (diag(rscale) A)[m, 32 nw] . wcat[:, kx:]^T + bias      (tcgen05, TF32 operands)
 *   forward : x = x_patient, wcat = [sum_r W_r,root | Y_lab | Y_dx | Y_med] with Y_r = x_r W_l,r^T, rscale_r = 1/deg_r(patient)
 *   backward: x = dout_patient, wcat = [W_root^T | dagg_r / deg_r(type)], rscale = NULL  ->  dx_patient
 * K-chunk pipelined: TMA (x chunk, wcat chunk) + expander warps (bits -> swizzled TF32 A chunk) -> tcgen05.mma -> TMEM,
 * double-buffered accumulator, epilogue TMEM -> staging -> coalesced stores.  stat_sums (optional, n <= 128): fp64
 * {sum, sum of squares} per column of y (BatchNorm statistics of the layer output, model.py:259-261) produced by the
 * epilogue; ws: b2g_layer_stats_ws_bytes(n) (only when stat_sums is given).
 * whalf / unscale (both or neither): the outputs of b2g_layer_cat_half -- the adjacency part then runs in half mode
 * (x part: kind::tf32, adjacency part: kind::f16 into the same fp32 TMEM accumulator; y = acc * unscale + bias). */
int b2g_layer_fwd_tc_supported(int64_t m, int n, int kx, int nw);
size_t b2g_layer_stats_ws_bytes(int n);
int b2g_layer_fwd_tc(const float* x, const float* wcat, const uint16_t* whalf, const float* unscale, const float* bias, const uint32_t* bits,
                     const b2g_bit_layout_t* h_layout, const float* const* h_rscale, int64_t m, int n, int kx, float* y,
                     double* stat_sums, void* ws, size_t ws_bytes, void* stream);
/* out[32 nw, 128] = col_scale[:, None] * (diag(rscale) A)^T . x[m, 128]       (tcgen05, MN-major TF32 operands)
 *   forward : x = x_patient, col_scale = 1/deg(type)  -> mean of the patient neighbours of every type node
 *   backward: x = dout_patient, rscale_r = 1/deg_r(patient) -> dY_r
 * with_colsum = 1: `out` has 32 more rows, [32 (nw + 1), 128], and row 32 nw holds the column sums of x (the bias gradient
 * when x = dout) -- one constant column of ones appended to the expanded adjacency; col_scale then has 32 (nw + 1) entries.
 * dense_b / dense_out (both or neither): dense_out[128, 128] = x^T . dense_b[m, 128] in the same pass (backward:
 * dW_root[out, in] = dout^T x_patient), i.e. D = x^T [dense_b | diag(rscale) A | 1] with one read of x.
 * ws: b2g_layer_adjT_tc_ws_bytes(nw + with_colsum).  Supported for d = 128 and 128 [dense_b] + 32 (nw + with_colsum) <= 512. */
int b2g_layer_adjT_tc_supported(int64_t m, int d, int nw);
size_t b2g_layer_adjT_tc_ws_bytes(int nw);
int b2g_layer_adjT_tc(const float* x, const uint32_t* bits, const b2g_bit_layout_t* h_layout, const float* const* h_rscale,
                      const float* col_scale, int64_t m, int with_colsum, const float* dense_b, float* dense_out, float* out,
                      void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (c) embedding tables -- nn.Embedding(arange(N)) fwd / dense grad (model.py:225-226)
 * ---------------------------------------------------------------------------------------------- */
/* out[i,:] = table[idx[i],:]   (idx int64, as the reference passes them) */
int b2g_gather_rows(const float* table, const int64_t* idx, int64_t m, int64_t n_table, int d, float* out,
                    void* stream);
/* fused decoder input: out[i,:] = act(a[ia[i],:] + b[ib[i],:]) with act = ReLU (optional) */
int b2g_gather_add_rows(const float* a, const int64_t* ia, const float* b, const int64_t* ib, int64_t m,
                        int d, int relu, float* out, void* stream);

/* out[idx[i]] = vals[i] / out[i] = src[idx[i]] for fp32 scalars: assembling the per-head predictions into
 * predictions[M] (model.py:317-333 masked writes) and its backward. idx entries must be unique for scatter. */
int b2g_scatter_values(const float* vals, const int64_t* idx, int64_t m, float* out, void* stream);
int b2g_gather_values(const float* src, const int64_t* idx, int64_t m, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (d) dense layers -- nn.Linear / SAGEConv.lin_l / lin_r / EdgeRegressionHead (model.py:93-103,373-386)
 * ---------------------------------------------------------------------------------------------- */
/* Grouped small GEMMs: the type-node side of a HeteroConv layer (lab / diagnosis / medication rows: three relations x
 * {lin_l on the sources, lin_r on the destinations, lin_l on the aggregates}, forward and backward) is ~50 launches of
 * <= 200-row problems per step; one launch takes up to 12 of them.  C[m,n] = (accumulate ? C : 0) + opA(a) opB(b)
 * (+ opA(a2) opB(b2)) + bias.  a_transposed: a is stored [k, m]; b_is_nk: b is stored [n, k] (nn.Linear weight). */
typedef struct {
  const float* a;  const float* b;      /* first product, reduction length k                       */
  const float* a2; const float* b2;     /* optional second product into the same C, length k2      */
  const float* bias;                    /* [n] or NULL                                             */
  float* c;                             /* [m, n]                                                  */
  int32_t m, n, k, k2;
  int32_t a_transposed, b_is_nk, accumulate, reserved;
} b2g_gemm_problem_t;
int b2g_small_gemm_group(const b2g_gemm_problem_t* h_probs, int n_probs, void* stream);
int b2g_small_colsum_group(const float* const* h_x, float* const* h_out, const int* h_m, const int* h_n, int count,
                           void* stream);
/* y[M,N] = (accumulate ? y : 0) + x[M,K] * w[N,K]^T + bias[N]   (bias may be NULL)  fp32 SIMT */
int b2g_linear_fwd(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y,
                   int accumulate, void* stream);
/* dx[M,K] = (accumulate ? dx : 0) + dy[M,N] * w[N,K] */
int b2g_linear_bwd_input(const float* dy, const float* w, int64_t m, int n, int k, float* dx, int accumulate,
                         void* stream);
/* dw[N,K] = dy[M,N]^T * x[M,K];  db[N] = column sums of dy (db may be NULL).  Deterministic split over M.
 * ws: b2g_linear_bwd_weight_ws_bytes(m, n, k). */
size_t b2g_linear_bwd_weight_ws_bytes(int64_t m, int n, int k);
int b2g_linear_bwd_weight(const float* dy, const float* x, int64_t m, int n, int k, float* dw, float* db,
                          void* ws, size_t ws_bytes, void* stream);

/* Same product on the tcgen05 tensor cores: fp32 operands consumed as TF32 (kind::tf32), fp32 accumulation in TMEM,
 * fp32 output; TMA (cp.async.bulk.tensor, 128-byte swizzle) stages X tiles and keeps W resident in shared memory.
 * Relative error ~1e-3 (north_star tolerance for tensor-core GEMM outputs: 1e-2).  _supported() says whether
 * (m, n, k) fits: n % 32 == 0, n <= 256, k % 32 == 0, W + at least one X stage within 227 KB of shared memory.
 * dx = dy W is the same call on W^T (b2g_transpose). */
int b2g_linear_fwd_tc_supported(int64_t m, int n, int k);
int b2g_linear_fwd_tc(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y,
                      int accumulate, void* stream);
/* The same kernel with an extended epilogue (n <= 128):
 *   stat_sums != NULL: fp64 {sum y, sum y^2} per column of y -- the statistics of the nn.BatchNorm1d that follows
 *                      (model.py:93-101) without a pass of its own; ws: b2g_linear_stats_ws_bytes(n);
 *   inv_norm  != NULL: the rows of y leave the kernel L2-normalised, y /= max(||y||_2, l2_eps) (F.normalize after the last
 *                      patient-MLP linear, model.py:103-105,232); inv_norm[m] = the reciprocal norms (saved for backward). */
size_t b2g_linear_stats_ws_bytes(int n);
int b2g_linear_fwd_tc_ex(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y,
                         double* stat_sums, void* ws, size_t ws_bytes, float* inv_norm, float l2_eps, void* stream);
int b2g_transpose(const float* in, int rows, int cols, float* out, void* stream);
/* dW[N,K] = dy[M,N]^T x[M,K] on tcgen05: both operands MN-major (the reduction index M is the slow one in memory), one
 * fp32 TMEM accumulator per CTA over its whole share of M, per-CTA partials added in fixed order.  Supported when one of
 * N, K is 128 and the other a multiple of 32 in [32, 256].  ws: b2g_linear_bwd_weight_tc_ws_bytes.  The bias gradient
 * is b2g_col_sums(dy) (ws: b2g_bn_ws_bytes(n)). */
int b2g_linear_bwd_weight_tc_supported(int64_t m, int n, int k);
size_t b2g_linear_bwd_weight_tc_ws_bytes(int64_t m, int n, int k);
int b2g_linear_bwd_weight_tc(const float* dy, const float* x, int64_t m, int n, int k, float* dw, void* ws,
                             size_t ws_bytes, void* stream);
int b2g_col_sums(const float* x, int64_t m, int d, float* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (d') BatchNorm1d + ReLU + Dropout, row L2 normalisation (model.py:93-105,134-139,259-269)
 * ---------------------------------------------------------------------------------------------- */
/* Batch statistics of x[M,d] (training mode): mean[d], rstd[d] = 1/sqrt(biased var + eps); optionally
 * updates running_mean / running_var in place with `momentum` and the UNBIASED variance, like
 * nn.BatchNorm1d.  fp64 accumulation, two deterministic stages.  ws: b2g_bn_ws_bytes(d). */
size_t b2g_bn_ws_bytes(int d);
int b2g_bn_stats(const float* x, int64_t m, int d, float eps, float momentum, float* mean, float* rstd,
                 float* running_mean, float* running_var, void* ws, size_t ws_bytes, void* stream);
/* Patient-partitioned (multi-GPU) BatchNorm: local fp64 column totals sums[2*d] = {sum x, sum x^2} -> (caller all-reduces)
 * -> mean / rstd from the GLOBAL totals and row count.  Backward likewise with {sum g, sum g*xhat}; m_total is the global
 * row count used in dx = gamma*rstd*(g - sum_g/m_total - xhat*sum_gx/m_total). */
int b2g_bn_local_sums(const float* x, int64_t m, int d, double* sums, void* ws, size_t ws_bytes, void* stream);
int b2g_bn_finalize_sums(const double* sums, int64_t m_total, int d, float eps, float momentum, float* mean,
                         float* rstd, float* running_mean, float* running_var, void* stream);
int b2g_bn_bwd_local_sums(const float* x, const float* dy, int64_t m, int d, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, int relu, float p_drop, uint64_t seed,
                          uint64_t stream_id, double* sums, void* ws, size_t ws_bytes, void* stream);
int b2g_bn_bwd_from_sums(const float* x, const float* dy, int64_t m, int64_t m_total, int d, const float* mean,
                         const float* rstd, const float* gamma, const float* beta, int relu, float p_drop,
                         uint64_t seed, uint64_t stream_id, const double* sums, float* dx, float* dgamma,
                         float* dbeta, void* stream);
/* eval mode: mean = running_mean, rstd = 1/sqrt(running_var + eps) */
int b2g_bn_eval_stats(const float* running_mean, const float* running_var, int d, float eps, float* mean,
                      float* rstd, void* stream);
/* y = dropout(act((x - mean) * rstd * gamma + beta)); `relu` is the activation code of model.py:145-153:
 * 0 none, 1 relu, 2 leaky_relu(0.01), 3 elu(1.0); dropout optional (p_drop == 0 -> none).
 * The keep mask is Philox4x32-7(seed, stream_id) per element, scaled by 1/(1-p) (F.dropout semantics). */
int b2g_bn_apply(const float* x, int64_t m, int d, const float* mean, const float* rstd, const float* gamma,
                 const float* beta, int relu, float p_drop, uint64_t seed, uint64_t stream_id, float* y,
                 void* stream);
/* backward of the above.  Phase 1 (train mode only) reduces dgamma = sum(g * xhat), dbeta = sum(g) with
 * g = dy * mask * relu'; phase 2 writes dx.  In eval mode (batch_stats == 0) dx = g * gamma * rstd.
 * dx_colsum (optional, [d]): column sums of dx, produced by the same pass that writes dx -- the bias gradient
 * of the nn.Linear whose output was normalised (model.py:93-101), which would otherwise re-read dx. */
int b2g_bn_bwd(const float* x, const float* dy, int64_t m, int d, const float* mean, const float* rstd,
               const float* gamma, const float* beta, int relu, float p_drop, uint64_t seed,
               uint64_t stream_id, int batch_stats, float* dx, float* dgamma, float* dbeta, float* dx_colsum,
               void* ws, size_t ws_bytes, void* stream);
/* ------------------------------------------------------------------------------------------------
 * (f) multi-GPU exchange over NVLink peer memory (replaces the NCCL launches of the
 *     latency-bound all-reduces of the patient-partitioned mode -- csrc/peer.cuh, csrc/comm.cu)
 * Setup (SYNC, once per process): every rank allocates its symmetric region and exports a CUDA-IPC handle,
 * the host side exchanges the 64-byte handles (torch.distributed.all_gather_object), every rank maps its
 * peers.  At most 8 ranks, all on one node.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b2g_comm b2g_comm_t;
size_t b2g_comm_region_bytes(void);
size_t b2g_comm_max_bytes(void);      /* largest payload of one b2g_comm_allreduce_* call */
int b2g_comm_local_alloc(void** region, unsigned char* h_handle64);
int b2g_comm_create(int rank, int world, void* local_region, const unsigned char* h_handles, b2g_comm_t** out);
int b2g_comm_destroy(b2g_comm_t* comm);
int b2g_comm_error(b2g_comm_t* comm);  /* SYNC: 1 if a wait timed out (a rank did not show up within the time-out) */
/* seconds a rendezvous waits for a peer before giving up (default ~30 s; also B2G_PEER_TIMEOUT_S at creation) */
int b2g_comm_set_timeout(b2g_comm_t* comm, double seconds);
/* out[i] = sum over ranks of in[i] (in == out allowed), one-shot: every rank stores its payload in its own
 * region, signals every peer, and adds everybody's payload in RANK ORDER (bit-identical on all ranks,
 * deterministic).  Payload: a multiple of 16 bytes, 16-byte aligned, <= b2g_comm_max_bytes().  Every rank
 * must issue the same calls in the same order.  Graph-capturable (sequence numbers live on the device). */
int b2g_comm_allreduce_f32(b2g_comm_t* comm, const float* in, float* out, int64_t n, void* stream);
int b2g_comm_allreduce_f64(b2g_comm_t* comm, const double* in, double* out, int64_t n, void* stream);
/* Patient-partitioned BatchNorm with the exchange FUSED into the reduction kernel: the CTA that finishes last
 * publishes this rank's fp64 column totals, meets the same CTA of every other rank, adds the totals in rank
 * order and finalises -- one launch instead of local sums -> all-reduce -> finalize.  m = local rows,
 * m_total = rows over all ranks; dgamma / dbeta receive the GLOBAL totals. */
int b2g_bn_stats_sync(b2g_comm_t* comm, const float* x, int64_t m, int64_t m_total, int d, float eps,
                      float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                      void* ws, size_t ws_bytes, void* stream);
int b2g_bn_bwd_sync(b2g_comm_t* comm, const float* x, const float* dy, int64_t m, int64_t m_total, int d,
                    const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                    float p_drop, uint64_t seed, uint64_t stream_id, float* dx, float* dgamma, float* dbeta,
                    float* dx_colsum, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (g) optimizer -- torch.optim.Adam(params, lr, weight_decay) as built by train.py:255-260 (betas (0.9, 0.999),
 *     eps 1e-8, L2 weight decay added to the gradient, no amsgrad) for ALL parameter tensors in one launch
 *     (SURVEY.md section 8f item 4).  Parameters without a gradient are simply not listed (note N8).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float* param;          /* [numel] updated in place            */
  const float* grad;     /* [numel]                             */
  float* exp_avg;        /* [numel] first moment, in place      */
  float* exp_avg_sq;     /* [numel] second moment, in place     */
  int64_t numel;
} b2g_adam_tensor_t;
int b2g_adam_chunk_elems(void);
int b2g_adam_step(const b2g_adam_tensor_t* d_tensors, int n_tensors, const int32_t* d_chunks, int n_chunks,
                  double step_size, double bc2_sqrt, double one_minus_beta1, double beta2, double one_minus_beta2,
                  double eps, double weight_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (h) evaluation metrics -- evaluate.py:36-82 (MAE / RMSE / R^2 / MAPE), :88-139 (per lab), :417-440 (per-lab
 *     +-3 sigma winsorisation of the residuals), one CTA per lab over a by-lab CSR of the pair list
 *     (SURVEY.md section 8f item 2).  out[n_lab][b2g_eval_fields()] doubles: n, mean residual, residual std,
 *     number of clipped residuals, sum |r|, sum r^2, sum t, sum t^2, sum |r/t| over t != 0, count of t != 0.
 * ---------------------------------------------------------------------------------------------- */
int b2g_eval_fields(void);
int b2g_eval_per_lab(const float* pred, const float* target, const int32_t* rowptr, const int32_t* pair_of, int n_lab,
                     int winsorize, float n_sigma, double* out, float* pred_w, void* stream);
/* b2g_eval_per_lab + the patient-degree strata of evaluate.py:237-287: out_bins[n_lab][3][7] (double) per lab and degree
 * group {1-5, 6-15, 16+ observed labs}: n, sum |r|, sum r^2, sum t, sum t^2, sum |r/t| (t != 0), count(t != 0) on the
 * winsorised residuals; patient_idx int64[M], degree int64[N_patient] (bincount of has_lab sources). */
int b2g_eval_per_lab_strata(const float* pred, const float* target, const int32_t* rowptr, const int32_t* pair_of,
                            const int64_t* patient_idx, const int64_t* degree, int n_lab, int winsorize, float n_sigma,
                            double* out, double* out_bins, float* pred_w, void* stream);

/* y = dropout(relu(x)) without normalisation (EdgeRegressionHead, model.py:377-380) and its backward
 * (dx = dy * mask * [y > 0]); in-place allowed. */
int b2g_relu_dropout_fwd(const float* x, int64_t n, int relu, float p_drop, uint64_t seed, uint64_t stream_id,
                         float* y, void* stream);
int b2g_relu_dropout_bwd(const float* y, const float* dy, int64_t n, int relu, float p_drop, uint64_t seed,
                         uint64_t stream_id, float* dx, void* stream);
/* the keep mask alone (0 or 1/(1-p)), for replaying device dropout on the CPU oracle */
int b2g_dropout_mask(int64_t n, float p_drop, uint64_t seed, uint64_t stream_id, float* mask, void* stream);

/* y[r,:] = x[r,:] / max(||x[r,:]||_2, eps); inv_norm[r] saved for backward  (F.normalize, model.py:232) */
int b2g_l2norm_fwd(const float* x, int64_t m, int d, float eps, float* y, float* inv_norm, void* stream);
int b2g_l2norm_bwd(const float* y, const float* dy, const float* inv_norm, int64_t m, int d, float* dx,
                   void* stream);
/* The same with the column sums of dx (dx_colsum[d]) from the same pass: dx is the `dy` of the nn.Linear in front of
 * F.normalize (model.py:101-105,232), whose bias gradient is exactly this sum.  ws: b2g_l2norm_bwd_cs_ws_bytes(d). */
size_t b2g_l2norm_bwd_cs_ws_bytes(int d);
int b2g_l2norm_bwd_cs(const float* y, const float* dy, const float* inv_norm, int64_t m, int d, float* dx, float* dx_colsum,
                      void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (e) fused edge decoder -- EdgeRegressionHead([64,32]) over (patient, lab) pairs, model.py:305-333,373-386
 * ---------------------------------------------------------------------------------------------- */
/* pred[i] = w3 . drop(relu(W2 drop(relu(U[pi[i]] + V[li[i]])) + b2)) + b3
 *   U [N_p,64] = h_patient W1[:, :d]^T, V [N_l,64] = h_lab W1[:, d:]^T + b1 (first layer factorised per node, so the
 *   [M, 2d] concatenation of model.py:319-333 is never built); W2 [32,64], b2 [32], w3 [32], b3 [1].
 * Dropout masks: Philox(seed, sid1) over the flat [M,64] layer-1 output, (seed, sid2) over the flat [M,32] one. */
int b2g_decoder_fwd(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2,
                    const float* b2, const float* w3, const float* b3, int64_t m, float p_drop, uint64_t seed,
                    uint64_t sid1, uint64_t sid2, float* pred, void* stream);
/* Same forward with the 64 -> 32 layer on tcgen05 (tf32 mode): each thread writes its pair's z1 row into a K-major
 * SWIZZLE_128B shared-memory tile, one thread issues the MMAs (M = 128 pairs, N = 32, TF32 operands, fp32 TMEM
 * accumulator), each thread reads its accumulator row back.  Identical dropout streams; ~1e-3 relative. */
int b2g_decoder_fwd_tc(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2,
                       const float* b2, const float* w3, const float* b3, int64_t m, float p_drop, uint64_t seed,
                       uint64_t sid1, uint64_t sid2, float* pred, void* stream);
/* Backward of the above for upstream gradient dpred[M].  Pairs with dpred == 0 (the unsupervised 80 %,
 * train.py:366-368) are compacted away first (stable order -> deterministic sums).  Outputs: g_rows[M,64] = d loss /
 * d (U[p]+V[l]) written ONLY for pairs with dpred != 0; active_flags[M] = 1.0 / 0.0 marks those rows (pass it as
 * col_scale to b2g_gather_reduce* to form dU / dV without touching unwritten rows); dW2 [32,64], db2 [32], dw3 [32],
 * db3 [1].  ws: b2g_decoder_bwd_ws_bytes(m). */
size_t b2g_decoder_bwd_ws_bytes(int64_t m);
int b2g_decoder_bwd(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2,
                    const float* b2, const float* w3, const float* dpred, int64_t m, float p_drop, uint64_t seed,
                    uint64_t sid1, uint64_t sid2, float* g_rows, float* active_flags, float* dW2, float* db2,
                    float* dw3, float* db3, void* ws, size_t ws_bytes, void* stream);

/* tf32-mode backward: per 128-pair tile both contractions (a2 = z1 W2^T, dz1 = da2 W2) run on tcgen05 from thread-written
 * swizzled shared-memory tiles; dW2 is accumulated from the same tiles with fp32 FMAs.  Same arguments / outputs. */
int b2g_decoder_bwd_tc(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2,
                       const float* b2, const float* w3, const float* dpred, int64_t m, float p_drop, uint64_t seed,
                       uint64_t sid1, uint64_t sid2, float* g_rows, float* active_flags, float* dW2, float* db2,
                       float* dw3, float* db3, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (f) loss -- train.py:364-386 weighted MAE / MSE over the supervised subset, model.py:602-605
 * ---------------------------------------------------------------------------------------------- */
/* loss = (1/n_sup) * sum_{i: sup[i]} w[lab[i]] * (|p-t| or (p-t)^2); grad[i] = dloss/dpred[i] (0 where
 * !sup[i]).  sup / w may be NULL (all edges, unit weights).  kind: 0 = mae, 1 = mse, 2 = huber(delta 1).
 * Deterministic tree reduction.  ws: b2g_loss_ws_bytes(m). */
size_t b2g_loss_ws_bytes(int64_t m);
int b2g_weighted_loss(const float* pred, const float* target, const int64_t* lab, const float* w,
                      const uint8_t* sup, int64_t m, int kind, float* loss, float* grad, void* ws,
                      size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2G_H_ */

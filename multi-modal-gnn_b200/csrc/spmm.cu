// Segmented gather-reduce over CSR (the message-passing contraction) and row gathers.
// Replaces SAGEConv.propagate's index_select -> [E,d] message tensor -> scatter_add (PyG; model.py:256):
// no message tensor is ever materialised and no floating-point atomics are used.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace b2g;

struct RelPack {
  b2g_rel_t r[4];
};

// sum_{j in [b,e)} col_scale[col[j]] * x[col[j],:]  accumulated into acc, edges visited in CSR order
template <int D>
__device__ __forceinline__ void reduce_span(const b2g_rel_t& rel, int b, int e, int lane, RowVec<D>& acc) {
  for (int j0 = b; j0 < e; j0 += 32) {
    int j = j0 + lane;
    int my = (j < e) ? __ldg(rel.col + j) : 0;
    float cs = (rel.col_scale != nullptr && j < e) ? __ldg(rel.col_scale + my) : 1.0f;
    int cnt = min(32, e - j0);
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {  // 4 independent row loads in flight
      int c0 = __shfl_sync(FULL, my, t), c1 = __shfl_sync(FULL, my, t + 1), c2 = __shfl_sync(FULL, my, t + 2),
          c3 = __shfl_sync(FULL, my, t + 3);
      float s0 = __shfl_sync(FULL, cs, t), s1 = __shfl_sync(FULL, cs, t + 1), s2 = __shfl_sync(FULL, cs, t + 2),
            s3 = __shfl_sync(FULL, cs, t + 3);
      // a zero column scale means "row not present" (e.g. decoder gradient rows of unsupervised pairs): never read it
      RowVec<D> r0, r1, r2, r3;
      if (s0 != 0.f) r0.load(rel.x + (size_t)c0 * D, lane); else r0.zero();
      if (s1 != 0.f) r1.load(rel.x + (size_t)c1 * D, lane); else r1.zero();
      if (s2 != 0.f) r2.load(rel.x + (size_t)c2 * D, lane); else r2.zero();
      if (s3 != 0.f) r3.load(rel.x + (size_t)c3 * D, lane); else r3.zero();
      acc.fma(s0, r0);
      acc.fma(s1, r1);
      acc.fma(s2, r2);
      acc.fma(s3, r3);
    }
    for (; t < cnt; ++t) {
      int c = __shfl_sync(FULL, my, t);
      float s = __shfl_sync(FULL, cs, t);
      RowVec<D> r;
      if (s != 0.f) r.load(rel.x + (size_t)c * D, lane); else r.zero();
      acc.fma(s, r);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_gather_reduce(RelPack rels, int n_rels, int64_t n_rows, float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  RowVec<D> acc;
  acc.zero();
  for (int k = 0; k < n_rels; ++k) {
    const b2g_rel_t& rel = rels.r[k];
    int b = __ldg(rel.rowptr + row), e = __ldg(rel.rowptr + row + 1);
    RowVec<D> part;
    part.zero();
    reduce_span<D>(rel, b, e, lane, part);
    float rs = rel.row_scale ? __ldg(rel.row_scale + row) : 1.0f;
    acc.fma(rs, part);
  }
  float* o = out + (size_t)row * D;
  if (accumulate) {
    RowVec<D> prev;
    prev.load_rw(o, lane);
    acc.add(prev);
  }
  acc.store(o, lane);
}

// ---- the same contraction with the SOURCE TABLES STAGED IN SHARED MEMORY ---------------------------------------------------
// Patient destinations gather from the lab / diagnosis / medication tables, a few hundred rows each (C4: 50 + 200 + 100 rows
// of 512 B = 175 KB): every CTA brings all of them into shared memory once with bulk copies (cp.async.bulk + mbarrier) and
// stays resident (one CTA per SM, grid-stride over the destination rows); a gathered row is then one conflict-free LDS.128
// per lane instead of a trip through the L1 tags.  Same arithmetic and edge order as k_gather_reduce: identical results.
struct StagedInfo {
  int n_src[4];
  int off[4];          // float offset of table k in shared memory
};

template <int D>
__device__ __forceinline__ void lds_row(RowVec<D>& r, const float* row, int lane) {
  if constexpr (D >= 128) {
#pragma unroll
    for (int j = 0; j < D / 128; ++j) {
      const float4 t = *(reinterpret_cast<const float4*>(row + j * 128) + lane);
      r.v[4 * j] = t.x; r.v[4 * j + 1] = t.y; r.v[4 * j + 2] = t.z; r.v[4 * j + 3] = t.w;
    }
  } else if constexpr (D == 64) {
    const float2 t = *(reinterpret_cast<const float2*>(row) + lane);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = row[lane];
  }
}

// A warp owns 32 consecutive destination rows: their row pointers are one coalesced load per relation, and their column
// indices are one contiguous stream per relation that the warp walks in 32-entry chunks held in registers (current + next
// chunk, the next one always in flight), so the only global-load latency left per row is that of the gathered rows
// themselves (k_gather_reduce: row pointer -> column index -> column scale -> row, four dependent loads per row).
// FMAs are applied in edge order, relation by relation: bit-identical to k_gather_reduce.
// STAGED: the source tables sit in shared memory (`tab`, offsets info.off[k]) and a gathered row is an LDS.128 per lane;
// otherwise rows come from global memory (read-only path).  MAXR: relations compiled in (their stream state lives in registers).
template <int D, int MAXR, bool STAGED>
__device__ __forceinline__ void gather_block32(const RelPack& rels, const StagedInfo& info, const float* tab, int n_rels, int64_t n_rows,
                                               int64_t row0, int lane, float* __restrict__ out, int accumulate) {
  const int nr = (int)min((int64_t)32, n_rows - row0);
  int rp[MAXR], eblk[MAXR], base[MAXR], buf[MAXR], nxt[MAXR], cur[MAXR];
  float rsv[MAXR], sbuf[MAXR], snxt[MAXR];
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    if (k < n_rels) {
      const b2g_rel_t& rel = rels.r[k];
      rp[k] = __ldg(rel.rowptr + row0 + min(lane, nr));
      eblk[k] = __ldg(rel.rowptr + row0 + nr);
      rsv[k] = (rel.row_scale && lane < nr) ? __ldg(rel.row_scale + row0 + lane) : 1.0f;
      base[k] = __shfl_sync(FULL, rp[k], 0);
      cur[k] = 0;
      const int j0 = base[k] + lane, j1 = base[k] + 32 + lane;
      buf[k] = j0 < eblk[k] ? __ldg(rel.col + j0) : 0;
      nxt[k] = j1 < eblk[k] ? __ldg(rel.col + j1) : 0;
      sbuf[k] = (rel.col_scale && j0 < eblk[k]) ? __ldg(rel.col_scale + buf[k]) : 1.0f;
      snxt[k] = (rel.col_scale && j1 < eblk[k]) ? __ldg(rel.col_scale + nxt[k]) : 1.0f;
    }
  }
  for (int r = 0; r < nr; ++r) {
    RowVec<D> acc;
    acc.zero();
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      if (k < n_rels) {
        const b2g_rel_t& rel = rels.r[k];
        const float* src = STAGED ? tab + info.off[k] : rel.x;
        const int b = __shfl_sync(FULL, rp[k], r);
        const int e = r + 1 < 32 ? __shfl_sync(FULL, rp[k], r + 1) : eblk[k];
        RowVec<D> part;
        part.zero();
        for (int j = b; j < e; j += 4) {
          const int idx = j - base[k];
          while ((idx >> 5) > cur[k]) {                  // (uniform) step to the chunk of edge j; keep the one after it in flight
            buf[k] = nxt[k];
            sbuf[k] = snxt[k];
            ++cur[k];
            const int jn = base[k] + (cur[k] + 1) * 32 + lane;
            nxt[k] = jn < eblk[k] ? __ldg(rel.col + jn) : 0;
            snxt[k] = (rel.col_scale && jn < eblk[k]) ? __ldg(rel.col_scale + nxt[k]) : 1.0f;
          }
          const int cnt = min(4, e - j);
          int c[4];
          float sc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int id = idx + u;
            const bool in_next = (id >> 5) > cur[k];     // uniform: a group of 4 touches at most the current and the next chunk
            c[u] = __shfl_sync(FULL, in_next ? nxt[k] : buf[k], id & 31);
            sc[u] = __shfl_sync(FULL, in_next ? snxt[k] : sbuf[k], id & 31);
          }
          RowVec<D> rw[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {                  // (a zero column scale means "row not present", as in k_gather_reduce)
            if (u < cnt && sc[u] != 0.f) {
              if (STAGED) lds_row<D>(rw[u], src + (size_t)c[u] * D, lane);
              else rw[u].load(src + (size_t)c[u] * D, lane);
            } else {
              rw[u].zero();
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < cnt) part.fma(sc[u], rw[u]);
        }
        acc.fma(__shfl_sync(FULL, rsv[k], r), part);
      }
    }
    float* o = out + (size_t)(row0 + r) * D;
    if (accumulate) {
      RowVec<D> prev;
      prev.load_rw(o, lane);
      acc.add(prev);
    }
    acc.store(o, lane);
  }
}

template <int D>
__global__ void __launch_bounds__(512, 1) k_gather_reduce_staged(RelPack rels, StagedInfo info, int n_rels, int64_t n_rows,
                                                                 float* __restrict__ out, int accumulate) {
  extern __shared__ __align__(16) float tab[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int k = 0; k < n_rels; ++k) total += (uint32_t)info.n_src[k] * D * 4u;
    mbar_expect_tx(&bar, total);
    for (int k = 0; k < n_rels; ++k) bulk_load(tab + info.off[k], rels.r[k].x, (uint32_t)info.n_src[k] * D * 4u, &bar);
  }
  mbar_wait(&bar, 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t n_blocks = (n_rows + 31) / 32;
  for (int64_t blk = (int64_t)blockIdx.x * nwarps + warp; blk < n_blocks; blk += (int64_t)gridDim.x * nwarps)
    gather_block32<D, 4, true>(rels, info, tab, n_rels, n_rows, blk * 32, lane, out, accumulate);
}

// the streaming form without staging: 8 warps per CTA, a warp per 32 destination rows
template <int D, int MAXR>
__global__ void __launch_bounds__(256) k_gather_reduce_stream(RelPack rels, int n_rels, int64_t n_rows, float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blk * 32 >= n_rows) return;
  StagedInfo none{};
  gather_block32<D, MAXR, false>(rels, none, nullptr, n_rels, n_rows, blk * 32, lane, out, accumulate);
}

template <int D>
__global__ void __launch_bounds__(256) k_chunk_partial(b2g_rel_t rel, const int32_t* __restrict__ item_row,
                                                       const int32_t* __restrict__ item_start, int64_t n_items, int chunk,
                                                       float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (it >= n_items) return;
  int row = __ldg(item_row + it);
  int b = __ldg(item_start + it);
  int e = min(b + chunk, __ldg(rel.rowptr + row + 1));
  RowVec<D> acc;
  acc.zero();
  reduce_span<D>(rel, b, e, lane, acc);
  acc.store(partial + (size_t)it * D, lane);
}

template <int D>
__global__ void __launch_bounds__(256) k_chunk_final(const float* __restrict__ partial, const int32_t* __restrict__ row_item_ptr,
                                                     const float* __restrict__ row_scale, int64_t n_rows, float* __restrict__ out,
                                                     int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  int b = __ldg(row_item_ptr + row), e = __ldg(row_item_ptr + row + 1);
  RowVec<D> acc;
  acc.zero();
  for (int i = b; i < e; ++i) {
    RowVec<D> p;
    p.load(partial + (size_t)i * D, lane);
    acc.add(p);
  }
  float rs = row_scale ? __ldg(row_scale + row) : 1.0f;
  RowVec<D> res;
  res.zero();
  res.fma(rs, acc);
  float* o = out + (size_t)row * D;
  if (accumulate) {
    RowVec<D> prev;
    prev.load_rw(o, lane);
    res.add(prev);
  }
  res.store(o, lane);
}

template <int D>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t m,
                                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= m) return;
  int64_t r = __ldg(idx + i);
  RowVec<D> v;
  v.load(table + (size_t)r * D, lane);
  v.store(out + (size_t)i * D, lane);
}

template <int D>
__global__ void __launch_bounds__(256) k_gather_add_rows(const float* __restrict__ a, const int64_t* __restrict__ ia,
                                                         const float* __restrict__ b, const int64_t* __restrict__ ib, int64_t m, int relu,
                                                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= m) return;
  RowVec<D> va, vb;
  va.load(a + (size_t)__ldg(ia + i) * D, lane);
  vb.load(b + (size_t)__ldg(ib + i) * D, lane);
  va.add(vb);
  if (relu) {
#pragma unroll
    for (int k = 0; k < RowVec<D>::N; ++k) va.v[k] = fmaxf(va.v[k], 0.f);
  }
  va.store(out + (size_t)i * D, lane);
}


__global__ void k_scatter_values(const float* __restrict__ vals, const int64_t* __restrict__ idx, int64_t m, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[idx[i]] = vals[i];
}
__global__ void k_gather_values(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t m, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = src[idx[i]];
}

// ---- dense adjacency for the tensor-core formulation of the aggregation -------------------------------------------------
// EHR relations are bipartite between millions of patients and a vocabulary of <= 256 labs / diagnoses / drugs, with
// 5-70 % of all possible pairs present: the adjacency is better held as a dense [n_big, pad] fp32 matrix (entries 0/1 or
// 0 / (1/deg_row), all exactly representable or uniformly rounded in TF32) and multiplied on the tensor cores.
__global__ void __launch_bounds__(256) k_dense_adj(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                   const float* __restrict__ row_val, int64_t n_rows, int pad, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  float* o = out + (size_t)row * pad;
  for (int c = lane; c < pad; c += 32) o[c] = 0.f;
  __syncwarp();
  const float v = row_val ? __ldg(row_val + row) : 1.0f;
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (int j = b + lane; j < e; j += 32) atomicAdd(o + __ldg(col + j), v);   // equal addends: order-independent, exact
}

// out[c, r] = (r < rows ? in[r, c] * scale[r] : 0)   in [rows, cols] -> out [cols, pad]
__global__ void k_transpose_pad(const float* __restrict__ in, const float* __restrict__ scale, int rows, int cols, int pad,
                                float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    int r = r0 + i;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] * (scale ? scale[r] : 1.f) : 0.f;
  }
  __syncthreads();
  const int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < cols && r < pad) out[(size_t)(c0 + i) * pad + r] = tile[threadIdx.x][i];
}

__global__ void k_row_scale(const float* __restrict__ in, const float* __restrict__ scale, int64_t n4, int d4, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(in)[i];
  float s = __ldg(scale + i / d4);
  reinterpret_cast<float4*>(out)[i] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
}
}  // namespace

/* Dense [n_rows, pad] adjacency of a CSR: out[r, col[j]] += (row_val ? row_val[r] : 1) for j in row r, zero elsewhere. */
extern "C" int b2g_dense_adjacency(const int32_t* rowptr, const int32_t* col, const float* row_val, int64_t n_rows, int pad, float* out,
                                   void* stream_) {
  B2G_CHECK_ARG(rowptr && col && out && n_rows > 0 && pad > 0, "dense_adjacency: bad args");
  k_dense_adj<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream_>>>(rowptr, col, row_val, n_rows, pad, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

/* out[cols, pad] = (in[rows, cols] * scale[rows, None])^T, zero-padded to `pad` columns (scale may be NULL). */
extern "C" int b2g_transpose_pad(const float* in, const float* scale, int rows, int cols, int pad, float* out, void* stream_) {
  B2G_CHECK_ARG(in && out && rows > 0 && cols > 0 && pad >= rows, "transpose_pad: bad args");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(pad, 32)), block(32, 8);
  k_transpose_pad<<<grid, block, 0, (cudaStream_t)stream_>>>(in, scale, rows, cols, pad, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

/* out[r, :] = in[r, :] * scale[r]   (d % 4 == 0) */
extern "C" int b2g_row_scale(const float* in, const float* scale, int64_t rows, int d, float* out, void* stream_) {
  B2G_CHECK_ARG(in && scale && out && rows > 0 && d > 0 && d % 4 == 0 && aligned16(in) && aligned16(out), "row_scale: bad args");
  int64_t n4 = rows * d / 4;
  k_row_scale<<<(unsigned)ceil_div(n4, 256), 256, 0, (cudaStream_t)stream_>>>(in, scale, n4, d / 4, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

namespace {
}  // namespace

extern "C" int b2g_scatter_values(const float* vals, const int64_t* idx, int64_t m, float* out, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (vals && idx && out)), "scatter_values: bad args");
  if (m == 0) return B2G_OK;
  k_scatter_values<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream_>>>(vals, idx, m, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_gather_values(const float* src, const int64_t* idx, int64_t m, float* out, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (src && idx && out)), "gather_values: bad args");
  if (m == 0) return B2G_OK;
  k_gather_values<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream_>>>(src, idx, m, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

namespace {
int gather_reduce_impl(const b2g_rel_t* h_rels, int n_rels, int64_t n_rows, int d, float* out, int accumulate, bool stream, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(h_rels && n_rels >= 1 && n_rels <= 4 && n_rows >= 0 && out, "gather_reduce: bad args (n_rels=%d)", n_rels);
  B2G_CHECK_ARG(aligned16(out), "gather_reduce: out not 16-byte aligned");
  if (n_rows == 0) return B2G_OK;
  RelPack pack{};
  for (int k = 0; k < n_rels; ++k) {
    B2G_CHECK_ARG(h_rels[k].rowptr && h_rels[k].x && aligned16(h_rels[k].x), "gather_reduce: relation %d has null/unaligned pointers", k);
    pack.r[k] = h_rels[k];
  }
  if (stream) {
    unsigned grid = (unsigned)ceil_div(n_rows, 8 * 32);
    if (n_rels == 1) {
      DISPATCH_D(d, (k_gather_reduce_stream<D, 1><<<grid, 256, 0, st>>>(pack, n_rels, n_rows, out, accumulate)));
    } else {
      DISPATCH_D(d, (k_gather_reduce_stream<D, 4><<<grid, 256, 0, st>>>(pack, n_rels, n_rows, out, accumulate)));
    }
  } else {
    unsigned grid = (unsigned)ceil_div(n_rows, 8);
    DISPATCH_D(d, (k_gather_reduce<D><<<grid, 256, 0, st>>>(pack, n_rels, n_rows, out, accumulate)));
  }
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}
}  // namespace

extern "C" int b2g_gather_reduce(const b2g_rel_t* h_rels, int n_rels, int64_t n_rows, int d, float* out, int accumulate, void* stream_) {
  return gather_reduce_impl(h_rels, n_rels, n_rows, d, out, accumulate, false, stream_);
}
/* same contraction, same results: a warp per 32 destination rows with the CSR streamed through registers -- for MANY SHORT rows
 * (hundreds of thousands of rows of a few entries: the per-patient reduction of the decoder's pair gradients) */
extern "C" int b2g_gather_reduce_stream(const b2g_rel_t* h_rels, int n_rels, int64_t n_rows, int d, float* out, int accumulate, void* stream_) {
  return gather_reduce_impl(h_rels, n_rels, n_rows, d, out, accumulate, true, stream_);
}

/* all source tables in shared memory: 16-byte aligned tables, sum of rows * d * 4 <= 200 KB */
extern "C" int b2g_gather_reduce_staged_supported(const int* h_n_src, int n_rels, int d) {
  if (!h_n_src || n_rels < 1 || n_rels > 4 || (d != 32 && d != 64 && d != 128 && d != 256)) return 0;
  size_t bytes = 0;
  for (int k = 0; k < n_rels; ++k) {
    if (h_n_src[k] < 1) return 0;
    bytes += (size_t)h_n_src[k] * d * 4;
  }
  return bytes <= 200 * 1024 ? 1 : 0;
}

extern "C" int b2g_gather_reduce_staged(const b2g_rel_t* h_rels, const int* h_n_src, int n_rels, int64_t n_rows, int d, float* out,
                                        int accumulate, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(h_rels && out && n_rows >= 0 && b2g_gather_reduce_staged_supported(h_n_src, n_rels, d),
                "gather_reduce_staged: the source tables must fit 200 KB of shared memory (n_rels=%d d=%d)", n_rels, d);
  B2G_CHECK_ARG(aligned16(out), "gather_reduce_staged: out not 16-byte aligned");
  if (n_rows == 0) return B2G_OK;
  RelPack pack{};
  StagedInfo info{};
  int off = 0;
  for (int k = 0; k < n_rels; ++k) {
    B2G_CHECK_ARG(h_rels[k].rowptr && h_rels[k].x && aligned16(h_rels[k].x), "gather_reduce_staged: relation %d has null/unaligned pointers", k);
    pack.r[k] = h_rels[k];
    info.n_src[k] = h_n_src[k];
    info.off[k] = off;
    off += h_n_src[k] * d;
  }
  const size_t smem = (size_t)off * 4;
  int64_t blocks = ceil_div(n_rows, 32 * 16);
  unsigned grid = (unsigned)(blocks < sm_count() ? blocks : sm_count());
  DISPATCH_D(d, ({
               static size_t smem_set = 0;
               if (smem > smem_set) {
                 B2G_CUDA(cudaFuncSetAttribute(k_gather_reduce_staged<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                 smem_set = smem;
               }
               k_gather_reduce_staged<D><<<grid, 512, smem, st>>>(pack, info, n_rels, n_rows, out, accumulate);
             }));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_gather_reduce_chunked(const b2g_rel_t* h_rel, const int32_t* item_row, const int32_t* item_start,
                                         const int32_t* row_item_ptr, int64_t n_items, int32_t chunk, int64_t n_rows, int d,
                                         float* out, int accumulate, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(h_rel && h_rel->rowptr && h_rel->x && item_row && item_start && row_item_ptr && out && chunk > 0,
                "gather_reduce_chunked: null pointer");
  B2G_CHECK_ARG(aligned16(out) && aligned16(h_rel->x) && aligned16(ws), "gather_reduce_chunked: unaligned pointer");
  if (n_rows == 0) return B2G_OK;
  if (ws_bytes < (size_t)n_items * d * sizeof(float)) {
    set_error("gather_reduce_chunked: workspace too small (%zu < %zu)", ws_bytes, (size_t)n_items * d * sizeof(float));
    return B2G_EWS;
  }
  b2g_rel_t rel = *h_rel;
  float* partial = (float*)ws;
  unsigned g1 = (unsigned)ceil_div(n_items, 8), g2 = (unsigned)ceil_div(n_rows, 8);
  DISPATCH_D(d, (k_chunk_partial<D><<<g1, 256, 0, st>>>(rel, item_row, item_start, n_items, chunk, partial)));
  B2G_LAUNCH_CHECK();
  DISPATCH_D(d, (k_chunk_final<D><<<g2, 256, 0, st>>>(partial, row_item_ptr, rel.row_scale, n_rows, out, accumulate)));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_gather_rows(const float* table, const int64_t* idx, int64_t m, int64_t n_table, int d, float* out, void* stream_) {
  (void)n_table;
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (table && idx && out)), "gather_rows: null pointer");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(table) && aligned16(out), "gather_rows: unaligned pointer");
  unsigned grid = (unsigned)ceil_div(m, 8);
  DISPATCH_D(d, (k_gather_rows<D><<<grid, 256, 0, (cudaStream_t)stream_>>>(table, idx, m, out)));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_gather_add_rows(const float* a, const int64_t* ia, const float* b, const int64_t* ib, int64_t m, int d, int relu,
                                   float* out, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (a && ia && b && ib && out)), "gather_add_rows: null pointer");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(out), "gather_add_rows: unaligned pointer");
  unsigned grid = (unsigned)ceil_div(m, 8);
  DISPATCH_D(d, (k_gather_add_rows<D><<<grid, 256, 0, (cudaStream_t)stream_>>>(a, ia, b, ib, m, relu, out)));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

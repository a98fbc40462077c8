// Device side of the graph ingest: the reference builds its COO edge lists with a Python `iterrows` loop
// per table row (graph_build.py:476-586: two NodeIndexer dictionary look-ups per row, rows with an unknown entity dropped).
// Here the per-row work runs on the GPU: a binary search of every row's entity id in the sorted dictionary, then a STABLE
// compaction of the rows whose two look-ups succeeded (edge order = table row order, exactly like the loop), straight into
// the [2, E] int64 edge_index layout the reference emits (graph_build.py:515), the edge attribute riding along.
#include "common.cuh"

namespace {
using namespace b2g;

constexpr int IG_THREADS = 256;
constexpr int IG_IPT = 8;
constexpr int IG_TILE = IG_THREADS * IG_IPT;

// out[i] = index[j] with sorted_ids[j] == query[i], else -1
__global__ void __launch_bounds__(IG_THREADS) k_id_lookup(const int64_t* __restrict__ sorted_ids, const int32_t* __restrict__ index,
                                                         int64_t n_dict, const int64_t* __restrict__ query, int64_t m,
                                                         int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t q = query[i];
  int64_t lo = 0, hi = n_dict;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(sorted_ids + mid) < q) lo = mid + 1; else hi = mid;
  }
  out[i] = (lo < n_dict && __ldg(sorted_ids + lo) == q) ? __ldg(index + lo) : -1;
}

// number of rows with both indices >= 0 per tile of IG_TILE rows
__global__ void __launch_bounds__(IG_THREADS) k_valid_count(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t m,
                                                           int32_t* __restrict__ tile_count) {
  __shared__ int s[IG_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * IG_TILE + (int64_t)threadIdx.x * IG_IPT;
  int c = 0;
#pragma unroll
  for (int k = 0; k < IG_IPT; ++k)
    if (base + k < m && a[base + k] >= 0 && b[base + k] >= 0) ++c;
  c = (int)warp_sum((float)c);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < IG_THREADS / 32; ++w) t += s[w];
    tile_count[blockIdx.x] = t;
  }
}

// exclusive scan of the tile counts (one block; a few thousand tiles per 10 M rows), total -> *total
__global__ void __launch_bounds__(1024) k_tile_scan(int32_t* __restrict__ tile_count, int64_t n_tiles, int64_t* __restrict__ total) {
  __shared__ long long carry_s;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t c0 = 0; c0 < n_tiles; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const int v = i < n_tiles ? tile_count[i] : 0;
    int inc = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      int s = wsum[lane], si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    const long long carry = carry_s;
    if (i < n_tiles) tile_count[i] = (int32_t)(carry + wsum[w] + inc - v);      // E < 2^31 (checked by the caller)
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

// stable scatter: thread t of a tile owns IG_IPT consecutive rows; its offset = tile offset + valid rows of the threads before it
__global__ void __launch_bounds__(IG_THREADS) k_valid_scatter(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                                             const float* __restrict__ attr, int64_t m,
                                                             const int32_t* __restrict__ tile_off, int64_t e_total,
                                                             int64_t* __restrict__ edge_index, float* __restrict__ attr_out,
                                                             int64_t* __restrict__ row_of_edge) {
  __shared__ int wsum[IG_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * IG_TILE + (int64_t)threadIdx.x * IG_IPT;
  int c = 0;
  bool ok[IG_IPT];
#pragma unroll
  for (int k = 0; k < IG_IPT; ++k) {
    ok[k] = base + k < m && a[base + k] >= 0 && b[base + k] >= 0;
    c += ok[k] ? 1 : 0;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int before = 0;
  for (int q = 0; q < w; ++q) before += wsum[q];
  int64_t pos = (int64_t)tile_off[blockIdx.x] + before + inc - c;
#pragma unroll
  for (int k = 0; k < IG_IPT; ++k) {
    if (ok[k]) {
      edge_index[pos] = a[base + k];
      edge_index[e_total + pos] = b[base + k];
      if (attr_out) attr_out[pos] = attr[base + k];
      if (row_of_edge) row_of_edge[pos] = base + k;
      ++pos;
    }
  }
}
}  // namespace

/* out[i] = index[j] where sorted_ids[j] == query[i], or -1: NodeIndexer.get_index (graph_build.py:84-89) for every table row at once.
 * sorted_ids int64[n_dict] ascending and distinct, index int32[n_dict] = the node index of each id. */
extern "C" int b2g_id_lookup(const int64_t* sorted_ids, const int32_t* index, int64_t n_dict, const int64_t* query, int64_t m, int32_t* out,
                             void* stream_) {
  B2G_CHECK_ARG(m >= 0 && n_dict >= 0 && (m == 0 || (query && out)) && (n_dict == 0 || (sorted_ids && index)), "id_lookup: bad args");
  if (m == 0) return B2G_OK;
  k_id_lookup<<<(unsigned)ceil_div(m, IG_THREADS), IG_THREADS, 0, (cudaStream_t)stream_>>>(sorted_ids, index, n_dict, query, m, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_edges_from_rows_ws_bytes(int64_t m) { return (size_t)(ceil_div(m > 0 ? m : 1, IG_TILE) + 1) * sizeof(int32_t) + 256; }

/* SYNC.  Rows i with src_idx[i] >= 0 and dst_idx[i] >= 0 become edges, in row order (the `if patient_idx is not None and lab_idx
 * is not None: edge_list.append(...)` of graph_build.py:502-508,545-551,579-585): edge_index[0, e] = src_idx[i],
 * edge_index[1, e] = dst_idx[i] (int64, [2, *h_n_edges] laid out with row stride *h_n_edges inside a buffer of 2 m entries),
 * attr_out[e] = attr[i] (optional), row_of_edge[e] = i (optional).  ws: b2g_edges_from_rows_ws_bytes(m). */
extern "C" int b2g_edges_from_rows(const int32_t* src_idx, const int32_t* dst_idx, const float* attr, int64_t m, int64_t* edge_index,
                                   float* attr_out, int64_t* row_of_edge, int64_t* h_n_edges, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m >= 0 && h_n_edges && (m == 0 || (src_idx && dst_idx && edge_index)) && (!attr_out || attr), "edges_from_rows: bad args");
  *h_n_edges = 0;
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(m < ((int64_t)1 << 31), "edges_from_rows: more than 2^31 rows");
  if (!ws || ws_bytes < b2g_edges_from_rows_ws_bytes(m)) {
    set_error("edges_from_rows: workspace too small");
    return B2G_EWS;
  }
  const int64_t n_tiles = ceil_div(m, IG_TILE);
  int32_t* tile = (int32_t*)ws;
  int64_t* d_total = nullptr;
  B2G_CUDA(cudaMallocAsync((void**)&d_total, sizeof(int64_t), st));
  k_valid_count<<<(unsigned)n_tiles, IG_THREADS, 0, st>>>(src_idx, dst_idx, m, tile);
  B2G_LAUNCH_CHECK();
  k_tile_scan<<<1, 1024, 0, st>>>(tile, n_tiles, d_total);
  B2G_LAUNCH_CHECK();
  int64_t total = 0;
  B2G_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  B2G_CUDA(cudaStreamSynchronize(st));
  B2G_CUDA(cudaFreeAsync(d_total, st));
  *h_n_edges = total;
  if (total == 0) return B2G_OK;
  k_valid_scatter<<<(unsigned)n_tiles, IG_THREADS, 0, st>>>(src_idx, dst_idx, attr, m, tile, total, edge_index, attr_out, row_of_edge);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

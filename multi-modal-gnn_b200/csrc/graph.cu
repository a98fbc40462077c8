// Graph structure kernels: stable COO -> CSR (hand-written LSD radix sort), degrees, gate, chunk items.
// Replaces the index bookkeeping PyG does per call (index_select / scatter_add on COO) with a one-off
// build: the graph is static for the whole run (SURVEY.md note N5).
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace b2g {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
}  // namespace b2g

extern "C" const char* b2g_last_error(void) { return b2g::g_err; }
extern "C" int b2g_version(void) { return 100; }
extern "C" unsigned long long b2g_launch_count(void) { return b2g::g_launches.load(); }
extern "C" void b2g_reset_launch_count(void) { b2g::g_launches.store(0); }

namespace {
using namespace b2g;

constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 8;                          // items per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;   // 2048 items per block
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;

// key64/val64 -> key32 (+ range validation: graph_build.py:611-633 raises ValueError on the host)
__global__ void k_prepare(const int64_t* __restrict__ key, const int64_t* __restrict__ val, int64_t n, int64_t n_rows,
                          int64_t n_vals, int32_t* __restrict__ key32, int* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = key[i], v = val[i];
  bool bad = (k < 0) | (k >= n_rows) | (v < 0) | (v >= n_vals);
  if (bad) {
    atomicOr(flag, 1);
    k = 0;
  }
  key32[i] = (int32_t)k;
}

__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const int32_t* __restrict__ keys, int64_t n, int shift,
                                                             int32_t* __restrict__ hist, int n_blocks) {
  __shared__ int h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    int64_t idx = base + (int64_t)i * SORT_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (RADIX - 1)], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// block-wide exclusive scan of one int per thread (blockDim.x == 1024 or 256), returns total in `total`
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int& total) {
  __shared__ int wsum[THREADS / 32];
  __shared__ int tot;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = (lane < THREADS / 32) ? wsum[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(FULL, si, o);
      if (lane >= o) si += t;
    }
    if (lane < THREADS / 32) wsum[lane] = si - s;
    if (lane == 31) tot = si;
  }
  __syncthreads();
  int r = wsum[w] + inc - v;
  total = tot;
  __syncthreads();
  return r;
}

// one block per digit row: exclusive scan of hist[d*n_blocks .. (d+1)*n_blocks) in place, total -> digit_total[d]
__global__ void __launch_bounds__(1024) k_scan_rows(int32_t* __restrict__ data, int64_t row_len, int32_t* __restrict__ row_total) {
  int32_t* row = data + (int64_t)blockIdx.x * row_len;
  int carry = 0;
  for (int64_t c0 = 0; c0 < row_len; c0 += 1024) {
    int64_t i = c0 + threadIdx.x;
    int v = (i < row_len) ? row[i] : 0;
    int tot;
    int ex = block_excl_scan<1024>(v, tot);
    if (i < row_len) row[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) row_total[blockIdx.x] = carry;
}

// stable scatter of one radix pass. Tile order is "warp-striped": warp w owns SORT_IPT consecutive groups of 32.
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const int32_t* __restrict__ keys_in,
                                                                const int32_t* __restrict__ eid_in, int64_t n, int shift,
                                                                const int32_t* __restrict__ hist_scanned,
                                                                const int32_t* __restrict__ digit_total, int n_blocks,
                                                                int32_t* __restrict__ keys_out, int32_t* __restrict__ eid_out) {
  __shared__ int warp_cnt[SORT_WARPS][RADIX];
  __shared__ int digit_base[RADIX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&warp_cnt[0][0])[i] = 0;
  {  // exclusive scan of the 256 digit totals (every block recomputes it; 256 adds)
    int tot;
    int ex = block_excl_scan<SORT_THREADS>(digit_total[threadIdx.x], tot);
    digit_base[threadIdx.x] = ex;
  }
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * SORT_TILE + (int64_t)w * (32 * SORT_IPT);
  int32_t k[SORT_IPT], e[SORT_IPT], lr[SORT_IPT];
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    int64_t idx = base + i * 32 + lane;
    bool valid = idx < n;
    k[i] = valid ? keys_in[idx] : 0;
    e[i] = valid ? (eid_in ? eid_in[idx] : (int32_t)idx) : 0;
    int d = valid ? ((k[i] >> shift) & (RADIX - 1)) : RADIX;  // RADIX = "no item"
    unsigned peers = __match_any_sync(FULL, d);
    int rank = __popc(peers & ((1u << lane) - 1u));
    int leader = __ffs(peers) - 1;
    int b = 0;
    if (lane == leader && valid) {
      b = warp_cnt[w][d];
      warp_cnt[w][d] = b + __popc(peers);
    }
    b = __shfl_sync(FULL, b, leader);
    lr[i] = b + rank;
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: running offset over the warps of this block + global offset of (digit, block)
    int d = threadIdx.x;
    int run = digit_base[d] + hist_scanned[(int64_t)d * n_blocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      int t = warp_cnt[ww][d];
      warp_cnt[ww][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_IPT; ++i) {
    int64_t idx = base + i * 32 + lane;
    if (idx < n) {
      int d = (k[i] >> shift) & (RADIX - 1);
      int pos = warp_cnt[w][d] + lr[i];
      keys_out[pos] = k[i];
      eid_out[pos] = e[i];
    }
  }
}

// rowptr from the sorted keys: thread i owns the boundary between sorted[i-1] and sorted[i]
__global__ void k_rowptr_from_sorted(const int32_t* __restrict__ sorted, int64_t n, int64_t n_rows, int32_t* __restrict__ rowptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t prev = (i == 0) ? -1 : (int64_t)sorted[i - 1];
  int64_t cur = (i == n) ? n_rows : (int64_t)sorted[i];
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)i;
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_col_from_eid(const int64_t* __restrict__ val, const int32_t* __restrict__ eid, int64_t n, int64_t n_vals,
                               int32_t* __restrict__ col) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = val[eid[i]];
  col[i] = (v < 0 || v >= n_vals) ? 0 : (int32_t)v;
}

__global__ void k_degrees(const int32_t* __restrict__ rowptr, int64_t n_rows, int64_t* __restrict__ deg, float* __restrict__ inv) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int d = rowptr[r + 1] - rowptr[r];
  if (deg) deg[r] = (int64_t)d;
  if (inv) inv[r] = 1.0f / (float)(d > 1 ? d : 1);
}

__global__ void k_gate(const int64_t* __restrict__ deg, const int64_t* __restrict__ pidx, int64_t m, int64_t thr, uint8_t* __restrict__ low) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) low[i] = deg[pidx[i]] < thr ? 1 : 0;
}

__global__ void k_chunk_counts(const int32_t* __restrict__ rowptr, int64_t n_rows, int chunk, int32_t* __restrict__ cnt) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int d = rowptr[r + 1] - rowptr[r];
  int c = (d + chunk - 1) / chunk;
  cnt[r] = c > 0 ? c : 1;
}

// single-block exclusive scan producing n+1 outputs (out[n] = total)
__global__ void __launch_bounds__(1024) k_scan_excl(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  int carry = 0;
  for (int64_t c0 = 0; c0 < n; c0 += 1024) {
    int64_t i = c0 + threadIdx.x;
    int v = (i < n) ? in[i] : 0;
    int tot;
    int ex = block_excl_scan<1024>(v, tot);
    if (i < n) out[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_chunk_fill(const int32_t* __restrict__ rowptr, int64_t n_rows, int chunk, const int32_t* __restrict__ row_item_ptr,
                             int32_t* __restrict__ item_row, int32_t* __restrict__ item_start) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int b = row_item_ptr[r], e = row_item_ptr[r + 1];
  int s = rowptr[r];
  for (int i = b; i < e; ++i) {
    item_row[i] = (int32_t)r;
    item_start[i] = s + (i - b) * chunk;
  }
}

inline int n_blocks_for(int64_t n) { return (int)ceil_div(n > 0 ? n : 1, SORT_TILE); }
inline int key_bits(int64_t n_rows) {
  int bits = 1;
  while (((int64_t)1 << bits) < n_rows) ++bits;
  return bits;
}
}  // namespace

extern "C" size_t b2g_csr_build_ws_bytes(int64_t n_edges, int64_t n_rows) {
  (void)n_rows;
  size_t e = (size_t)(n_edges > 0 ? n_edges : 1);
  size_t nb = (size_t)n_blocks_for(n_edges);
  return align_up(e * 4, 256) * 3 + align_up(nb * RADIX * 4, 256) + align_up(RADIX * 4, 256) + 256;
}

extern "C" int b2g_csr_build(const int64_t* key, const int64_t* val, int64_t n_edges, int64_t n_rows, int64_t n_vals,
                             int32_t* rowptr, int32_t* col, int32_t* eid, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(n_edges >= 0 && n_rows > 0 && n_vals > 0, "csr_build: bad sizes E=%lld rows=%lld vals=%lld", (long long)n_edges,
                (long long)n_rows, (long long)n_vals);
  B2G_CHECK_ARG(n_edges < (int64_t)2147483647 - SORT_TILE && n_rows < 2147483647LL && n_vals < 2147483647LL,
                "csr_build: int32 index range exceeded");
  B2G_CHECK_ARG(rowptr && (n_edges == 0 || (key && val && col && eid)), "csr_build: null pointer");
  if (ws_bytes < b2g_csr_build_ws_bytes(n_edges, n_rows) || !ws) {
    set_error("csr_build: workspace too small");
    return B2G_EWS;
  }
  if (n_edges == 0) {
    k_fill_i32<<<(unsigned)ceil_div(n_rows + 1, 256), 256, 0, st>>>(rowptr, n_rows + 1, 0);
    B2G_LAUNCH_CHECK();
    return B2G_OK;
  }
  const int nb = n_blocks_for(n_edges);
  char* p = (char*)ws;
  size_t eb = align_up((size_t)n_edges * 4, 256);
  int32_t* kbuf[2] = {(int32_t*)p, (int32_t*)(p + eb)};
  int32_t* ebuf_ws = (int32_t*)(p + 2 * eb);
  int32_t* hist = (int32_t*)(p + 3 * eb);
  int32_t* dtot = (int32_t*)((char*)hist + align_up((size_t)nb * RADIX * 4, 256));
  int* flag = (int*)((char*)dtot + align_up(RADIX * 4, 256));
  B2G_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  k_prepare<<<(unsigned)ceil_div(n_edges, 256), 256, 0, st>>>(key, val, n_edges, n_rows, n_vals, kbuf[0], flag);
  B2G_LAUNCH_CHECK();

  const int passes = (key_bits(n_rows) + 7) / 8;
  int32_t* ebufs[2] = {eid, ebuf_ws};
  const int32_t* e_in = nullptr;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = 8 * pass;
    int32_t* k_in = kbuf[pass & 1];
    int32_t* k_out = kbuf[(pass + 1) & 1];
    int32_t* e_out = ebufs[(passes - 1 - pass) & 1];
    k_radix_hist<<<nb, SORT_THREADS, 0, st>>>(k_in, n_edges, shift, hist, nb);
    B2G_LAUNCH_CHECK();
    k_scan_rows<<<RADIX, 1024, 0, st>>>(hist, nb, dtot);
    B2G_LAUNCH_CHECK();
    k_radix_scatter<<<nb, SORT_THREADS, 0, st>>>(k_in, e_in, n_edges, shift, hist, dtot, nb, k_out, e_out);
    B2G_LAUNCH_CHECK();
    e_in = e_out;
  }
  const int32_t* sorted = kbuf[passes & 1];
  k_rowptr_from_sorted<<<(unsigned)ceil_div(n_edges + 1, 256), 256, 0, st>>>(sorted, n_edges, n_rows, rowptr);
  B2G_LAUNCH_CHECK();
  k_col_from_eid<<<(unsigned)ceil_div(n_edges, 256), 256, 0, st>>>(val, eid, n_edges, n_vals, col);
  B2G_LAUNCH_CHECK();
  int h_flag = 0;
  B2G_CUDA(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2G_CUDA(cudaStreamSynchronize(st));
  if (h_flag) {
    set_error("csr_build: edge endpoint out of range (rows=%lld, vals=%lld)", (long long)n_rows, (long long)n_vals);
    return B2G_ERANGE;
  }
  return B2G_OK;
}

extern "C" int b2g_csr_degrees(const int32_t* rowptr, int64_t n_rows, int64_t* deg, float* inv, void* stream_) {
  B2G_CHECK_ARG(rowptr && n_rows >= 0, "csr_degrees: bad args");
  if (n_rows == 0) return B2G_OK;
  k_degrees<<<(unsigned)ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream_>>>(rowptr, n_rows, deg, inv);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_degree_gate(const int64_t* deg, const int64_t* patient_idx, int64_t m, int64_t threshold, uint8_t* low,
                               void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (deg && patient_idx && low)), "degree_gate: bad args");
  if (m == 0) return B2G_OK;
  k_gate<<<(unsigned)ceil_div(m, 256), 256, 0, (cudaStream_t)stream_>>>(deg, patient_idx, m, threshold, low);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_csr_chunk_ws_bytes(int64_t n_rows) { return align_up((size_t)(n_rows + 1) * 4, 256); }

extern "C" int b2g_csr_chunk_count(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int32_t* row_item_ptr,
                                   int64_t* h_n_items, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(rowptr && row_item_ptr && h_n_items && n_rows > 0 && chunk > 0, "csr_chunk_count: bad args");
  if (!ws || ws_bytes < b2g_csr_chunk_ws_bytes(n_rows)) {
    set_error("csr_chunk_count: workspace too small");
    return B2G_EWS;
  }
  int32_t* cnt = (int32_t*)ws;
  k_chunk_counts<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(rowptr, n_rows, chunk, cnt);
  B2G_LAUNCH_CHECK();
  k_scan_excl<<<1, 1024, 0, st>>>(cnt, n_rows, row_item_ptr);
  B2G_LAUNCH_CHECK();
  int32_t total = 0;
  B2G_CUDA(cudaMemcpyAsync(&total, row_item_ptr + n_rows, 4, cudaMemcpyDeviceToHost, st));
  B2G_CUDA(cudaStreamSynchronize(st));
  *h_n_items = total;
  return B2G_OK;
}

extern "C" int b2g_csr_chunk_fill(const int32_t* rowptr, int64_t n_rows, int32_t chunk, const int32_t* row_item_ptr,
                                  int32_t* item_row, int32_t* item_start, void* stream_) {
  B2G_CHECK_ARG(rowptr && row_item_ptr && item_row && item_start && n_rows > 0 && chunk > 0, "csr_chunk_fill: bad args");
  k_chunk_fill<<<(unsigned)ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream_>>>(rowptr, n_rows, chunk, row_item_ptr, item_row,
                                                                                  item_start);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

// fp32 SIMT dense products for nn.Linear / SAGEConv.lin_l / lin_r / EdgeRegressionHead (model.py:93-103,373-386).
// This is the exact-fp32 ("parity") path; the bf16 tcgen05 path lives in dense_tc.cu.
#include "common.cuh"

namespace {
using namespace b2g;

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;

// C[M,N] = (acc ? C : 0) + A[M,K] * B + bias,   B(k,n) = B_NK ? Bp[n*K + k] : Bp[k*N + n]
template <bool B_NK>
__global__ void __launch_bounds__(GEMM_THREADS) k_sgemm(const float* __restrict__ A, const float* __restrict__ Bp,
                                                        const float* __restrict__ bias, int64_t M, int N, int K,
                                                        float* __restrict__ C, int accumulate, int vec_ok) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread tile 8 (m) x 4 (n)
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile: 128 rows x 16 k, two float4 (along k) per thread, stored transposed
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = (tid >> 2) + h * 64;
      int kq = (tid & 3) * 4;
      int64_t gm = m0 + r;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < M) {
        const float* src = A + (size_t)gm * K + k0 + kq;
        if (vec_ok && k0 + kq + 3 < K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k0 + kq + q < K) v[q] = __ldg(src + q);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) As[kq + q][r] = v[q];
    }
    // ---- B tile: 16 k x 64 n
    if (B_NK) {
      int n = tid >> 2, kq = (tid & 3) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n0 + n < N) {
        const float* src = Bp + (size_t)(n0 + n) * K + k0 + kq;
        if (vec_ok && k0 + kq + 3 < K) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (k0 + kq + q < K) v[q] = __ldg(src + q);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[kq + q][n] = v[q];
    } else {
      int k = tid >> 4, nq = (tid & 15) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + k < K) {
        const float* src = Bp + (size_t)(k0 + k) * N + n0 + nq;
        if (vec_ok && n0 + nq + 3 < N) {
          float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (n0 + nq + q < N) v[q] = __ldg(src + q);
        }
      }
      *reinterpret_cast<float4*>(&Bs[k][nq]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
  const int n = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < N) bv[j] = __ldg(bias + n + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
    float* dst = C + (size_t)gm * N + n;
    if ((N & 3) == 0 && n + 3 < N) {
      float4 o = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
      if (accumulate) {
        float4 p = *reinterpret_cast<const float4*>(dst);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      *reinterpret_cast<float4*>(dst) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) dst[j] = (accumulate ? dst[j] : 0.f) + acc[i][j] + bv[j];
    }
  }
}

// ---- small problems (type-node rows: a few hundred rows at most) ------------------------------------------------------
// C[M,N] = (acc ? C : 0) + opA(A) opB(B) + bias with 32 x 32 output tiles so that even a 50-row problem spreads over
// many CTAs, BK = 32 (4 trips for K = 128) and register prefetch of the next tile: these launches are latency-bound.
//   A_T  : A is stored [K, M] row-major (A(m,k) = Ap[k*M + m])   -- used for dW = dy^T x
//   B_NK : B is stored [N, K] row-major (B(k,n) = Bp[n*K + k])   -- nn.Linear weights
template <bool A_T, bool B_NK>
__global__ void __launch_bounds__(256) k_sgemm_small(const float* __restrict__ Ap, const float* __restrict__ Bp,
                                                     const float* __restrict__ bias, int M, int N, int K, float* __restrict__ C,
                                                     int accumulate) {
  constexpr int KT = 128;          // the whole reduction of a d = 128 layer in ONE load phase: these launches are latency-bound
  __shared__ float As[KT][33];     // [k][m]
  __shared__ float Bs[KT][33];     // [k][n]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int lr = tid >> 3, lq = (tid & 7) * 4;     // loader coordinates: row 0..31, 4 consecutive elements (+ 32 * j)
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += KT) {
    float ra[16], rb[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (A_T) {          // A stored [K, M]: row = k, consecutive m
          int k = k0 + lr + 32 * j, m = m0 + lq + q;
          ra[4 * j + q] = (k < K && m < M) ? __ldg(Ap + (size_t)k * M + m) : 0.f;
        } else {            // A stored [M, K]: row = m, consecutive k
          int m = m0 + lr, k = k0 + lq + q + 32 * j;
          ra[4 * j + q] = (m < M && k < K) ? __ldg(Ap + (size_t)m * K + k) : 0.f;
        }
        if (B_NK) {         // B stored [N, K]: row = n, consecutive k
          int n = n0 + lr, k = k0 + lq + q + 32 * j;
          rb[4 * j + q] = (n < N && k < K) ? __ldg(Bp + (size_t)n * K + k) : 0.f;
        } else {            // B stored [K, N]: row = k, consecutive n
          int k = k0 + lr + 32 * j, n = n0 + lq + q;
          rb[4 * j + q] = (k < K && n < N) ? __ldg(Bp + (size_t)k * N + n) : 0.f;
        }
      }
    }
    __syncthreads();       // previous chunk fully consumed
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (A_T) As[lr + 32 * j][lq + q] = ra[4 * j + q]; else As[lq + q + 32 * j][lr] = ra[4 * j + q];
        if (B_NK) Bs[lq + q + 32 * j][lr] = rb[4 * j + q]; else Bs[lr + 32 * j][lq + q] = rb[4 * j + q];
      }
    }
    __syncthreads();
    const int kmax = min(KT, K - k0);
#pragma unroll 8
    for (int k = 0; k < kmax; ++k) {
      float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
      float b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int n = n0 + tx * 2 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      float* dst = C + (size_t)m * N + n;
      *dst = accumulate ? *dst + v : v;
    }
  }
}
constexpr int SMALL_M = 1024;

// ---- grouped small problems ---------------------------------------------------------------------------------------------
// The type-node side of a HeteroConv layer is a swarm of tiny GEMMs (<= a few hundred rows, K = N = 128): three relations
// x {lin_l on the sources, lin_r on the destinations, lin_l on the aggregates} forward, and twice that backward -- ~50
// launches of ~5 us per step.  One launch takes a list of independent problems
//      C[M,N] = (acc ? C : 0) + opA(A) opB(B) (+ opA(A2) opB(B2)) + bias        (the optional second product shares C)
// and spreads their 32 x 32 tiles over one grid.  Same tile code and operand conventions as k_sgemm_small (A_T / B_NK
// are per-problem flags here), so each problem's result is bit-identical to the single-problem kernel's.
constexpr int GROUP_MAX = 12;
struct GroupPack {
  b2g_gemm_problem_t p[GROUP_MAX];
  int tile_begin[GROUP_MAX + 1];
  int n;
};

__global__ void __launch_bounds__(256) k_sgemm_group(GroupPack g) {
  constexpr int KT = 128;
  __shared__ float As[KT][33];     // [k][m]
  __shared__ float Bs[KT][33];     // [k][n]
  int pi = 0;
  while (pi + 1 < g.n && (int)blockIdx.x >= g.tile_begin[pi + 1]) ++pi;
  const b2g_gemm_problem_t& P = g.p[pi];
  const int M = P.m, N = P.n;
  const int tiles_n = (N + 31) / 32;
  const int t = blockIdx.x - g.tile_begin[pi];
  const int m0 = (t / tiles_n) * 32, n0 = (t % tiles_n) * 32;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 3, lq = (tid & 7) * 4;
  const bool A_T = P.a_transposed != 0, B_NK = P.b_is_nk != 0;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int src = 0; src < 2; ++src) {
    const float* __restrict__ Ap = src == 0 ? P.a : P.a2;
    const float* __restrict__ Bp = src == 0 ? P.b : P.b2;
    const int K = src == 0 ? P.k : P.k2;
    if (Ap == nullptr || K <= 0) continue;
    for (int k0 = 0; k0 < K; k0 += KT) {
      float ra[16], rb[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (A_T) {
            int k = k0 + lr + 32 * j, m = m0 + lq + q;
            ra[4 * j + q] = (k < K && m < M) ? __ldg(Ap + (size_t)k * M + m) : 0.f;
          } else {
            int m = m0 + lr, k = k0 + lq + q + 32 * j;
            ra[4 * j + q] = (m < M && k < K) ? __ldg(Ap + (size_t)m * K + k) : 0.f;
          }
          if (B_NK) {
            int n = n0 + lr, k = k0 + lq + q + 32 * j;
            rb[4 * j + q] = (n < N && k < K) ? __ldg(Bp + (size_t)n * K + k) : 0.f;
          } else {
            int k = k0 + lr + 32 * j, n = n0 + lq + q;
            rb[4 * j + q] = (k < K && n < N) ? __ldg(Bp + (size_t)k * N + n) : 0.f;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (A_T) As[lr + 32 * j][lq + q] = ra[4 * j + q]; else As[lq + q + 32 * j][lr] = ra[4 * j + q];
          if (B_NK) Bs[lq + q + 32 * j][lr] = rb[4 * j + q]; else Bs[lr + 32 * j][lq + q] = rb[4 * j + q];
        }
      }
      __syncthreads();
      const int kmax = min(KT, K - k0);
#pragma unroll 8
      for (int k = 0; k < kmax; ++k) {
        float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
        float b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
        acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
        acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int n = n0 + tx * 2 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (P.bias ? __ldg(P.bias + n) : 0.f);
      float* dst = P.c + (size_t)m * N + n;
      *dst = P.accumulate ? *dst + v : v;
    }
  }
}

// column sums of several small matrices in one launch (bias gradients of the grouped problems): 32 columns per CTA
struct ColsumPack {
  const float* x[GROUP_MAX];
  float* out[GROUP_MAX];
  int m[GROUP_MAX], n[GROUP_MAX];
  int cta_begin[GROUP_MAX + 1];
  int count;
};
__global__ void __launch_bounds__(256) k_colsum_group(ColsumPack g) {
  __shared__ float sh[8][32];
  int pi = 0;
  while (pi + 1 < g.count && (int)blockIdx.x >= g.cta_begin[pi + 1]) ++pi;
  const float* __restrict__ dy = g.x[pi];
  const int M = g.m[pi], N = g.n[pi];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = (blockIdx.x - g.cta_begin[pi]) * 32 + lane;
  float s = 0.f;
  if (n < N) {
#pragma unroll 8
    for (int m = slice; m < M; m += 8) s += __ldg(dy + (size_t)m * N + n);
  }
  sh[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || n >= N) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) s += sh[k][lane];
  g.out[pi][n] = s;
}


// db[n] = sum_m dy[m, n] for a few hundred rows: 32 columns per CTA, 8 row slices (one per warp, coalesced 128-byte row
// segments, 8 loads in flight), slices combined in fixed order: deterministic
__global__ void __launch_bounds__(256) k_colsum_small(const float* __restrict__ dy, int M, int N, float* __restrict__ db) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (n < N) {
#pragma unroll 8
    for (int m = slice; m < M; m += 8) s += __ldg(dy + (size_t)m * N + n);
  }
  sh[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || n >= N) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) s += sh[k][lane];
  db[n] = s;
}

// ---- weight gradient: dW[N,K] = dy[M,N]^T x[M,K], split over M -------------------------------------------------
constexpr int WG_T = 64;      // output tile 64 (n) x 64 (k)
constexpr int WG_MC = 16;     // m rows per smem stage

__global__ void __launch_bounds__(256) k_wgrad_partial(const float* __restrict__ dy, const float* __restrict__ x, int64_t M, int N, int K,
                                                       int64_t rows_per_slice, float* __restrict__ part_w, float* __restrict__ part_b) {
  __shared__ __align__(16) float Ys[WG_MC][WG_T + 4];
  __shared__ __align__(16) float Xs[WG_MC][WG_T + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // thread tile 4 (n) x 4 (k)
  const int n0 = blockIdx.y * WG_T, k0 = blockIdx.z * WG_T;
  const int64_t ms = (int64_t)blockIdx.x * rows_per_slice;
  const int64_t me = min(M, ms + rows_per_slice);
  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t mb = ms; mb < me; mb += WG_MC) {
    {  // 16 x 64 floats each = 256 float4 -> one per thread and matrix
      int r = tid >> 4, c = (tid & 15) * 4;
      int64_t gm = mb + r;
      float vy[4] = {0.f, 0.f, 0.f, 0.f}, vx[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < me) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (n0 + c + q < N) vy[q] = __ldg(dy + (size_t)gm * N + n0 + c + q);
          if (k0 + c + q < K) vx[q] = __ldg(x + (size_t)gm * K + k0 + c + q);
        }
      }
      *reinterpret_cast<float4*>(&Ys[r][c]) = make_float4(vy[0], vy[1], vy[2], vy[3]);
      *reinterpret_cast<float4*>(&Xs[r][c]) = make_float4(vx[0], vx[1], vx[2], vx[3]);
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WG_MC; ++mm) {
      float4 a = *reinterpret_cast<const float4*>(&Ys[mm][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Xs[mm][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* pw = part_w + (size_t)blockIdx.x * N * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < K) pw[(size_t)n * K + k] = acc[i][j];
    }
    if (part_b && blockIdx.z == 0 && tx == 0) part_b[(size_t)blockIdx.x * N + n] = bsum[i];
  }
}

// out[i] = sum over slices of part[slice][i]: one warp per output, lanes stride over the slices, fixed-order butterfly
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ part, int slices, int64_t count, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= count) return;
  float s = 0.f;
  for (int sl = lane; sl < slices; sl += 32) s += part[(size_t)sl * count + i];
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

inline int wgrad_slices(int64_t m, int n, int k) {
  int tiles = (int)(ceil_div(n, WG_T) * ceil_div(k, WG_T));
  int64_t by_rows = ceil_div(m, 1024);
  int64_t by_sms = (4LL * 148 + tiles - 1) / tiles;
  int64_t s = by_rows < by_sms ? by_rows : by_sms;
  return (int)(s < 1 ? 1 : s);
}
}  // namespace

extern "C" int b2g_linear_fwd(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y, int accumulate,
                              void* stream_) {
  B2G_CHECK_ARG(m >= 0 && n > 0 && k > 0 && (m == 0 || (x && w && y)), "linear_fwd: bad args m=%lld n=%d k=%d", (long long)m, n, k);
  if (m == 0) return B2G_OK;
  if (m <= SMALL_M) {
    dim3 g((unsigned)ceil_div(m, 32), (unsigned)ceil_div(n, 32));
    k_sgemm_small<false, true><<<g, 256, 0, (cudaStream_t)stream_>>>(x, w, bias, (int)m, n, k, y, accumulate);
    B2G_LAUNCH_CHECK();
    return B2G_OK;
  }
  int vec_ok = (k % 4 == 0) && aligned16(x) && aligned16(w);
  B2G_CHECK_ARG((n % 4 != 0) || aligned16(y), "linear_fwd: y not 16-byte aligned");
  dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(n, BN));
  k_sgemm<true><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream_>>>(x, w, bias, m, n, k, y, accumulate, vec_ok);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_linear_bwd_input(const float* dy, const float* w, int64_t m, int n, int k, float* dx, int accumulate, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && n > 0 && k > 0 && (m == 0 || (dy && w && dx)), "linear_bwd_input: bad args");
  if (m == 0) return B2G_OK;
  if (m <= SMALL_M) {
    dim3 g((unsigned)ceil_div(m, 32), (unsigned)ceil_div(k, 32));
    k_sgemm_small<false, false><<<g, 256, 0, (cudaStream_t)stream_>>>(dy, w, nullptr, (int)m, k, n, dx, accumulate);
    B2G_LAUNCH_CHECK();
    return B2G_OK;
  }
  // dx[M,K] = dy[M,N] * W[N,K]: reduction dim is N, output width K, B stored [N(k-dim), K(n-dim)] row-major
  int vec_ok = (n % 4 == 0) && (k % 4 == 0) && aligned16(dy) && aligned16(w);
  B2G_CHECK_ARG((k % 4 != 0) || aligned16(dx), "linear_bwd_input: dx not 16-byte aligned");
  dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(k, BN));
  k_sgemm<false><<<grid, GEMM_THREADS, 0, (cudaStream_t)stream_>>>(dy, w, nullptr, m, k, n, dx, accumulate, vec_ok);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_linear_bwd_weight_ws_bytes(int64_t m, int n, int k) {
  int s = wgrad_slices(m > 0 ? m : 1, n, k);
  return align_up((size_t)s * n * k * 4, 256) + align_up((size_t)s * n * 4, 256);
}

extern "C" int b2g_linear_bwd_weight(const float* dy, const float* x, int64_t m, int n, int k, float* dw, float* db, void* ws,
                                     size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m >= 0 && n > 0 && k > 0 && dw && (m == 0 || (dy && x)), "linear_bwd_weight: bad args");
  if (m == 0) {
    B2G_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * k * 4, st));
    if (db) B2G_CUDA(cudaMemsetAsync(db, 0, (size_t)n * 4, st));
    return B2G_OK;
  }
  if (m <= SMALL_M) {
    dim3 g((unsigned)ceil_div(n, 32), (unsigned)ceil_div(k, 32));
    k_sgemm_small<true, false><<<g, 256, 0, st>>>(dy, x, nullptr, n, k, (int)m, dw, 0);
    B2G_LAUNCH_CHECK();
    if (db) {
      k_colsum_small<<<(unsigned)ceil_div(n, 32), 256, 0, st>>>(dy, (int)m, n, db);
      B2G_LAUNCH_CHECK();
    }
    return B2G_OK;
  }
  if (!ws || ws_bytes < b2g_linear_bwd_weight_ws_bytes(m, n, k)) {
    set_error("linear_bwd_weight: workspace too small");
    return B2G_EWS;
  }
  const int slices = wgrad_slices(m, n, k);
  const int64_t rows = ceil_div(ceil_div(m, slices), WG_MC) * WG_MC;
  float* part_w = (float*)ws;
  float* part_b = (float*)((char*)ws + align_up((size_t)slices * n * k * 4, 256));
  dim3 grid((unsigned)slices, (unsigned)ceil_div(n, WG_T), (unsigned)ceil_div(k, WG_T));
  k_wgrad_partial<<<grid, 256, 0, st>>>(dy, x, m, n, k, rows, part_w, db ? part_b : nullptr);
  B2G_LAUNCH_CHECK();
  int64_t cnt = (int64_t)n * k;
  k_wgrad_reduce<<<(unsigned)ceil_div(cnt, 8), 256, 0, st>>>(part_w, slices, cnt, dw);
  B2G_LAUNCH_CHECK();
  if (db) {
    k_wgrad_reduce<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(part_b, slices, n, db);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}


/* Up to 12 independent small problems C = (acc ? C : 0) + opA(A) opB(B) (+ opA(A2) opB(B2)) + bias in ONE launch
 * (h_probs is a HOST array; every m <= 1024).  Operand conventions of b2g_gemm_problem_t are those of the three small
 * linear entry points: forward (a = x [m,k], b = W [n,k], b_is_nk = 1), input gradient (a = dy [m,k], b = W [k,n],
 * b_is_nk = 0), weight gradient (a = dy stored [k,m], a_transposed = 1, b = x [k,n], b_is_nk = 0). */
extern "C" int b2g_small_gemm_group(const b2g_gemm_problem_t* h_probs, int n_probs, void* stream_) {
  B2G_CHECK_ARG(n_probs >= 0 && n_probs <= GROUP_MAX && (n_probs == 0 || h_probs), "small_gemm_group: at most %d problems", GROUP_MAX);
  GroupPack g{};
  int tiles = 0, cnt = 0;
  for (int i = 0; i < n_probs; ++i) {
    const b2g_gemm_problem_t& P = h_probs[i];
    B2G_CHECK_ARG(P.m >= 0 && P.n > 0 && P.k > 0 && P.m <= SMALL_M && P.c && P.a && P.b && (P.a2 == nullptr || (P.b2 && P.k2 > 0)),
                  "small_gemm_group: bad problem %d (m=%d n=%d k=%d)", i, P.m, P.n, P.k);
    if (P.m == 0) continue;
    g.p[cnt] = P;
    g.tile_begin[cnt] = tiles;
    tiles += (int)(ceil_div(P.m, 32) * ceil_div(P.n, 32));
    ++cnt;
  }
  g.tile_begin[cnt] = tiles;
  g.n = cnt;
  if (tiles == 0) return B2G_OK;
  k_sgemm_group<<<(unsigned)tiles, 256, 0, (cudaStream_t)stream_>>>(g);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

/* out_i[n] = sum_m x_i[m, n] for up to 12 small matrices (m <= 1024) in one launch; h_* are HOST arrays. */
extern "C" int b2g_small_colsum_group(const float* const* h_x, float* const* h_out, const int* h_m, const int* h_n, int count,
                                      void* stream_) {
  B2G_CHECK_ARG(count >= 0 && count <= GROUP_MAX && (count == 0 || (h_x && h_out && h_m && h_n)), "small_colsum_group: bad args");
  ColsumPack g{};
  int ctas = 0, c = 0;
  for (int i = 0; i < count; ++i) {
    B2G_CHECK_ARG(h_x[i] && h_out[i] && h_m[i] >= 0 && h_n[i] > 0, "small_colsum_group: bad entry %d", i);
    g.x[c] = h_x[i]; g.out[c] = h_out[i]; g.m[c] = h_m[i]; g.n[c] = h_n[i];
    g.cta_begin[c] = ctas;
    ctas += (int)ceil_div(h_n[i], 32);
    ++c;
  }
  g.cta_begin[c] = ctas;
  g.count = c;
  if (ctas == 0) return B2G_OK;
  k_colsum_group<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream_>>>(g);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

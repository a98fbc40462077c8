// BatchNorm1d (+ReLU +Dropout) forward/backward, row L2 normalisation, ReLU/Dropout, weighted loss.
// Reference ops: nn.BatchNorm1d / F.relu / F.dropout / F.normalize (model.py:93-105,134-139,259-269),
// weighted MAE/MSE (train.py:364-386), compute_regression_loss (model.py:579-612).
// All reductions are two-stage and order-fixed (deterministic); column statistics accumulate in fp64.
#include <atomic>
#include "peer.cuh"

namespace {
using namespace b2g;

constexpr int MAX_PARTIALS = 296;  // upper bound on the CTAs of a column reduction (workspace sizing)
constexpr int COL_THREADS = 512;

// activation codes (model.py:145-153): 0 none, 1 relu, 2 leaky_relu(0.01), 3 elu(alpha 1)
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.01f * v;
  if (act == 3) return v > 0.f ? v : expm1f(v);
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : 0.01f;
  if (act == 3) return pre > 0.f ? 1.f : expf(pre);
  return 1.f;
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- column reductions ------------------------------------------------------------------------------------------
// One kernel per reduction: every CTA reduces its rows to a per-CTA fp64 partial (fixed order), takes a ticket, and the
// CTA that draws the last ticket adds the partials in CTA order and applies the finalisation (so the result does not
// depend on which CTA finishes last: deterministic), all in the same launch.
//   MODE 0: s0 = sum x,        s1 = sum x^2                    (batch statistics, column sums)
//   MODE 1: s0 = sum g,        s1 = sum g * xhat                (BN backward), g = dy * dropmask * act'
//   FIN  0: mean / rstd (+ running-stat update like nn.BatchNorm1d in training mode)
//        1: sums[2d] (double) + dbeta / dgamma     2: out[c] = (float)s0     3: sums[2d] (double) only
struct ColFin {
  int kind;
  int64_t m;
  float eps, momentum;
  float *mean, *rstd, *running_mean, *running_var;   // kind 0
  double* sums;                                       // kind 1, 3
  float *dgamma, *dbeta;                              // kind 1
  float* out;                                         // kind 2
  int peer_on;                                        // multi-GPU: the column totals are summed over the ranks (peer.cuh)
  PeerCtx peer;                                       //            before the finalisation, inside this kernel
};

__device__ unsigned int g_tickets[1024];   // zero at load; the last CTA of a launch resets its slot.  Slots rotate per launch
                                           // (host counter), so concurrent streams collide only 1024 launches apart.

template <int MODE>
__global__ void __launch_bounds__(COL_THREADS) k_col_reduce(const float* __restrict__ x, const float* __restrict__ dy, int64_t m, int d,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                                            float p_drop, uint64_t seed, uint64_t sid, double* __restrict__ partial,
                                                            int ticket_slot, ColFin fin) {
  extern __shared__ double sm[];  // [rows_per_pass][2][d]
  __shared__ int is_last;
  if (MODE == 1 && p_drop > 0.f) resolve_seed(seed, sid);
  const int tpr = d >> 3;               // threads per row (8 columns each = one Philox call)
  const int rpp = COL_THREADS / tpr;    // rows per pass of the block
  const int cg = threadIdx.x % tpr, rs = threadIdx.x / tpr;
  const int c = cg * 8;
  double a0[8], a1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a0[e] = a1[e] = 0.0;
  float mu[8], rsd[8], ga[8], be[8];
  if (MODE == 1) {
    ld8(mean + c, mu); ld8(rstd + c, rsd); ld8(gamma + c, ga); ld8(beta + c, be);
  }
#pragma unroll(MODE == 0 ? 4 : 2)
  for (int64_t r = (int64_t)blockIdx.x * rpp + rs; r < m; r += (int64_t)gridDim.x * rpp) {
    const size_t off = (size_t)r * d + c;
    float xs[8];
    ld8(x + off, xs);
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a0[e] += (double)xs[e];
        a1[e] += (double)xs[e] * (double)xs[e];
      }
    } else {
      float g[8];
      ld8(dy + off, g);
      if (p_drop > 0.f) {
        float mk[8];
        dropout_scale8(seed, sid, off >> 3, p_drop, mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] *= mk[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xs[e] - mu[e]) * rsd[e];
        g[e] *= act_grad(fmaf(xh, ga[e], be[e]), act);
        a0[e] += (double)g[e];
        a1[e] += (double)g[e] * (double)xh;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sm[(size_t)(rs * 2 + 0) * d + c + e] = a0[e];
    sm[(size_t)(rs * 2 + 1) * d + c + e] = a1[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += COL_THREADS) {
    const int which = i / d, col = i % d;
    double s = 0;
    for (int rr = 0; rr < rpp; ++rr) s += sm[(size_t)(rr * 2 + which) * d + col];
    partial[((size_t)blockIdx.x * 2 + which) * d + col] = s;
  }
  // ---- ticket: the last CTA finalises ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&g_tickets[ticket_slot], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int n_part = gridDim.x, outs = 2 * d;
  const int n_slices = COL_THREADS / outs > 0 ? COL_THREADS / outs : 1;   // d = 256: 1, 128: 2, 64: 4, 32: 8
  double* comb = sm;                                                        // [n_slices][outs]
  for (int o0 = 0; o0 < outs; o0 += COL_THREADS) {
    const int idx = o0 + threadIdx.x;
    const int o = idx % outs, sl = (idx / outs) % n_slices;
    if (idx < outs * n_slices) {
      double acc = 0;
#pragma unroll 16
      for (int p = sl; p < n_part; p += n_slices) acc += __ldcg(partial + (size_t)p * outs + o);   // 16 L2 loads in flight
      comb[sl * outs + o] = acc;
    }
  }
  __syncthreads();
  {
    const int cc = threadIdx.x;                       // d <= 256 < COL_THREADS: one column per thread
    const bool col = cc < d;
    double s0 = 0, s1 = 0;
    if (col) {
      for (int sl = 0; sl < n_slices; ++sl) {
        s0 += comb[sl * outs + cc];
        s1 += comb[sl * outs + d + cc];
      }
    }
    if (fin.peer_on) {
      // patient-partitioned BatchNorm: this rank's totals -> own symmetric slice, rendezvous with the same CTA of every
      // other rank, global totals = sum in rank order of everybody's slice (bit-identical on all ranks)
      const uint32_t seq = peer_next_seq(fin.peer, PEER_SLOT_BN);
      const size_t off = peer_slice_off(PEER_SLOT_BN, seq & 1u);
      double* mine = reinterpret_cast<double*>(fin.peer.base[fin.peer.rank] + off);
      if (col) {
        mine[cc] = s0;
        mine[d + cc] = s1;
      }
      peer_signal_wait(fin.peer, PEER_SLOT_BN, seq);
      if (col) {
        s0 = s1 = 0;
        for (int r = 0; r < fin.peer.world; ++r) {
          const double* theirs = reinterpret_cast<const double*>(fin.peer.base[r] + off);
          s0 += ld_volatile_f64(theirs + cc);
          s1 += ld_volatile_f64(theirs + d + cc);
        }
      }
      if (threadIdx.x == 0) peer_commit_seq(fin.peer, PEER_SLOT_BN, seq);
    }
    if (!col) {
      // nothing to finalise
    } else if (fin.kind == 0) {
      const double mm = (double)fin.m;
      const double mean_ = s0 / mm;
      double var = s1 / mm - mean_ * mean_;
      if (var < 0) var = 0;
      fin.mean[cc] = (float)mean_;
      fin.rstd[cc] = (float)(1.0 / sqrt(var + (double)fin.eps));
      if (fin.running_mean) fin.running_mean[cc] = (1.f - fin.momentum) * fin.running_mean[cc] + fin.momentum * (float)mean_;
      if (fin.running_var) {
        const double unb = fin.m > 1 ? var * (mm / (double)(fin.m - 1)) : var;
        fin.running_var[cc] = (1.f - fin.momentum) * fin.running_var[cc] + fin.momentum * (float)unb;
      }
    } else if (fin.kind == 1 || fin.kind == 3) {
      fin.sums[cc] = s0;
      fin.sums[d + cc] = s1;
      if (fin.kind == 1) {
        if (fin.dbeta) fin.dbeta[cc] = (float)s0;
        if (fin.dgamma) fin.dgamma[cc] = (float)s1;
      }
    } else {
      fin.out[cc] = (float)s0;
    }
  }
  if (threadIdx.x == 0) g_tickets[ticket_slot] = 0u;
}

__global__ void k_bn_finalize_sums(const double* __restrict__ sums, double m, int d, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double mu = sums[c] / m;
  double var = sums[d + c] / m - mu * mu;
  if (var < 0) var = 0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) {
    double unb = m > 1 ? var * (m / (m - 1)) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void k_sums_to_float(const double* __restrict__ sums, int d, float* __restrict__ dbeta, float* __restrict__ dgamma) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  if (dbeta) dbeta[c] = (float)sums[c];
  if (dgamma) dgamma[c] = (float)sums[d + c];
}

__global__ void k_bn_eval_stats(const float* __restrict__ rm, const float* __restrict__ rv, int d, float eps, float* __restrict__ mean,
                                float* __restrict__ rstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  mean[c] = rm[c];
  rstd[c] = 1.0f / sqrtf(rv[c] + eps);
}

// Element-wise passes: a thread owns 8 consecutive elements (one Philox call, two 16-byte accesses).  blockDim x 8
// elements is a multiple of every supported d, so a thread's columns never change over the grid-stride loop: the
// per-column constants live in 32 registers (in shared memory they cost 8 LDS.128 per octet -- 4x the wavefronts of the
// global accesses themselves: ncu showed the L1 / shared-memory pipe 91 % busy and the kernel at 4.9 TB/s); two octets
// per trip keep four 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256, 3) k_bn_apply(const float* __restrict__ x, int64_t n8, int d, const float* __restrict__ mean,
                                                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  int act, float p_drop, uint64_t seed, uint64_t sid, float* __restrict__ y) {
  if (p_drop > 0.f) resolve_seed(seed, sid);
  const int c = (int)((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8) % d);
  float mu[8], rs[8], ga[8], be[8];
  ld8(mean + c, mu); ld8(rstd + c, rs); ld8(gamma + c, ga); ld8(beta + c, be);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += 2 * stride) {
    float v[2][8];
    const bool second = i + stride < n8;
    ld8(x + i * 8, v[0]);
    if (second) ld8(x + (i + stride) * 8, v[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !second) break;
      const int64_t iu = i + u * stride;
      float mk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (p_drop > 0.f) dropout_scale8(seed, sid, (uint64_t)iu, p_drop, mk);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[u][e] = act_fwd(fmaf((v[u][e] - mu[e]) * rs[e], ga[e], be[e]), act) * mk[e];
      st8(y + iu * 8, v[u]);
    }
  }
}

// dx of BatchNorm(+act+dropout).  One CTA of 512 threads per SM, grid-stride over 8-element octets (a thread's columns
// never change).  Optionally also the column sums of dx (= the bias gradient of the Linear that produced x, which would
// otherwise cost another full pass over dx): per-thread fp64 sums -> per-CTA record -> the last CTA (ticket) adds the
// records in CTA order: deterministic.
constexpr int BWA_THREADS = 512;
__global__ void __launch_bounds__(BWA_THREADS, 1) k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy, int64_t n8, int64_t m, int d,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta, int act, float p_drop,
                                                      uint64_t seed, uint64_t sid, int batch_stats, const double* __restrict__ sums,
                                                      float* __restrict__ dx, double* __restrict__ cs_partial, float* __restrict__ colsum,
                                                      int ticket_slot) {
  extern __shared__ double dyn_cs[];             // [BWA_THREADS][8]   (column-sum variant only)
  __shared__ int is_last;
  if (p_drop > 0.f) resolve_seed(seed, sid);
  const float inv_m = 1.0f / (float)m;
  const int c = (int)((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8) % d);
  float mu[8], rs[8], ga[8], be[8], sg[8], sgx[8];
  ld8(mean + c, mu); ld8(rstd + c, rs); ld8(gamma + c, ga); ld8(beta + c, be);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[e] = batch_stats ? (float)sums[c + e] * inv_m : 0.f;
    sgx[e] = batch_stats ? (float)sums[d + c + e] * inv_m : 0.f;
  }
  double cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += 2 * stride) {
    float xs[2][8], g[2][8];
    const bool second = i0 + stride < n8;          // two octets per trip: twice the bytes in flight per thread
    ld8(x + i0 * 8, xs[0]);
    ld8(dy + i0 * 8, g[0]);
    if (second) {
      ld8(x + (i0 + stride) * 8, xs[1]);
      ld8(dy + (i0 + stride) * 8, g[1]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !second) break;
      const int64_t i = i0 + u * stride;
      if (p_drop > 0.f) {
        float mk[8];
        dropout_scale8(seed, sid, (uint64_t)i, p_drop, mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[u][e] *= mk[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xs[u][e] - mu[e]) * rs[e];
        float gg = g[u][e] * act_grad(fmaf(xh, ga[e], be[e]), act);
        g[u][e] = batch_stats ? ga[e] * rs[e] * (gg - sg[e] - xh * sgx[e]) : ga[e] * rs[e] * gg;
        cs[e] += (double)g[u][e];
      }
      st8(dx + i * 8, g[u]);
    }
  }
  if (colsum == nullptr) return;
  // ---- column sums of dx: threads t and t' share columns iff t = t' (mod d/8) ----
  const int tpr = d >> 3;                           // 4 .. 32 threads cover one row
  // serial, fixed-order accumulation over the BWA_THREADS / tpr threads of each column set, done by the first tpr threads
#pragma unroll
  for (int e = 0; e < 8; ++e) dyn_cs[threadIdx.x * 8 + e] = cs[e];
  __syncthreads();
  if ((int)threadIdx.x < tpr) {
    double tot[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) tot[e] = 0.0;
    for (int t = threadIdx.x; t < BWA_THREADS; t += tpr)
#pragma unroll
      for (int e = 0; e < 8; ++e) tot[e] += dyn_cs[t * 8 + e];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs_partial[(size_t)blockIdx.x * d + threadIdx.x * 8 + e] = tot[e];
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&g_tickets[ticket_slot], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int n_part = gridDim.x;
  const int n_slices = BWA_THREADS / d;             // d = 256: 2 ... d = 32: 16
  {
    const int col = threadIdx.x % d, sl = threadIdx.x / d;
    double acc = 0;
#pragma unroll 16
    for (int p = sl; p < n_part; p += n_slices) acc += __ldcg(cs_partial + (size_t)p * d + col);
    dyn_cs[sl * d + col] = acc;
  }
  __syncthreads();
  if ((int)threadIdx.x < d) {
    double t = 0;
    for (int sl = 0; sl < n_slices; ++sl) t += dyn_cs[sl * d + threadIdx.x];
    colsum[threadIdx.x] = (float)t;
  }
  if (threadIdx.x == 0) g_tickets[ticket_slot] = 0u;
}

// ---- ReLU / Dropout --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_relu_dropout_fwd(const float* __restrict__ x, int64_t n, int relu, float p, uint64_t seed, uint64_t sid,
                                                          float* __restrict__ y) {
  if (p > 0.f) resolve_seed(seed, sid);
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p > 0.f) mk = dropout_scale4(seed, sid, (uint64_t)i, p);
    float mks[4] = {mk.x, mk.y, mk.z, mk.w};
    if (i * 4 + 3 < n) {
      float4 v = reinterpret_cast<const float4*>(x)[i];
      float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) vs[q] = (relu ? fmaxf(vs[q], 0.f) : vs[q]) * mks[q];
      reinterpret_cast<float4*>(y)[i] = make_float4(vs[0], vs[1], vs[2], vs[3]);
    } else {
      for (int q = 0; q < 4 && i * 4 + q < n; ++q) {
        float v = x[i * 4 + q];
        y[i * 4 + q] = (relu ? fmaxf(v, 0.f) : v) * mks[q];
      }
    }
  }
}

// dx = dy * mask * [y > 0]  (y is the forward OUTPUT: y > 0 <=> pre-activation > 0 and kept)
__global__ void __launch_bounds__(256) k_relu_dropout_bwd(const float* __restrict__ y, const float* __restrict__ dy, int64_t n, int relu, float p,
                                                          uint64_t seed, uint64_t sid, float* __restrict__ dx) {
  if (p > 0.f) resolve_seed(seed, sid);
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p > 0.f) mk = dropout_scale4(seed, sid, (uint64_t)i, p);
    float mks[4] = {mk.x, mk.y, mk.z, mk.w};
    for (int q = 0; q < 4 && i * 4 + q < n; ++q) {
      int64_t e = i * 4 + q;
      float g = dy[e] * mks[q];
      if (relu && !(y[e] > 0.f)) g = 0.f;
      dx[e] = g;
    }
  }
}

__global__ void __launch_bounds__(256) k_dropout_mask(int64_t n, float p, uint64_t seed, uint64_t sid, float* __restrict__ mask) {
  if (p > 0.f) resolve_seed(seed, sid);
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p > 0.f) mk = dropout_scale4(seed, sid, (uint64_t)i, p);
    float mks[4] = {mk.x, mk.y, mk.z, mk.w};
    for (int q = 0; q < 4 && i * 4 + q < n; ++q) mask[i * 4 + q] = mks[q];
  }
}

// ---- row L2 normalisation -----------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_l2norm_fwd(const float* __restrict__ x, int64_t m, float eps, float* __restrict__ y,
                                                    float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= m) return;
  RowVec<D> v;
  v.load(x + (size_t)r * D, lane);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RowVec<D>::N; ++i) ss = fmaf(v.v[i], v.v[i], ss);
  ss = warp_sum(ss);
  float inv = 1.0f / fmaxf(sqrtf(ss), eps);
#pragma unroll
  for (int i = 0; i < RowVec<D>::N; ++i) v.v[i] *= inv;
  v.store(y + (size_t)r * D, lane);
  if (lane == 0) inv_norm[r] = inv;
}

// y = x * inv  =>  dx = inv * (dy - y * <y, dy>)      (rows clamped by eps have <.,.> ~ 0 and reduce to inv * dy)
template <int D>
__global__ void __launch_bounds__(256) k_l2norm_bwd(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ inv_norm,
                                                    int64_t m, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= m) return;
  RowVec<D> yv, gv;
  yv.load(y + (size_t)r * D, lane);
  gv.load(dy + (size_t)r * D, lane);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < RowVec<D>::N; ++i) dot = fmaf(yv.v[i], gv.v[i], dot);
  dot = warp_sum(dot);
  float inv = __ldg(inv_norm + r);
#pragma unroll
  for (int i = 0; i < RowVec<D>::N; ++i) gv.v[i] = inv * (gv.v[i] - yv.v[i] * dot);
  gv.store(dx + (size_t)r * D, lane);
}

// The same pass with the column sums of dx riding along (dx is the `dy` of the Linear in front of the normalisation: its bias
// gradient would otherwise cost a pass of its own).  Persistent grid; per-thread fp64 sums in row order, the CTA's 8 warps
// combined in warp order, one record per CTA; k_rec_reduce adds the records in CTA order: deterministic.
template <int D>
__global__ void __launch_bounds__(256) k_l2norm_bwd_cs(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ inv_norm,
                                                       int64_t m, float* __restrict__ dx, double* __restrict__ rec) {
  __shared__ double sh[8][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int N = RowVec<D>::N;
  double cs[N];
#pragma unroll
  for (int i = 0; i < N; ++i) cs[i] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * 8;
  for (int64_t r0 = (int64_t)blockIdx.x * 8 + warp; r0 < m; r0 += 2 * stride) {     // two rows per trip: four row loads in flight
    const int64_t r1 = r0 + stride;
    const bool two = r1 < m;
    RowVec<D> yv[2], gv[2];
    yv[0].load(y + (size_t)r0 * D, lane);
    gv[0].load(dy + (size_t)r0 * D, lane);
    if (two) {
      yv[1].load(y + (size_t)r1 * D, lane);
      gv[1].load(dy + (size_t)r1 * D, lane);
    }
    float inv[2];
    inv[0] = __ldg(inv_norm + r0);
    inv[1] = two ? __ldg(inv_norm + r1) : 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) dot = fmaf(yv[u].v[i], gv[u].v[i], dot);
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        gv[u].v[i] = inv[u] * (gv[u].v[i] - yv[u].v[i] * dot);
        cs[i] += (double)gv[u].v[i];
      }
      gv[u].store(dx + (size_t)(u ? r1 : r0) * D, lane);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {                        // column of element i of a lane's row slice (RowVec layout)
    const int col = D >= 128 ? (i >> 2) * 128 + lane * 4 + (i & 3) : (D == 64 ? lane * 2 + i : lane);
    sh[warp][col] = cs[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sh[w][c];
    rec[(size_t)blockIdx.x * D + c] = a;
  }
}
// out[c] = (float) sum over CTAs of rec[cta][c]: 8 interleaved slices of the records, combined in fixed order
__global__ void __launch_bounds__(256) k_rec_reduce(const double* __restrict__ rec, int n_cta, int d, float* __restrict__ out) {
  __shared__ double sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (c < d)
    for (int k = slice; k < n_cta; k += 8) a += rec[(size_t)k * d + c];
  sh[slice][lane] = a;
  __syncthreads();
  if (slice == 0 && c < d) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += sh[k][lane];
    out[c] = (float)a;
  }
}

// ---- loss -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float loss_term(float diff, int kind) {
  float a = fabsf(diff);
  if (kind == 0) return a;
  if (kind == 1) return diff * diff;
  return a < 1.f ? 0.5f * diff * diff : a - 0.5f;
}
__device__ __forceinline__ float loss_dterm(float diff, int kind) {
  float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
  if (kind == 0) return sgn;
  if (kind == 1) return 2.f * diff;
  return fabsf(diff) < 1.f ? diff : sgn;
}

__global__ void __launch_bounds__(256) k_loss_partial(const float* __restrict__ pred, const float* __restrict__ target,
                                                      const int64_t* __restrict__ lab, const float* __restrict__ w, const uint8_t* __restrict__ sup,
                                                      int64_t m, int kind, double* __restrict__ part_sum, unsigned long long* __restrict__ part_cnt) {
  __shared__ double ss[8];
  __shared__ unsigned long long sc[8];
  double s = 0;
  unsigned long long cnt = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += 4 * stride) {
    float term[4];
    bool on[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {            // 4 independent element chains in flight per thread
      const int64_t i = i0 + u * stride;
      on[u] = i < m && (!sup || sup[i]);
      term[u] = 0.f;
      if (on[u]) {
        const float wt = w ? __ldg(w + __ldg(lab + i)) : 1.f;
        term[u] = wt * loss_term(pred[i] - target[i], kind);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {            // element order i0, i0 + stride, ... : fixed
      if (on[u]) {
        s += (double)term[u];
        ++cnt;
      }
    }
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    ss[threadIdx.x >> 5] = s;
    sc[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    unsigned long long c = 0;
    for (int i = 0; i < 8; ++i) {
      t += ss[i];
      c += sc[i];
    }
    part_sum[blockIdx.x] = t;
    part_cnt[blockIdx.x] = c;
  }
}

// totals of the per-CTA partials by one warp (lanes stride over the partials, fixed-order butterfly): every caller gets the
// same bits
__device__ __forceinline__ void loss_totals(const double* __restrict__ part_sum, const unsigned long long* __restrict__ part_cnt, int n_part,
                                            int lane, double& t, unsigned long long& c) {
  t = 0;
  c = 0;
  for (int i = lane; i < n_part; i += 32) {
    t += part_sum[i];
    c += part_cnt[i];
  }
  t = warp_sum(t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
}

__global__ void k_loss_final(const double* __restrict__ part_sum, const unsigned long long* __restrict__ part_cnt, int n_part,
                             float* __restrict__ loss) {
  double t;
  unsigned long long c;
  loss_totals(part_sum, part_cnt, n_part, threadIdx.x & 31, t, c);
  if (threadIdx.x == 0) *loss = (float)(t * (1.0 / (double)c));   // c == 0 -> inf -> loss NaN, like torch's mean over an empty selection
}

// gradient + (block 0) the loss value: every CTA re-derives the supervised count from the partial counts itself
__global__ void __launch_bounds__(256) k_loss_grad(const float* __restrict__ pred, const float* __restrict__ target, const int64_t* __restrict__ lab,
                                                   const float* __restrict__ w, const uint8_t* __restrict__ sup, int64_t m, int kind,
                                                   const double* __restrict__ part_sum, const unsigned long long* __restrict__ part_cnt,
                                                   int n_part, float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ float inv_s;
  if (threadIdx.x < 32) {
    double t;
    unsigned long long c;
    loss_totals(part_sum, part_cnt, n_part, threadIdx.x, t, c);
    if (threadIdx.x == 0) {
      const double inv = 1.0 / (double)c;
      inv_s = (float)inv;
      if (blockIdx.x == 0) *loss = (float)(t * inv);
    }
  }
  __syncthreads();
  const float inv = inv_s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (!sup || sup[i]) {
      float wt = w ? __ldg(w + __ldg(lab + i)) : 1.f;
      g = wt * loss_dterm(pred[i] - target[i], kind) * inv;
    }
    grad[i] = g;
  }
}

inline int ew_grid(int64_t n_items) {
  int64_t g = ceil_div(n_items, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline int col_parts(int64_t m, int d) {
  int rpp = COL_THREADS / (d / 8);
  int64_t g = ceil_div(m, (int64_t)rpp * 4);          // >= 4 passes per CTA before another CTA is worth its partial
  int cap = sm_count() < MAX_PARTIALS ? sm_count() : MAX_PARTIALS;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline size_t col_smem(int d) { return (size_t)(COL_THREADS / (d / 8)) * 2 * d * sizeof(double); }
inline int next_ticket_slot() {
  static std::atomic<unsigned> n{0};
  return (int)(n.fetch_add(1u) & 1023u);
}
template <int MODE>
int launch_col_reduce(const float* x, const float* dy, int64_t m, int d, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, int act, float p_drop, uint64_t seed, uint64_t sid, double* partial, const ColFin& fin,
                      cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_col_reduce<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr_set = true;
  }
  k_col_reduce<MODE><<<col_parts(m, d), COL_THREADS, col_smem(d), st>>>(x, dy, m, d, mean, rstd, gamma, beta, act, p_drop, seed, sid, partial,
                                                                        next_ticket_slot(), fin);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}
inline bool d_ok(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }
}  // namespace

extern "C" size_t b2g_bn_ws_bytes(int d) { return align_up((size_t)MAX_PARTIALS * 2 * d * 8, 256) + align_up((size_t)2 * d * 8, 256); }

// dx pass of the BatchNorm backward (+ optional column sums of dx through `cs_partial`, >= sm_count * d doubles)
static int launch_bwd_apply(const float* x, const float* dy, int64_t m, int64_t m_stat, int d, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, int act, float p_drop, uint64_t seed, uint64_t sid, int batch_stats,
                            const double* sums, float* dx, double* cs_partial, float* colsum, cudaStream_t st) {
  const int64_t n8 = m * d / 8;
  int64_t g = ceil_div(n8, BWA_THREADS);
  if (g > sm_count()) g = sm_count();
  if (g < 1) g = 1;
  k_bn_bwd_apply<<<(unsigned)g, BWA_THREADS, colsum ? BWA_THREADS * 8 * sizeof(double) : 0, st>>>(
      x, dy, n8, m_stat, d, mean, rstd, gamma, beta, act, p_drop, seed, sid, batch_stats, sums, dx, cs_partial, colsum,
      colsum ? next_ticket_slot() : 0);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

#define COL_WS_CHECK(name)                                   \
  if (!ws || ws_bytes < b2g_bn_ws_bytes(d)) {                \
    set_error(name ": workspace too small");                 \
    return B2G_EWS;                                          \
  }

extern "C" int b2g_bn_stats(const float* x, int64_t m, int d, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                            float* running_var, void* ws, size_t ws_bytes, void* stream_) {
  B2G_CHECK_ARG(x && mean && rstd && m > 0 && d_ok(d) && aligned16(x), "bn_stats: bad args (m=%lld d=%d)", (long long)m, d);
  COL_WS_CHECK("bn_stats");
  ColFin fin{};
  fin.kind = 0; fin.m = m; fin.eps = eps; fin.momentum = momentum;
  fin.mean = mean; fin.rstd = rstd; fin.running_mean = running_mean; fin.running_var = running_var;
  return launch_col_reduce<0>(x, nullptr, m, d, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, (double*)ws, fin, (cudaStream_t)stream_);
}

/* out[c] = sum_r x[r, c]  (bias gradient of a linear layer), fp64 accumulation, fixed order */
extern "C" int b2g_col_sums(const float* x, int64_t m, int d, float* out, void* ws, size_t ws_bytes, void* stream_) {
  B2G_CHECK_ARG(x && out && m > 0 && d_ok(d) && aligned16(x), "col_sums: bad args (m=%lld d=%d)", (long long)m, d);
  COL_WS_CHECK("col_sums");
  ColFin fin{};
  fin.kind = 2; fin.out = out;
  return launch_col_reduce<0>(x, nullptr, m, d, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, (double*)ws, fin, (cudaStream_t)stream_);
}

/* Patient-partitioned BatchNorm (multi-GPU): each rank reduces its rows to fp64 column totals sums[2*d] = {sum x, sum x^2}
 * (b2g_bn_local_sums), the ranks all-reduce them, and b2g_bn_finalize_sums turns the global totals + global row count
 * into mean / rstd (+ running-stat update).  Backward: b2g_bn_bwd_local_sums gives {sum g, sum g*xhat}; after the
 * all-reduce b2g_bn_bwd_from_sums writes dx (m_total = global row count) and dgamma / dbeta. */
extern "C" int b2g_bn_local_sums(const float* x, int64_t m, int d, double* sums, void* ws, size_t ws_bytes, void* stream_) {
  B2G_CHECK_ARG(x && sums && m > 0 && d_ok(d) && aligned16(x), "bn_local_sums: bad args");
  COL_WS_CHECK("bn_local_sums");
  ColFin fin{};
  fin.kind = 3; fin.sums = sums;
  return launch_col_reduce<0>(x, nullptr, m, d, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, (double*)ws, fin, (cudaStream_t)stream_);
}

extern "C" int b2g_bn_finalize_sums(const double* sums, int64_t m_total, int d, float eps, float momentum, float* mean, float* rstd,
                                    float* running_mean, float* running_var, void* stream_) {
  B2G_CHECK_ARG(sums && mean && rstd && m_total > 0 && d > 0, "bn_finalize_sums: bad args");
  k_bn_finalize_sums<<<(unsigned)ceil_div(d, 128), 128, 0, (cudaStream_t)stream_>>>(sums, (double)m_total, d, eps, momentum, mean, rstd,
                                                                                 running_mean, running_var);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_bn_bwd_local_sums(const float* x, const float* dy, int64_t m, int d, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, int relu, float p_drop, uint64_t seed, uint64_t stream_id,
                                     double* sums, void* ws, size_t ws_bytes, void* stream_) {
  B2G_CHECK_ARG(m > 0 && d_ok(d) && x && dy && mean && rstd && gamma && beta && sums, "bn_bwd_local_sums: bad args");
  B2G_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "bn_bwd_local_sums: unaligned pointer");
  COL_WS_CHECK("bn_bwd_local_sums");
  ColFin fin{};
  fin.kind = 3; fin.sums = sums;
  return launch_col_reduce<1>(x, dy, m, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, (double*)ws, fin, (cudaStream_t)stream_);
}

extern "C" int b2g_bn_bwd_from_sums(const float* x, const float* dy, int64_t m, int64_t m_total, int d, const float* mean,
                                    const float* rstd, const float* gamma, const float* beta, int relu, float p_drop, uint64_t seed,
                                    uint64_t stream_id, const double* sums, float* dx, float* dgamma, float* dbeta, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m > 0 && m_total >= m && d_ok(d) && x && dy && mean && rstd && gamma && beta && sums && dx, "bn_bwd_from_sums: bad args");
  B2G_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "bn_bwd_from_sums: unaligned pointer");
  int rc = launch_bwd_apply(x, dy, m, m_total, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, 1, sums, dx, nullptr, nullptr, st);
  if (rc != B2G_OK) return rc;
  if (dgamma || dbeta) {
    k_sums_to_float<<<(unsigned)ceil_div(d, 128), 128, 0, st>>>(sums, d, dbeta, dgamma);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}

/* Fused multi-GPU variants (peer.cuh): the per-rank column totals are exchanged over NVLink peer memory by the last CTA
 * of the reduction kernel itself -- statistics, exchange and finalisation are ONE launch (instead of local sums ->
 * NCCL all-reduce -> finalize).  m = local rows, m_total = rows over all ranks.  dgamma / dbeta are the GLOBAL totals. */
extern "C" const void* b2g_comm_ctx(const struct b2g_comm* c);

extern "C" int b2g_bn_stats_sync(struct b2g_comm* comm, const float* x, int64_t m, int64_t m_total, int d, float eps, float momentum,
                                 float* mean, float* rstd, float* running_mean, float* running_var, void* ws, size_t ws_bytes,
                                 void* stream_) {
  B2G_CHECK_ARG(comm && x && mean && rstd && m > 0 && m_total >= m && d_ok(d) && aligned16(x), "bn_stats_sync: bad args");
  COL_WS_CHECK("bn_stats_sync");
  ColFin fin{};
  fin.kind = 0; fin.m = m_total; fin.eps = eps; fin.momentum = momentum;
  fin.mean = mean; fin.rstd = rstd; fin.running_mean = running_mean; fin.running_var = running_var;
  fin.peer_on = 1;
  fin.peer = *reinterpret_cast<const PeerCtx*>(b2g_comm_ctx(comm));
  return launch_col_reduce<0>(x, nullptr, m, d, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0, 0, (double*)ws, fin, (cudaStream_t)stream_);
}

extern "C" int b2g_bn_bwd_sync(struct b2g_comm* comm, const float* x, const float* dy, int64_t m, int64_t m_total, int d, const float* mean,
                               const float* rstd, const float* gamma, const float* beta, int relu, float p_drop, uint64_t seed,
                               uint64_t stream_id, float* dx, float* dgamma, float* dbeta, float* dx_colsum, void* ws, size_t ws_bytes,
                               void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(comm && m > 0 && m_total >= m && d_ok(d) && x && dy && mean && rstd && gamma && beta && dx, "bn_bwd_sync: bad args");
  B2G_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "bn_bwd_sync: unaligned pointer");
  COL_WS_CHECK("bn_bwd_sync");
  double* partial = (double*)ws;
  double* sums = (double*)((char*)ws + align_up((size_t)MAX_PARTIALS * 2 * d * 8, 256));
  ColFin fin{};
  fin.kind = 1; fin.sums = sums; fin.dgamma = dgamma; fin.dbeta = dbeta;
  fin.peer_on = 1;
  fin.peer = *reinterpret_cast<const PeerCtx*>(b2g_comm_ctx(comm));
  int rc = launch_col_reduce<1>(x, dy, m, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, partial, fin, st);
  if (rc != B2G_OK) return rc;
  return launch_bwd_apply(x, dy, m, m_total, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, 1, sums, dx, partial, dx_colsum, st);
}

extern "C" int b2g_bn_eval_stats(const float* running_mean, const float* running_var, int d, float eps, float* mean, float* rstd,
                                 void* stream_) {
  B2G_CHECK_ARG(running_mean && running_var && mean && rstd && d > 0, "bn_eval_stats: bad args");
  k_bn_eval_stats<<<(unsigned)ceil_div(d, 128), 128, 0, (cudaStream_t)stream_>>>(running_mean, running_var, d, eps, mean, rstd);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_bn_apply(const float* x, int64_t m, int d, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            int relu, float p_drop, uint64_t seed, uint64_t stream_id, float* y, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && d_ok(d) && (m == 0 || (x && y && mean && rstd && gamma && beta)), "bn_apply: bad args");
  B2G_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "bn_apply: dropout p must be in [0,1)");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "bn_apply: unaligned pointer");
  int64_t n8 = m * d / 8;
  k_bn_apply<<<ew_grid(n8), 256, 0, (cudaStream_t)stream_>>>(x, n8, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, y);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_bn_bwd(const float* x, const float* dy, int64_t m, int d, const float* mean, const float* rstd, const float* gamma,
                          const float* beta, int relu, float p_drop, uint64_t seed, uint64_t stream_id, int batch_stats, float* dx,
                          float* dgamma, float* dbeta, float* dx_colsum, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m > 0 && d_ok(d) && x && dy && mean && rstd && gamma && beta && dx, "bn_bwd: bad args");
  B2G_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(mean) && aligned16(rstd) && aligned16(gamma) && aligned16(beta),
                "bn_bwd: unaligned pointer");
  COL_WS_CHECK("bn_bwd");
  double* partial = (double*)ws;
  double* sums = (double*)((char*)ws + align_up((size_t)MAX_PARTIALS * 2 * d * 8, 256));
  ColFin fin{};
  fin.kind = 1; fin.sums = sums; fin.dgamma = dgamma; fin.dbeta = dbeta;
  int rc = launch_col_reduce<1>(x, dy, m, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, partial, fin, st);
  if (rc != B2G_OK) return rc;
  return launch_bwd_apply(x, dy, m, m, d, mean, rstd, gamma, beta, relu, p_drop, seed, stream_id, batch_stats, sums, dx, partial, dx_colsum, st);
}

extern "C" int b2g_relu_dropout_fwd(const float* x, int64_t n, int relu, float p_drop, uint64_t seed, uint64_t stream_id, float* y,
                                    void* stream_) {
  B2G_CHECK_ARG(n >= 0 && (n == 0 || (x && y)) && p_drop >= 0.f && p_drop < 1.f, "relu_dropout_fwd: bad args");
  if (n == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(x) && aligned16(y), "relu_dropout_fwd: unaligned pointer");
  k_relu_dropout_fwd<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream_>>>(x, n, relu, p_drop, seed, stream_id, y);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_relu_dropout_bwd(const float* y, const float* dy, int64_t n, int relu, float p_drop, uint64_t seed, uint64_t stream_id,
                                    float* dx, void* stream_) {
  B2G_CHECK_ARG(n >= 0 && (n == 0 || (y && dy && dx)) && p_drop >= 0.f && p_drop < 1.f, "relu_dropout_bwd: bad args");
  if (n == 0) return B2G_OK;
  k_relu_dropout_bwd<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream_>>>(y, dy, n, relu, p_drop, seed, stream_id, dx);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_dropout_mask(int64_t n, float p_drop, uint64_t seed, uint64_t stream_id, float* mask, void* stream_) {
  B2G_CHECK_ARG(n >= 0 && (n == 0 || mask) && p_drop >= 0.f && p_drop < 1.f, "dropout_mask: bad args");
  if (n == 0) return B2G_OK;
  k_dropout_mask<<<ew_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream_>>>(n, p_drop, seed, stream_id, mask);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_l2norm_fwd(const float* x, int64_t m, int d, float eps, float* y, float* inv_norm, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (x && y && inv_norm)), "l2norm_fwd: bad args");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(x) && aligned16(y), "l2norm_fwd: unaligned pointer");
  unsigned grid = (unsigned)ceil_div(m, 8);
  DISPATCH_D(d, (k_l2norm_fwd<D><<<grid, 256, 0, (cudaStream_t)stream_>>>(x, m, eps, y, inv_norm)));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_l2norm_bwd(const float* y, const float* dy, const float* inv_norm, int64_t m, int d, float* dx, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (y && dy && inv_norm && dx)), "l2norm_bwd: bad args");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(y) && aligned16(dy) && aligned16(dx), "l2norm_bwd: unaligned pointer");
  unsigned grid = (unsigned)ceil_div(m, 8);
  DISPATCH_D(d, (k_l2norm_bwd<D><<<grid, 256, 0, (cudaStream_t)stream_>>>(y, dy, inv_norm, m, dx)));
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_l2norm_bwd_cs_ws_bytes(int d) { return (size_t)sm_count() * 8 * d * sizeof(double) + 256; }
extern "C" int b2g_l2norm_bwd_cs(const float* y, const float* dy, const float* inv_norm, int64_t m, int d, float* dx, float* dx_colsum,
                                 void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m > 0 && d_ok(d) && y && dy && inv_norm && dx && dx_colsum, "l2norm_bwd_cs: bad args");
  B2G_CHECK_ARG(aligned16(y) && aligned16(dy) && aligned16(dx) && aligned16(ws), "l2norm_bwd_cs: unaligned pointer");
  if (!ws || ws_bytes < b2g_l2norm_bwd_cs_ws_bytes(d)) {
    set_error("l2norm_bwd_cs: workspace too small");
    return B2G_EWS;
  }
  int64_t blocks = ceil_div(m, 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  double* rec = (double*)ws;
  DISPATCH_D(d, (k_l2norm_bwd_cs<D><<<grid, 256, 0, st>>>(y, dy, inv_norm, m, dx, rec)));
  B2G_LAUNCH_CHECK();
  k_rec_reduce<<<(unsigned)ceil_div(d, 32), 256, 0, st>>>(rec, grid, d, dx_colsum);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

constexpr int LOSS_MAX_PARTS = 2048;     // the reduction is latency-bound (dependent lab -> weight loads): many CTAs in flight
extern "C" size_t b2g_loss_ws_bytes(int64_t m) {
  (void)m;
  return align_up((size_t)LOSS_MAX_PARTS * 8, 256) * 2 + 256;
}

extern "C" int b2g_weighted_loss(const float* pred, const float* target, const int64_t* lab, const float* w, const uint8_t* sup, int64_t m,
                                 int kind, float* loss, float* grad, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m >= 0 && loss && (m == 0 || (pred && target)) && kind >= 0 && kind <= 2 && (!w || lab), "weighted_loss: bad args");
  if (!ws || ws_bytes < b2g_loss_ws_bytes(m)) {
    set_error("weighted_loss: workspace too small");
    return B2G_EWS;
  }
  double* part_sum = (double*)ws;
  unsigned long long* part_cnt = (unsigned long long*)((char*)ws + align_up((size_t)LOSS_MAX_PARTS * 8, 256));
  int parts = (int)ceil_div(m > 0 ? m : 1, 256 * 8);
  const int cap = sm_count() * 8 < LOSS_MAX_PARTS ? sm_count() * 8 : LOSS_MAX_PARTS;
  if (parts > cap) parts = cap;
  k_loss_partial<<<parts, 256, 0, st>>>(pred, target, lab, w, sup, m, kind, part_sum, part_cnt);
  B2G_LAUNCH_CHECK();
  if (grad && m > 0) {
    k_loss_grad<<<ew_grid(m), 256, 0, st>>>(pred, target, lab, w, sup, m, kind, part_sum, part_cnt, parts, loss, grad);
    B2G_LAUNCH_CHECK();
  } else {
    k_loss_final<<<1, 32, 0, st>>>(part_sum, part_cnt, parts, loss);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}

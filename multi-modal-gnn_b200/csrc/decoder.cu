// Fused edge decoder: EdgeRegressionHead([64, 32]) over (patient, lab) pairs -- model.py:305-333,373-386.
//
//   z1 = dropout(relu(U[p] + V[l]))          U = h_p W1[:, :d]^T  [N_p, 64],  V = h_l W1[:, d:]^T + b1  [N_l, 64]
//   z2 = dropout(relu(W2 z1 + b2))           W2 [32, 64]
//   y  = w3 . z2 + b3
//
// Forward: one thread per pair, W2 broadcast from shared memory; nothing of size [M, 64] is written.
// Backward: pairs whose upstream gradient is exactly 0 (the 80 % unsupervised pairs, train.py:366-368) are
// compacted away (stable, deterministic); each remaining pair recomputes its forward, writes its 64-wide
// gradient row g (consumed by the per-patient / per-lab segmented reducers) and the CTA accumulates
// dW2 / db2 / dw3 / db3 in registers across its tiles; a second stage adds the per-CTA partials in order.
#include "common.cuh"

namespace {
using namespace b2g;

constexpr int H1 = 64, H2 = 32;
constexpr int DEC_THREADS = 256;
constexpr int PART_STRIDE = H2 * H1 + 8 * (H2 + H2 + 1);  // dW2 | per-warp {db2[32], dw3[32], db3}

__device__ __forceinline__ void load_z1(const float* __restrict__ U, const float* __restrict__ V, int64_t p, int64_t l, float p_drop,
                                        uint64_t seed, uint64_t sid1, int64_t pair, float (&z)[H1]) {
  const float4* u4 = reinterpret_cast<const float4*>(U + (size_t)p * H1);
  const float4* v4 = reinterpret_cast<const float4*>(V + (size_t)l * H1);
#pragma unroll
  for (int q = 0; q < H1 / 4; ++q) {
    float4 a = __ldg(u4 + q), b = __ldg(v4 + q);
    float4 r = make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
    if (p_drop > 0.f) {
      float4 mk = dropout_scale4(seed, sid1, (uint64_t)pair * (H1 / 4) + q, p_drop);
      r.x *= mk.x; r.y *= mk.y; r.z *= mk.z; r.w *= mk.w;
    }
    z[4 * q] = r.x; z[4 * q + 1] = r.y; z[4 * q + 2] = r.z; z[4 * q + 3] = r.w;
  }
}

// a2[j..j+3] (pre-activation of layer 2) for one group of 4 outputs
__device__ __forceinline__ void layer2_group(const float (*sW2)[H1], const float* sb2, int jg, const float (&z)[H1], float (&a)[4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) a[t] = sb2[jg * 4 + t];
#pragma unroll
  for (int q = 0; q < H1 / 4; ++q) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float4 w = *reinterpret_cast<const float4*>(&sW2[jg * 4 + t][4 * q]);
      a[t] = fmaf(w.x, z[4 * q], a[t]);
      a[t] = fmaf(w.y, z[4 * q + 1], a[t]);
      a[t] = fmaf(w.z, z[4 * q + 2], a[t]);
      a[t] = fmaf(w.w, z[4 * q + 3], a[t]);
    }
  }
}

__global__ void __launch_bounds__(DEC_THREADS, 2) k_decoder_fwd(const float* __restrict__ U, const float* __restrict__ V,
                                                            const int64_t* __restrict__ pi, const int64_t* __restrict__ li,
                                                            const float* __restrict__ W2, const float* __restrict__ b2,
                                                            const float* __restrict__ w3, const float* __restrict__ b3, int64_t M,
                                                            float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                                            float* __restrict__ pred) {
  __shared__ __align__(16) float sW2[H2][H1];
  __shared__ float sb2[H2], sw3[H2];
  for (int i = threadIdx.x; i < H2 * H1; i += DEC_THREADS) (&sW2[0][0])[i] = W2[i];
  if (threadIdx.x < H2) {
    sb2[threadIdx.x] = b2[threadIdx.x];
    sw3[threadIdx.x] = w3[threadIdx.x];
  }
  __syncthreads();
  if (p_drop > 0.f) {
    resolve_seed(seed, sid1);
    sid2 &= ~SEED_IS_POINTER;
  }
  const float bias3 = __ldg(b3);
  for (int64_t i = (int64_t)blockIdx.x * DEC_THREADS + threadIdx.x; i < M; i += (int64_t)gridDim.x * DEC_THREADS) {
    float z[H1];
    load_z1(U, V, __ldg(pi + i), __ldg(li + i), p_drop, seed, sid1, i, z);
    float out = bias3;
#pragma unroll 1
    for (int jg = 0; jg < H2 / 4; ++jg) {
      float a[4];
      layer2_group(sW2, sb2, jg, z, a);
      float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p_drop > 0.f) mk = dropout_scale4(seed, sid2, (uint64_t)i * (H2 / 4) + jg, p_drop);
      out = fmaf(sw3[jg * 4 + 0], fmaxf(a[0], 0.f) * mk.x, out);
      out = fmaf(sw3[jg * 4 + 1], fmaxf(a[1], 0.f) * mk.y, out);
      out = fmaf(sw3[jg * 4 + 2], fmaxf(a[2], 0.f) * mk.z, out);
      out = fmaf(sw3[jg * 4 + 3], fmaxf(a[3], 0.f) * mk.w, out);
    }
    pred[i] = out;
  }
}

// ---- stable compaction of the pairs with a non-zero upstream gradient ------------------------------------------------
constexpr int CMP_TILE = 2048;  // 256 threads x 8

__global__ void __launch_bounds__(256) k_count_nonzero(const float* __restrict__ g, int64_t M, int32_t* __restrict__ block_cnt) {
  __shared__ int wc[8];
  int64_t base = (int64_t)blockIdx.x * CMP_TILE + (int64_t)threadIdx.x * 8;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (base + k < M && g[base + k] != 0.f) ++c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += wc[w];
    block_cnt[blockIdx.x] = t;
  }
}

// single-block exclusive scan over the block counts (n_blocks <= a few thousand); out[n] = total
__global__ void __launch_bounds__(1024) k_scan_blocks(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    int i = c0 + threadIdx.x;
    int v = (i < n) ? in[i] : 0;
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
      int s = wsum[lane], si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    int carry = carry_s;
    if (i < n) out[i] = carry + wsum[w] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wsum[31] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
}

__global__ void __launch_bounds__(256) k_compact_nonzero(const float* __restrict__ g, int64_t M, const int32_t* __restrict__ block_off,
                                                         int32_t* __restrict__ ids, float* __restrict__ flags) {
  __shared__ int wbase[8];
  int64_t base = (int64_t)blockIdx.x * CMP_TILE + (int64_t)threadIdx.x * 8;
  bool nz[8];
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    nz[k] = (base + k < M) && (g[base + k] != 0.f);
    c += nz[k];
    if (base + k < M && flags) flags[base + k] = nz[k] ? 1.f : 0.f;
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wbase[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int ww = 0; ww < 8; ++ww) {
      int t = wbase[ww];
      wbase[ww] = run;
      run += t;
    }
  }
  __syncthreads();
  int pos = block_off[blockIdx.x] + wbase[w] + inc - c;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (nz[k]) ids[pos++] = (int32_t)(base + k);
}

// ---- backward ------------------------------------------------------------------------------------------------------
constexpr int ZS = H1 + 4;   // padded row strides: conflict-free 16-byte / 4-byte per-thread-row accesses
constexpr int AS = H2 + 1;
// dynamic smem: sZ [256][68] (z1 after dropout) + sA [256][33] (d loss / d a2)  ~ 101 KB
__global__ void __launch_bounds__(DEC_THREADS) k_decoder_bwd(const float* __restrict__ U, const float* __restrict__ V,
                                                            const int64_t* __restrict__ pi, const int64_t* __restrict__ li,
                                                            const float* __restrict__ W2, const float* __restrict__ b2,
                                                            const float* __restrict__ w3, const float* __restrict__ dpred,
                                                            const int32_t* __restrict__ ids, const int32_t* __restrict__ n_active_ptr,
                                                            float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                                            float* __restrict__ g_out, float* __restrict__ partial) {
  extern __shared__ __align__(16) float dyn[];
  float* sZ = dyn;                          // [256][ZS]
  float* sA = dyn + DEC_THREADS * ZS;       // [256][AS]
  __shared__ __align__(16) float sW2[H2][H1];
  __shared__ float sb2[H2], sw3[H2];
  for (int i = threadIdx.x; i < H2 * H1; i += DEC_THREADS) (&sW2[0][0])[i] = W2[i];
  if (threadIdx.x < H2) {
    sb2[threadIdx.x] = b2[threadIdx.x];
    sw3[threadIdx.x] = w3[threadIdx.x];
  }
  __syncthreads();
  if (p_drop > 0.f) {
    resolve_seed(seed, sid1);
    sid2 &= ~SEED_IS_POINTER;
  }
  const int n_active = *n_active_ptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const int oj = tid >> 3, okb = (tid & 7) * 8;  // this thread owns dW2[oj][okb .. okb+8)
  float accW[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) accW[t] = 0.f;
  float acc_b2 = 0.f, acc_w3 = 0.f, acc_b3 = 0.f;  // lane j of each warp holds column j

  for (int t0 = blockIdx.x * DEC_THREADS; t0 < n_active; t0 += gridDim.x * DEC_THREADS) {
    const int slot = t0 + tid;
    const bool live = slot < n_active;
    float z[H1];
    float dy = 0.f;
    int64_t pair = 0;
    if (live) {
      pair = ids[slot];
      dy = dpred[pair];
      load_z1(U, V, __ldg(pi + pair), __ldg(li + pair), p_drop, seed, sid1, pair, z);
    } else {
#pragma unroll
      for (int k = 0; k < H1; ++k) z[k] = 0.f;
    }
    unsigned long long zpos = 0ull;           // bit k: z1[k] > 0 (relu active AND kept by dropout)
#pragma unroll
    for (int k = 0; k < H1; ++k) zpos |= (unsigned long long)(z[k] > 0.f) << k;
#pragma unroll
    for (int q = 0; q < H1 / 4; ++q)
      *reinterpret_cast<float4*>(&sZ[tid * ZS + 4 * q]) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
    // layer 2 recompute + its local gradient (kept in this thread's sA row)
#pragma unroll 1
    for (int jg = 0; jg < H2 / 4; ++jg) {
      float a[4];
      layer2_group(sW2, sb2, jg, z, a);
      float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p_drop > 0.f) mk = dropout_scale4(seed, sid2, (uint64_t)pair * (H2 / 4) + jg, p_drop);
      float mks[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = jg * 4 + t;
        float z2d = fmaxf(a[t], 0.f) * mks[t];
        float d = (live && a[t] > 0.f) ? dy * sw3[j] * mks[t] : 0.f;   // d loss / d a2[j]
        sA[tid * AS + j] = d;
        float s_b2 = warp_sum(d);                                      // db2[j] += d, dw3[j] += dy * z2d
        float s_w3 = warp_sum(live ? dy * z2d : 0.f);
        if (lane == j) {
          acc_b2 += s_b2;
          acc_w3 += s_w3;
        }
      }
    }
    {
      float s = warp_sum(dy);
      if (lane == 0) acc_b3 += s;
    }
    // dz1 = W2^T da2, masked by relu' and the layer-1 dropout mask; written as this pair's gradient row g
    if (live) {
      float4* grow = reinterpret_cast<float4*>(g_out + (size_t)pair * H1);
#pragma unroll 1
      for (int kc = 0; kc < H1 / 16; ++kc) {
        float dz[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) dz[t] = 0.f;
#pragma unroll
        for (int j = 0; j < H2; ++j) {
          const float aj = sA[tid * AS + j];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 w = *reinterpret_cast<const float4*>(&sW2[j][kc * 16 + 4 * q]);
            dz[4 * q] = fmaf(w.x, aj, dz[4 * q]);
            dz[4 * q + 1] = fmaf(w.y, aj, dz[4 * q + 1]);
            dz[4 * q + 2] = fmaf(w.z, aj, dz[4 * q + 2]);
            dz[4 * q + 3] = fmaf(w.w, aj, dz[4 * q + 3]);
          }
        }
        const unsigned bits = (unsigned)(zpos >> (kc * 16)) & 0xffffu;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o;
          o.x = (bits >> (4 * q)) & 1u ? dz[4 * q] * keep_scale : 0.f;
          o.y = (bits >> (4 * q + 1)) & 1u ? dz[4 * q + 1] * keep_scale : 0.f;
          o.z = (bits >> (4 * q + 2)) & 1u ? dz[4 * q + 2] * keep_scale : 0.f;
          o.w = (bits >> (4 * q + 3)) & 1u ? dz[4 * q + 3] * keep_scale : 0.f;
          grow[kc * 4 + q] = o;
        }
      }
    }
    // accumulate dW2 += da2^T z1 over the tile with a [32 x 8] thread grid of 1 x 8 output tiles
    __syncthreads();
    const int n_tile = min(DEC_THREADS, n_active - t0);
#pragma unroll 4
    for (int p = 0; p < n_tile; ++p) {
      float a = sA[p * AS + oj];
      float4 z0 = *reinterpret_cast<const float4*>(&sZ[p * ZS + okb]);
      float4 z1 = *reinterpret_cast<const float4*>(&sZ[p * ZS + okb + 4]);
      accW[0] = fmaf(a, z0.x, accW[0]); accW[1] = fmaf(a, z0.y, accW[1]);
      accW[2] = fmaf(a, z0.z, accW[2]); accW[3] = fmaf(a, z0.w, accW[3]);
      accW[4] = fmaf(a, z1.x, accW[4]); accW[5] = fmaf(a, z1.y, accW[5]);
      accW[6] = fmaf(a, z1.z, accW[6]); accW[7] = fmaf(a, z1.w, accW[7]);
    }
    __syncthreads();   // the next tile overwrites sZ / sA
  }
  float* part = partial + (size_t)blockIdx.x * PART_STRIDE;
#pragma unroll
  for (int t = 0; t < 8; ++t) part[oj * H1 + okb + t] = accW[t];
  float* pw = part + H2 * H1 + warp * (2 * H2 + 1);
  pw[lane] = acc_b2;
  pw[H2 + lane] = acc_w3;
  if (lane == 0) pw[2 * H2] = acc_b3;
}

__global__ void __launch_bounds__(256) k_decoder_bwd_final(const float* __restrict__ partial, int n_cta, float* __restrict__ dW2,
                                                           float* __restrict__ db2, float* __restrict__ dw3, float* __restrict__ db3) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H2 * H1) {
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += partial[(size_t)c * PART_STRIDE + i];
    dW2[i] = s;
  } else if (i < H2 * H1 + 2 * H2 + 1) {
    int r = i - H2 * H1;  // 0..31 db2, 32..63 dw3, 64 db3
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c)
      for (int w = 0; w < 8; ++w) s += partial[(size_t)c * PART_STRIDE + H2 * H1 + w * (2 * H2 + 1) + r];
    if (r < H2) db2[r] = s;
    else if (r < 2 * H2) dw3[r - H2] = s;
    else db3[0] = s;
  }
}

inline int bwd_ctas() { return 2 * sm_count(); }
}  // namespace

extern "C" int b2g_decoder_fwd(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
                               const float* w3, const float* b3, int64_t m, float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                               float* pred, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (U && V && pi && li && W2 && b2 && w3 && b3 && pred)), "decoder_fwd: null pointer");
  B2G_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "decoder_fwd: dropout p must be in [0,1)");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(U) && aligned16(V), "decoder_fwd: U / V must be 16-byte aligned");
  int64_t want = ceil_div(m, DEC_THREADS);
  int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
  k_decoder_fwd<<<grid, DEC_THREADS, 0, (cudaStream_t)stream_>>>(U, V, pi, li, W2, b2, w3, b3, m, p_drop, seed, sid1, sid2, pred);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_decoder_bwd_ws_bytes(int64_t m) {
  size_t nb = (size_t)ceil_div(m > 0 ? m : 1, CMP_TILE);
  return align_up((nb + 1) * 4, 256) * 2 + align_up((size_t)(m > 0 ? m : 1) * 4, 256) + align_up((size_t)bwd_ctas() * PART_STRIDE * 4, 256);
}

extern "C" int b2g_decoder_bwd(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
                               const float* w3, const float* dpred, int64_t m, float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                               float* g_rows, float* active_flags, float* dW2, float* db2, float* dw3, float* db3, void* ws,
                               size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(m > 0 && U && V && pi && li && W2 && b2 && w3 && dpred && g_rows && active_flags && dW2 && db2 && dw3 && db3,
                "decoder_bwd: null pointer");
  B2G_CHECK_ARG(m < 2147483647LL, "decoder_bwd: more than 2^31 pairs");
  B2G_CHECK_ARG(aligned16(U) && aligned16(V) && aligned16(g_rows), "decoder_bwd: unaligned pointer");
  if (!ws || ws_bytes < b2g_decoder_bwd_ws_bytes(m)) {
    set_error("decoder_bwd: workspace too small");
    return B2G_EWS;
  }
  const int nb = (int)ceil_div(m, CMP_TILE);
  char* p = (char*)ws;
  int32_t* block_cnt = (int32_t*)p;
  p += align_up((size_t)(nb + 1) * 4, 256);
  int32_t* block_off = (int32_t*)p;
  p += align_up((size_t)(nb + 1) * 4, 256);
  int32_t* ids = (int32_t*)p;
  p += align_up((size_t)m * 4, 256);
  float* partial = (float*)p;
  k_count_nonzero<<<nb, 256, 0, st>>>(dpred, m, block_cnt);
  B2G_LAUNCH_CHECK();
  k_scan_blocks<<<1, 1024, 0, st>>>(block_cnt, nb, block_off);
  B2G_LAUNCH_CHECK();
  k_compact_nonzero<<<nb, 256, 0, st>>>(dpred, m, block_off, ids, active_flags);
  B2G_LAUNCH_CHECK();
  static bool attr_set = false;
  const size_t dyn = (size_t)DEC_THREADS * (ZS + AS) * sizeof(float);
  if (!attr_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_decoder_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    attr_set = true;
  }
  const int ctas = bwd_ctas();
  k_decoder_bwd<<<ctas, DEC_THREADS, dyn, st>>>(U, V, pi, li, W2, b2, w3, dpred, ids, block_off + nb, p_drop, seed, sid1, sid2, g_rows,
                                                 partial);
  B2G_LAUNCH_CHECK();
  k_decoder_bwd_final<<<(unsigned)ceil_div(H2 * H1 + 2 * H2 + 1, 256), 256, 0, st>>>(partial, ctas, dW2, db2, dw3, db3);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

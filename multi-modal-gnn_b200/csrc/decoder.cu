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
#include "tc_common.cuh"

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
  for (int o = 0; o < H1 / 8; ++o) {          // 8 elements = one Philox call
    float mk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (p_drop > 0.f) dropout_scale8(seed, sid1, (uint64_t)pair * (H1 / 8) + o, p_drop, mk);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 a = __ldg(u4 + 2 * o + h), b = __ldg(v4 + 2 * o + h);
      z[8 * o + 4 * h + 0] = fmaxf(a.x + b.x, 0.f) * mk[4 * h + 0];
      z[8 * o + 4 * h + 1] = fmaxf(a.y + b.y, 0.f) * mk[4 * h + 1];
      z[8 * o + 4 * h + 2] = fmaxf(a.z + b.z, 0.f) * mk[4 * h + 2];
      z[8 * o + 4 * h + 3] = fmaxf(a.w + b.w, 0.f) * mk[4 * h + 3];
    }
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

// ---- tensor-core forward (tf32 mode) ----------------------------------------------------------------------------------
// Same function as k_decoder_fwd, but the 64 -> 32 layer of a 128-pair tile runs on tcgen05.
//   gather : 8 lanes per pair -- each lane loads 32 contiguous bytes of the pair's U row and V row (a warp instruction
//            covers 4 whole 256-byte rows = 8 cache lines instead of 32 scattered lines with a thread per pair, which made
//            the L1 tag stage the bottleneck), applies ReLU and the layer-1 dropout (one Philox call = the lane's 8
//            elements) and writes two 16-byte chunks into a K-major SWIZZLE_128B shared-memory tile (row = pair);
//   MMA    : one thread issues 8 tcgen05.mma (M = 128 pairs, N = 32, K = 8, TF32 operands, fp32 accumulation in TMEM);
//   epilog : every thread reads its own accumulator row back (TMEM lane = pair): bias / ReLU / dropout / 32 -> 1 dot.
// 128 threads per CTA, ~41 KB of shared memory and 32 TMEM columns, so several CTAs share an SM and hide the gathers.
constexpr int TCD_THREADS = 256;
constexpr int TCD_MIN_CTAS = 4;

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B sub-tile
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// z1[8o .. 8o+8) of one pair = dropout(relu(U[p] + V[l])) for this lane's octet o
__device__ __forceinline__ void load_z1_octet(const float* __restrict__ U, const float* __restrict__ V, int p, int l, int o, float p_drop,
                                              uint64_t seed, uint64_t sid1, int64_t pair, float (&z)[8]) {
  const float4* u4 = reinterpret_cast<const float4*>(U + (size_t)p * H1) + 2 * o;
  const float4* v4 = reinterpret_cast<const float4*>(V + (size_t)l * H1) + 2 * o;
  const float4 a0 = __ldg(u4), a1 = __ldg(u4 + 1), b0 = __ldg(v4), b1 = __ldg(v4 + 1);
  float mk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (p_drop > 0.f) dropout_scale8(seed, sid1, (uint64_t)pair * (H1 / 8) + o, p_drop, mk);
  z[0] = fmaxf(a0.x + b0.x, 0.f) * mk[0];
  z[1] = fmaxf(a0.y + b0.y, 0.f) * mk[1];
  z[2] = fmaxf(a0.z + b0.z, 0.f) * mk[2];
  z[3] = fmaxf(a0.w + b0.w, 0.f) * mk[3];
  z[4] = fmaxf(a1.x + b1.x, 0.f) * mk[4];
  z[5] = fmaxf(a1.y + b1.y, 0.f) * mk[5];
  z[6] = fmaxf(a1.z + b1.z, 0.f) * mk[6];
  z[7] = fmaxf(a1.w + b1.w, 0.f) * mk[7];
}

// store the octet (elements 8o .. 8o+8 of tile row `row`) into the two-sub-tile K-major z1 tile.  Lanes o < 4 write
// their even chunk first, lanes o >= 4 their odd chunk first: the 8 lanes of a row then hit 8 different bank groups.
__device__ __forceinline__ void store_z1_octet(uint8_t* sZ, uint32_t row, int o, const float (&z)[8]) {
  const int first = o >> 2;
  const float4 lo = make_float4(z[0], z[1], z[2], z[3]), hi = make_float4(z[4], z[5], z[6], z[7]);
  const uint32_t cA = 2 * o + first, cB = 2 * o + 1 - first;
  *reinterpret_cast<float4*>(sZ + (cA >> 3) * 16384 + sw128(row, cA & 7)) = first ? hi : lo;
  *reinterpret_cast<float4*>(sZ + (cB >> 3) * 16384 + sw128(row, cB & 7)) = first ? lo : hi;
}

__global__ void __launch_bounds__(TCD_THREADS, TCD_MIN_CTAS) k_decoder_fwd_tc(const float* __restrict__ U, const float* __restrict__ V,
                                                                  const int64_t* __restrict__ pi, const int64_t* __restrict__ li,
                                                                  const float* __restrict__ W2, const float* __restrict__ b2,
                                                                  const float* __restrict__ w3, const float* __restrict__ b3, int64_t M,
                                                                  float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                                                  float* __restrict__ pred) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sb2[H2], sw3[H2];
  __shared__ float sPart[TILE_M];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* base = dsm + ((1024u - (smem_u32(dsm) & 1023u)) & 1023u);
  uint8_t* sZ = base;                 // 2 sub-tiles [128 rows x 128 B]   = 32 KB   (z1: A operand)
  uint8_t* sW = base + 2 * 16384;     // 2 sub-tiles [ 32 rows x 128 B]   =  8 KB   (W2: B operand, N = 32 rows, K-major)

  // W2 [32][64] -> swizzled K-major tile (once per CTA)
  for (int i = tid; i < H2 * (H1 / 4); i += TCD_THREADS) {
    const int n = i / (H1 / 4), c16 = i % (H1 / 4);              // row n, 16-byte chunk c16 (0..15) of its 64 floats
    const float4 w = __ldg(reinterpret_cast<const float4*>(W2 + n * H1) + c16);
    *reinterpret_cast<float4*>(sW + (c16 >> 3) * 4096 + sw128(n, c16 & 7)) = w;
  }
  if (tid < H2) {
    sb2[tid] = b2[tid];
    sw3[tid] = w3[tid];
  }
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (p_drop > 0.f) {
    resolve_seed(seed, sid1);
    sid2 &= ~SEED_IS_POINTER;
  }
  const float bias3 = __ldg(b3);
  const uint32_t idesc = make_idesc(H2);
  const uint32_t za = smem_u32(sZ), wa = smem_u32(sW);
  const int64_t n_tiles = (M + TILE_M - 1) / TILE_M;
  const int o = lane & 7, q = lane >> 3;                     // gather: octet o of row 16*warp + 4j + q
  const int h = warp >> 2;                                   // epilogue: accumulator columns [16h, 16h + 16) ...
  const int row = (warp & 3) * 32 + lane;                    // ... of tile row `row` (= TMEM lane)
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t phase = 0;
  // pair indices of this warp's 16 rows, fetched one tile ahead (they stream from HBM: ~1 us of latency otherwise exposed)
  int nxt_p = -1, nxt_l = 0;
  {
    const int64_t ig = (int64_t)blockIdx.x * TILE_M + warp * 16 + lane;
    if (lane < 16 && ig < M) {
      nxt_p = (int)__ldg(pi + ig);
      nxt_l = (int)__ldg(li + ig);
    }
  }
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    {  // this warp's 16 pairs: rows gathered 4 at a time by 8 lanes each
      const int my_p = nxt_p, my_l = nxt_l;
      {
        const int64_t ig = (t + gridDim.x) * TILE_M + warp * 16 + lane;
        nxt_p = -1;
        nxt_l = 0;
        if (lane < 16 && ig < M) {
          nxt_p = (int)__ldg(pi + ig);
          nxt_l = (int)__ldg(li + ig);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = 4 * j + q;
        const int p = __shfl_sync(FULL, my_p, r), l = __shfl_sync(FULL, my_l, r);
        float z[8];
        if (p >= 0) {
          load_z1_octet(U, V, p, l, o, p_drop, seed, sid1, t * TILE_M + warp * 16 + r, z);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) z[k] = 0.f;
        }
        store_z1_octet(sZ, (uint32_t)(warp * 16 + r), o, z);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); // (and order the previous tile's TMEM reads)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8)
          umma_tf32(tmem_base, make_desc(za + kb * 16384 + k8 * 32), make_desc(wa + kb * 4096 + k8 * 32), idesc, (kb | k8) != 0);
      umma_commit(&bar_mma);
    }
    __syncwarp();
    const int64_t i = t * TILE_M + row;
    const bool live = i < M;
    mbar_wait(&bar_mma, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    tmem_ld16(tmem_base + lane_base + 16 * h, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float out = h == 0 ? bias3 : 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {                // one Philox call = 8 layer-2 dropout lanes
      float mk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (p_drop > 0.f) dropout_scale8(seed, sid2, (uint64_t)(live ? i : 0) * (H2 / 8) + 2 * h + s, p_drop, mk);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int jj = 8 * s + e, j = 16 * h + jj;
        out = fmaf(sw3[j], fmaxf(__uint_as_float(r[jj]) + sb2[j], 0.f) * mk[e], out);
      }
    }
    if (h == 1) sPart[row] = out;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                             // (also: every TMEM read of this tile is done before the next MMA)
    if (h == 0 && live) pred[i] = out + sPart[row];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
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

// ---- tensor-core backward (tf32 mode) --------------------------------------------------------------------------------
// Per 128-pair tile of ACTIVE pairs, 256 threads (8 warps), all three contractions on tcgen05:
//   G  gather  : warp w gathers rows 16w..16w+15, 8 lanes per row (coalesced, see the forward kernel) -> z1 tile (K-major)
//   MMA1       : a2 = z1 W2^T                               -> TMEM cols [0, 32)
//   T  (under MMA1) thread (row, h) reads back half h of its row of z1, records relu/dropout bits and writes the half
//                   TRANSPOSED into the z1^T tile ([64 k-rows] x [128 pairs], K-major in the pair index)
//   E1 epilogue: thread (row, h) turns columns [16h, 16h+16) of a2 into da2 (bias, ReLU', dropout, dy w3): K-major da2 tile
//                and transposed da2^T tile; db2 / dw3 / db3 accumulate in registers over the CTA's tiles
//   MMA2       : dz1 = da2 W2     (B = W2^T held K-major)    -> TMEM cols [32, 96)
//   MMA3       : dW2^T += z1^T da2  (A = z1^T tile, M = 128 of which rows 0..63 are real, B = da2^T, K = 128 pairs)
//                                                            -> TMEM cols [96, 128), accumulated over ALL tiles of the CTA
//   E2 epilogue: thread (row, h) masks columns [32h, 32h+32) of dz1 and stages them in the (now free) z1 tile
//   S  store   : warp w writes rows 16w..16w+15 of the staged tile to g_out, 8 lanes per 256-byte row (coalesced)
// At the end warps 0-1 read dW2^T (TMEM lane = k) into the CTA's partial record; a second kernel adds the records in order.
constexpr int TCB_THREADS = 256;
constexpr int PART_TC = H2 * H1 + 8 * (2 * H2 + 1);


__global__ void __launch_bounds__(TCB_THREADS, 2) k_decoder_bwd_tc(const float* __restrict__ U, const float* __restrict__ V,
                                                                  const int64_t* __restrict__ pi, const int64_t* __restrict__ li,
                                                                  const float* __restrict__ W2, const float* __restrict__ b2,
                                                                  const float* __restrict__ w3, const float* __restrict__ dpred,
                                                                  const int32_t* __restrict__ ids, const int32_t* __restrict__ n_active_ptr,
                                                                  float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                                                  float* __restrict__ g_out, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sb2[H2], sw3[H2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 112 KB of tiles + < 1 KB static + 1 KB reserved = 114 KB per CTA: exactly two CTAs per SM (228 KB), so there is no
  // slack for a manual round-up; the 1024-byte alignment the swizzled tiles need is the declared one (checked below).
  uint8_t* base = dsm;
  if ((smem_u32(dsm) & 1023u) != 0u) __trap();
  uint8_t* sZ = base;                          // [2][128 x 128 B]  z1   (A of MMA1; later the dz1 staging tile)    32 KB
  uint8_t* sA = base + 32768;                  // [1][128 x 128 B]  da2  (A of MMA2)                                16 KB
  uint8_t* sW = base + 49152;                  // [2][ 32 x 128 B]  W2   [n=32 rows][k=64]  (B of MMA1)              8 KB
  uint8_t* sWT = base + 57344;                 // [1][ 64 x 128 B]  W2^T [n=64 rows][k=32]  (B of MMA2)              8 KB
  uint8_t* sZT = base + 65536;                 // [4][ 64 x 128 B]  z1^T [k rows][32 pairs] (A of MMA3)             32 KB
  uint8_t* sAT = base + 98304;                 // [4][ 32 x 128 B]  da2^T [j rows][32 pairs] (B of MMA3)            16 KB
  // MMA3 runs with M = 128: rows 64..127 of z1^T sub-tile s alias sub-tile s+1 (for s = 3: the da2^T tile).  Those rows
  // only feed accumulator lanes 64..127, which nobody reads.
  for (int i = tid; i < H2 * (H1 / 4); i += TCB_THREADS) {
    const int n = i / (H1 / 4), c16 = i % (H1 / 4);
    const float4 w = __ldg(reinterpret_cast<const float4*>(W2 + n * H1) + c16);
    *reinterpret_cast<float4*>(sW + (c16 >> 3) * 4096 + sw128(n, c16 & 7)) = w;
    const float wv[4] = {w.x, w.y, w.z, w.w};      // transposed copy: element (j = n, k) -> row k, column j of W2^T
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * c16 + e;
      *reinterpret_cast<float*>(sWT + sw128(k, n >> 2) + (n & 3) * 4) = wv[e];
    }
  }
  if (tid < H2) {
    sb2[tid] = b2[tid];
    sw3[tid] = w3[tid];
  }
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (p_drop > 0.f) {
    resolve_seed(seed, sid1);
    sid2 &= ~SEED_IS_POINTER;
  }
  const int n_active = *n_active_ptr;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  const uint32_t idesc32 = make_idesc(H2), idesc64 = make_idesc(H1);
  const uint32_t za = smem_u32(sZ), aa = smem_u32(sA), wa = smem_u32(sW), wta = smem_u32(sWT), zta = smem_u32(sZT), ata = smem_u32(sAT);
  const int h = warp >> 2;                                   // column half owned in the thread-per-row phases
  const int row = (warp & 3) * 32 + lane;                    // tile row (= TMEM lane) owned in the thread-per-row phases
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int o = lane & 7, q = lane >> 3;                     // gather / store phases: octet o of row 16*warp + 4j + q
  float acc_b2[16], acc_w3[16], acc_b3 = 0.f;                // columns 16h .. 16h+15, summed over this thread's rows
#pragma unroll
  for (int e = 0; e < 16; ++e) acc_b2[e] = acc_w3[e] = 0.f;
  uint32_t phase = 0;
  bool any_tile = false;

  for (int t0 = blockIdx.x * TILE_M; t0 < n_active; t0 += gridDim.x * TILE_M) {
    // ---- G: gather ----
    {
      int my_pair = -1, my_p = 0, my_l = 0;
      if (lane < 16) {
        const int slot = t0 + warp * 16 + lane;
        if (slot < n_active) {
          my_pair = __ldg(ids + slot);
          my_p = (int)__ldg(pi + my_pair);
          my_l = (int)__ldg(li + my_pair);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = 4 * j + q;
        const int pr = __shfl_sync(FULL, my_pair, r), p = __shfl_sync(FULL, my_p, r), l = __shfl_sync(FULL, my_l, r);
        float z[8];
        if (pr >= 0) {
          load_z1_octet(U, V, p, l, o, p_drop, seed, sid1, pr, z);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) z[k] = 0.f;
        }
        store_z1_octet(sZ, (uint32_t)(warp * 16 + r), o, z);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {                                    // MMA1: a2 = z1 W2^T
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8)
          umma_tf32(tmem_base, make_desc(za + kb * 16384 + k8 * 32), make_desc(wa + kb * 4096 + k8 * 32), idesc32, (kb | k8) != 0);
      umma_commit(&bar_mma);
    }
    __syncwarp();
    // ---- T: read back half h of this thread's z1 row, record its sign bits, write it transposed ----
    uint32_t zpos = 0u;                                // bit k: z1[32h + k] > 0 (relu active AND kept by dropout)
    {
      uint8_t* ztile = sZT + (row >> 5) * 8192;        // pair index = row: sub-tile row/32, chunk (row%32)/4, float row%4
      const uint32_t pc = (uint32_t)(row & 31) >> 2, pf = (uint32_t)(row & 3) * 4;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(sZ + h * 16384 + sw128(row, c));
        const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * c + e;
          zpos |= (uint32_t)(vs[e] > 0.f) << k;
          *reinterpret_cast<float*>(ztile + sw128(32 * h + k, pc) + pf) = vs[e];
        }
      }
    }
    const bool live = t0 + row < n_active;             // this thread's row: pair id and upstream gradient (cache hits)
    const int pair = live ? __ldg(ids + t0 + row) : -1;
    const float dy = live ? __ldg(dpred + pair) : 0.f;
    mbar_wait(&bar_mma, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- E1: da2 for columns [16h, 16h + 16) ----
    {
      uint32_t r[16];
      tmem_ld16(tmem_base + lane_base + 16 * h, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float da[16];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        float mk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (p_drop > 0.f) dropout_scale8(seed, sid2, (uint64_t)(live ? pair : 0) * (H2 / 8) + 2 * h + s, p_drop, mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int jj = 8 * s + e, j = 16 * h + jj;
          const float a = __uint_as_float(r[jj]) + sb2[j];
          const float z2d = fmaxf(a, 0.f) * mk[e];
          const float d = (live && a > 0.f) ? dy * sw3[j] * mk[e] : 0.f;   // d loss / d a2[j]
          da[jj] = d;
          acc_b2[jj] += d;
          acc_w3[jj] = fmaf(dy, z2d, acc_w3[jj]);                          // dy == 0 on dead rows
        }
      }
      if (h == 0) acc_b3 += dy;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(sA + sw128(row, 4 * h + c)) = make_float4(da[4 * c], da[4 * c + 1], da[4 * c + 2], da[4 * c + 3]);
      uint8_t* atile = sAT + (row >> 5) * 4096;
      const uint32_t pc = (uint32_t)(row & 31) >> 2, pf = (uint32_t)(row & 3) * 4;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) *reinterpret_cast<float*>(atile + sw128(16 * h + jj, pc) + pf) = da[jj];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8)                    // MMA2: dz1 = da2 W2   (B = W2^T, K = 32)
        umma_tf32(tmem_base + 32, make_desc(aa + k8 * 32), make_desc(wta + k8 * 32), idesc64, k8 != 0);
#pragma unroll
      for (int s = 0; s < 4; ++s)                       // MMA3: dW2^T += z1^T da2   (K = 128 pairs)
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8)
          umma_tf32(tmem_base + 96, make_desc(zta + s * 8192 + k8 * 32), make_desc(ata + s * 4096 + k8 * 32), idesc32,
                    (any_tile || (s | k8) != 0) ? 1u : 0u);
      umma_commit(&bar_mma);
    }
    __syncwarp();
    any_tile = true;
    mbar_wait(&bar_mma, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- E2: masked dz1 columns [32h, 32h + 32) -> staging tile (the z1 tile: MMA1 and phase T are done with it) ----
    {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + 32 + 32 * h, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 ov;
        ov.x = (zpos >> (4 * c)) & 1u ? __uint_as_float(r[4 * c]) * keep_scale : 0.f;
        ov.y = (zpos >> (4 * c + 1)) & 1u ? __uint_as_float(r[4 * c + 1]) * keep_scale : 0.f;
        ov.z = (zpos >> (4 * c + 2)) & 1u ? __uint_as_float(r[4 * c + 2]) * keep_scale : 0.f;
        ov.w = (zpos >> (4 * c + 3)) & 1u ? __uint_as_float(r[4 * c + 3]) * keep_scale : 0.f;
        *reinterpret_cast<float4*>(sZ + h * 16384 + sw128(row, c)) = ov;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- S: coalesced store of the gradient rows ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = warp * 16 + 4 * j + q;
      if (t0 + rr < n_active) {
        const int pr = __ldg(ids + t0 + rr);
        float4* grow = reinterpret_cast<float4*>(g_out + (size_t)pr * H1) + 2 * o;
        const uint32_t c0 = 2 * o, c1 = 2 * o + 1;
        grow[0] = *reinterpret_cast<const float4*>(sZ + (c0 >> 3) * 16384 + sw128(rr, c0 & 7));
        grow[1] = *reinterpret_cast<const float4*>(sZ + (c1 >> 3) * 16384 + sw128(rr, c1 & 7));
      }
    }
    __syncthreads();                                   // next tile overwrites sZ
  }

  // ---- per-CTA partial record: dW2 [32][64] | per warp {db2[32], dw3[32], db3} ----
  float* part = partial + (size_t)blockIdx.x * PART_TC;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 2) {                                      // TMEM lanes 0..63 = k, columns 96..127 = j
    uint32_t r[32];
    tmem_ld32(tmem_base + lane_base + 96, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < H2; ++j) part[j * H1 + row] = any_tile ? __uint_as_float(r[j]) : 0.f;
  }
  float* pw = part + H2 * H1 + warp * (2 * H2 + 1);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float s_b2 = warp_sum(acc_b2[e]), s_w3 = warp_sum(acc_w3[e]);
    if (lane == e) {
      pw[16 * h + e] = s_b2;
      pw[H2 + 16 * h + e] = s_w3;
      pw[16 * (1 - h) + e] = 0.f;                      // the other half of the columns belongs to warp (w ^ 4)
      pw[H2 + 16 * (1 - h) + e] = 0.f;
    }
  }
  {
    const float s3 = warp_sum(acc_b3);
    if (lane == 0) pw[2 * H2] = s3;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

// out = sum over CTAs (and, for the bias-like sums, over the per-warp slots) of the partial records: one warp per output,
// lanes stride over the records, fixed-order butterfly at the end (deterministic)
__global__ void __launch_bounds__(256) k_decoder_bwd_final(const float* __restrict__ partial, int n_cta, int stride, int n_warps,
                                                           float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dw3,
                                                           float* __restrict__ db3) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= H2 * H1 + 2 * H2 + 1) return;
  float s = 0.f;
  if (i < H2 * H1) {
    for (int c = lane; c < n_cta; c += 32) s += partial[(size_t)c * stride + i];
  } else {
    const int r = i - H2 * H1;  // 0..31 db2, 32..63 dw3, 64 db3
    for (int q = lane; q < n_cta * n_warps; q += 32)
      s += partial[(size_t)(q / n_warps) * stride + H2 * H1 + (q % n_warps) * (2 * H2 + 1) + r];
  }
  s = warp_sum(s);
  if (lane != 0) return;
  if (i < H2 * H1) dW2[i] = s;
  else if (i < H2 * H1 + H2) db2[i - H2 * H1] = s;
  else if (i < H2 * H1 + 2 * H2) dw3[i - H2 * H1 - H2] = s;
  else db3[0] = s;
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

/* tf32-mode forward: the 64 -> 32 layer on tcgen05 (TF32 operands, fp32 accumulate); same arguments and dropout streams */
extern "C" int b2g_decoder_fwd_tc(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
                                  const float* w3, const float* b3, int64_t m, float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                  float* pred, void* stream_) {
  B2G_CHECK_ARG(m >= 0 && (m == 0 || (U && V && pi && li && W2 && b2 && w3 && b3 && pred)), "decoder_fwd_tc: null pointer");
  B2G_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "decoder_fwd_tc: dropout p must be in [0,1)");
  if (m == 0) return B2G_OK;
  B2G_CHECK_ARG(aligned16(U) && aligned16(V) && aligned16(W2), "decoder_fwd_tc: U / V / W2 must be 16-byte aligned");
  const size_t dyn = 2 * 16384 + 2 * 4096 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_decoder_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    attr_set = true;
  }
  int64_t tiles = ceil_div(m, TILE_M);
  int64_t cap = (int64_t)sm_count() * TCD_MIN_CTAS;
  int grid = (int)(tiles < cap ? tiles : cap);
  k_decoder_fwd_tc<<<grid, TCD_THREADS, dyn, (cudaStream_t)stream_>>>(U, V, pi, li, W2, b2, w3, b3, m, p_drop, seed, sid1, sid2, pred);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" size_t b2g_decoder_bwd_ws_bytes(int64_t m) {
  size_t nb = (size_t)ceil_div(m > 0 ? m : 1, CMP_TILE);
  size_t part = (size_t)bwd_ctas() * PART_STRIDE;
  size_t part_tc = (size_t)3 * sm_count() * PART_TC;
  return align_up((nb + 1) * 4, 256) * 2 + align_up((size_t)(m > 0 ? m : 1) * 4, 256) + align_up((part > part_tc ? part : part_tc) * 4, 256);
}

static int decoder_bwd_impl(bool use_tc, const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
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
  if (use_tc) {
    static bool attr_tc = false;
    const size_t dyn_tc = 114688;
    if (!attr_tc) {
      B2G_CUDA(cudaFuncSetAttribute(k_decoder_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_tc));
      attr_tc = true;
    }
    const int ctas_tc = 2 * sm_count();
    k_decoder_bwd_tc<<<ctas_tc, TCB_THREADS, dyn_tc, st>>>(U, V, pi, li, W2, b2, w3, dpred, ids, block_off + nb, p_drop, seed, sid1, sid2,
                                                            g_rows, partial);
    B2G_LAUNCH_CHECK();
    k_decoder_bwd_final<<<(unsigned)ceil_div(H2 * H1 + 2 * H2 + 1, 8), 256, 0, st>>>(partial, ctas_tc, PART_TC, 8, dW2, db2, dw3, db3);
    B2G_LAUNCH_CHECK();
    return B2G_OK;
  }
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
  k_decoder_bwd_final<<<(unsigned)ceil_div(H2 * H1 + 2 * H2 + 1, 8), 256, 0, st>>>(partial, ctas, PART_STRIDE, 8, dW2, db2, dw3, db3);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_decoder_bwd(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
                               const float* w3, const float* dpred, int64_t m, float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                               float* g_rows, float* active_flags, float* dW2, float* db2, float* dw3, float* db3, void* ws,
                               size_t ws_bytes, void* stream_) {
  return decoder_bwd_impl(false, U, V, pi, li, W2, b2, w3, dpred, m, p_drop, seed, sid1, sid2, g_rows, active_flags, dW2, db2, dw3, db3, ws,
                          ws_bytes, stream_);
}

/* tf32-mode backward: both per-pair contractions (a2 = z1 W2^T, dz1 = da2 W2) on tcgen05 */
extern "C" int b2g_decoder_bwd_tc(const float* U, const float* V, const int64_t* pi, const int64_t* li, const float* W2, const float* b2,
                                  const float* w3, const float* dpred, int64_t m, float p_drop, uint64_t seed, uint64_t sid1, uint64_t sid2,
                                  float* g_rows, float* active_flags, float* dW2, float* db2, float* dw3, float* db3, void* ws,
                                  size_t ws_bytes, void* stream_) {
  return decoder_bwd_impl(true, U, V, pi, li, W2, b2, w3, dpred, m, p_drop, seed, sid1, sid2, g_rows, active_flags, dW2, db2, dw3, db3, ws,
                          ws_bytes, stream_);
}

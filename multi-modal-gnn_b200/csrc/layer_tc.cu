// One launch per direction for the patient side of a HeteroConv layer (model.py:125-131,256 -- PyG HeteroConv(aggr='sum')
// over six SAGEConv(mean) relations), on the tcgen05 tensor cores with the adjacency held as BITS.
//
// Every edge of the reference's graph is patient <-> {lab, diagnosis, medication} (graph_build.py:216-247), and the three
// vocabularies are small (a few hundred nodes), so the patient x type adjacency of all relations is one bit matrix
// `bits[N_patient, nw]` (32 types per word, ~50 bytes per patient instead of ~1.7 KB of fp32 zeros and ones or ~50 B of CSR
// that would need a gather).  With the small operands pre-multiplied on the type rows (Y_r = x_r W_l,r^T, <= a few hundred
// rows), the patient side of the layer is two contractions:
//
//   k_layer_tf32   out_p[M, N]  = [x_p | diag(s) A] . [W_root | Y]^T + b        (forward; the same kernel computes
//                                 dx_p = [dout_p | A] . [W_root^T | G]^T in backward, G = dagg_t / deg_t)
//   k_adjT_tf32    T[types, 128] = (diag(s) A)^T . x_p                          (forward: neighbour sums onto the type nodes;
//                                 backward: dY = (diag(1/deg_p) A)^T . dout_p)
//
// Both expand the bits into TF32 operand tiles in shared memory with dedicated "expander" warps (generic-proxy stores in the
// tensor core's swizzled layout + fence.proxy.async), so the adjacency costs 4 bytes per 32 potential edges of HBM traffic.
// HBM traffic per patient row: forward 512 B (x_p) + 512 B (out_p) + bits; the reduction operands (W_root | Y, 128 x K fp32)
// stream from L2 one 32-column chunk at a time next to the matching A chunk.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {
using namespace b2g;

constexpr int LY_MAXW = 24;                      // adjacency words per patient row (<= 768 type nodes over all relations)
constexpr int LY_MAXREL = 4;
constexpr int LY_THREADS = 384;                  // warps: 0 TMA, 1 MMA, 4-7 epilogue, 2-3 and 8-11 expanders (2 also owns TMEM)
constexpr int LY_MAX_STAGES = 6;
constexpr int A_CHUNK = TILE_M * KB * 4;         // [128 rows x 128 B] = 16 KB
constexpr int STG_ROW = 128 + 16;                // epilogue staging row: 32 floats + 16 B pad (conflict-free 16-byte accesses)
constexpr int STG_WARP = 32 * STG_ROW;
constexpr int STG_BYTES = 4 * STG_WARP;

// description of the bit layout: word w of a row holds columns [32 w, 32 w + 32) of the concatenated type axis.  Bits
// [0, split) of the word belong to relation relA, bits [split, 32) to relation relB (split == 32: one relation).  The host
// lays relations out so that no word holds more than one boundary.
struct BitLayout {
  int nw;
  int8_t rel_a[LY_MAXW];
  int8_t rel_b[LY_MAXW];
  int8_t split[LY_MAXW];
};

struct LayerParams {
  const uint32_t* bits;                // [m, nw]
  const float* rscale[LY_MAXREL];      // per relation: row scale [m] (1/deg of the patient for a mean) or null (1.0)
  BitLayout bl;
  const float* bias;                   // [n] or null
  float* y;                            // [m, n]
  double* stats;                       // null, or per-CTA column sums [grid][2][n] (sum, sum of squares) of y
  int64_t m;
  int n;                               // output columns: 32..256, multiple of 32
  int kx;                              // dense reduction columns (x), multiple of 32 (may be 0)
  int tmem_cols;
  int stages;
  long long* dbg_out;                  // diagnosis only (dbg & 8): per-role wait cycles of CTA 0
  int dbg;                             // diagnosis only (B2G_LAYER_DBG): 1 = skip the bit expansion, 2 = skip the B loads, 4 = skip x loads
};

__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc, bool on) {
  if (on) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

__device__ __forceinline__ float pick_scale(const float (&s)[LY_MAXREL], int r) {
  return r == 0 ? s[0] : (r == 1 ? s[1] : (r == 2 ? s[2] : s[3]));
}
// 4 consecutive adjacency entries (bits b0 .. b0+3 of `word`) as {0 | scale}
__device__ __forceinline__ uint4 expand4(uint32_t word, int b0, uint32_t sbits) {
  uint4 v;                                                   // (one LOP3 with predicate output + one SEL per element)
  v.x = (word & (1u << b0)) ? sbits : 0u;
  v.y = (word & (2u << b0)) ? sbits : 0u;
  v.z = (word & (4u << b0)) ? sbits : 0u;
  v.w = (word & (8u << b0)) ? sbits : 0u;
  return v;
}

// dynamic smem (1024-byte aligned): stages [A chunk 16 KB | B chunk n x 128 B] | epilogue staging [4 warps][32 rows x 144 B]
__global__ void __launch_bounds__(LY_THREADS, 1) k_layer_tf32(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_w,
                                                              const __grid_constant__ LayerParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[LY_MAX_STAGES], bar_empty[LY_MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_stat[4][2][32];
  __shared__ uint4 s_lut[16];                          // nibble -> four {0 | ~0} masks (a byte table + PRMT costs more ALU than it saves)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 16)
    s_lut[threadIdx.x] = make_uint4((threadIdx.x & 1) ? ~0u : 0u, (threadIdx.x & 2) ? ~0u : 0u, (threadIdx.x & 4) ? ~0u : 0u, (threadIdx.x & 8) ? ~0u : 0u);
  const int nxc = prm.kx / KB;                         // chunks fed by TMA from x
  const int nw = prm.bl.nw;
  const int nc = nxc + nw;                             // chunks per tile
  const uint32_t b_bytes = (uint32_t)prm.n * KB * 4;
  const uint32_t stage_bytes = A_CHUNK + b_bytes;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* smem_stg = base + (size_t)prm.stages * stage_bytes;
  const int64_t n_tiles = (prm.m + TILE_M - 1) / TILE_M;
  const int nst = prm.stages;
  const bool tm = (prm.dbg & 8) && blockIdx.x == 0;
  long long w0 = 0, w1 = 0;
  const long long t_begin = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < LY_MAX_STAGES; ++s) {
      mbar_init(&bar_full[s], 2);                      // TMA producer (expect_tx) + the stage's expander warp
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(prm.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer: per chunk the B columns (W_root | Y), and for the first kx/32 chunks the x columns =====
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 1;                                 // parity to wait for on bar_empty (the first pass over the ring is free)
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int c = 0; c < nc; ++c) {
          mbar_wait_t(&bar_empty[s], ph, w0, tm);
          uint8_t* st = base + (size_t)s * stage_bytes;
          const bool ldb = !(prm.dbg & 2), ldx = c < nxc && !(prm.dbg & 4);
          mbar_expect_tx(&bar_full[s], (ldb ? b_bytes : 0u) + (ldx ? (uint32_t)A_CHUNK : 0u));
          if (ldb) tma_load_2d(st + A_CHUNK, &map_w, &bar_full[s], c * KB, 0);
          if (ldx) tma_load_2d(st, &map_x, &bar_full[s], c * KB, (int)(t * TILE_M));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE thread runs the whole loop; stage / phase counters are incremental and the shared-memory
    // descriptors are formed by adding the stage offset to a descriptor with a zero start address =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc(prm.n);
      const uint64_t dhi = make_desc(0);
      const uint32_t base16 = smem_u32(base) >> 4, st16 = stage_bytes >> 4, a16 = A_CHUNK >> 4;
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait_t(&bar_tempty[a], ((it >> 1) & 1) ^ 1, w1, tm);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * prm.n);
        for (int c = 0; c < nc; ++c) {
          mbar_wait_t(&bar_full[s], ph, w0, tm);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = dhi + (base16 + (uint32_t)s * st16);
          const uint64_t db = da + a16;
#pragma unroll
          for (int k8 = 0; k8 < KB / 8; ++k8) umma_tf32(d_tmem, da + 2 * k8, db + 2 * k8, idesc, (c | k8) != 0);
          umma_commit(&bar_empty[s]);
          if (c == nc - 1) umma_commit(&bar_tfull[a]);
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 8) {
    // ===== expanders: adjacency bits -> TF32 A chunks (K-major SWIZZLE_128B: 16-byte chunk j of row r at j ^ (r & 7)) =====
    // Expander warp e is bound to pipeline stage e: it handles every chunk g = e (mod stages) -- all 128 rows of the chunk,
    // 4 rows per lane -- so up to `stages` chunks are being expanded at the same time, and it sees every phase of its stage's
    // barriers in order (a warp that skipped phases could not use parity waits).  A nibble of the word indexes a 16-entry
    // look-up table of four {0 | ~0} masks (one LDS.128 instead of ~12 ALU instructions per 4 entries).
    // For an x chunk there is nothing to expand: the warp only contributes the stage's second arrival.
    const int e = warp >= 8 ? warp - 6 : warp - 2;     // 0..5
    if (e < nst) {
      const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const uint32_t total = (uint32_t)(my_tiles * nc);
      uint32_t wn[4], san[4], sbn[4];
      int cn = e % nc;                                 // chunk-in-tile and tile counter of the chunk being prefetched
      int64_t tn = blockIdx.x + (int64_t)(e / nc) * gridDim.x;
      auto prefetch = [&](uint32_t g) {                // words + row scales of chunk g (if it is an adjacency chunk)
        if (g >= total || cn < nxc) return;
        const int k = cn - nxc;
        const float* ra = prm.rscale[prm.bl.rel_a[k]];
        const float* rb = prm.rscale[prm.bl.rel_b[k]];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t row = tn * TILE_M + i * 32 + lane;
          const bool live = row < prm.m;
          wn[i] = live ? __ldg(prm.bits + (size_t)row * nw + k) : 0u;
          san[i] = (live && ra) ? __float_as_uint(__ldg(ra + row)) : 0x3f800000u;
          sbn[i] = (live && rb) ? __float_as_uint(__ldg(rb + row)) : 0x3f800000u;
        }
      };
      prefetch((uint32_t)e);
      uint32_t ph = 1;                                 // parity to wait for on bar_empty[e]
      for (uint32_t g = (uint32_t)e; g < total; g += (uint32_t)nst, ph ^= 1) {
        const int c = cn;
        uint32_t wc[4], sac[4], sbc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { wc[i] = wn[i]; sac[i] = san[i]; sbc[i] = sbn[i]; }
        cn += nst;                                     // advance (cn, tn) to chunk g + stages
        while (cn >= nc) { cn -= nc; tn += gridDim.x; }
        prefetch(g + (uint32_t)nst);
        mbar_wait_t(&bar_empty[e], ph, w0, tm);
        if (c >= nxc && !(prm.dbg & 1)) {
          const int k = c - nxc;
          uint8_t* abase = base + (size_t)e * stage_bytes;
          const int split = prm.bl.split[k];
          if (split >= 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 32 + lane;
              uint8_t* arow = abase + r * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 m4 = s_lut[(wc[i] >> (4 * j)) & 15u];
                *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) = make_uint4(m4.x & sac[i], m4.y & sac[i], m4.z & sac[i], m4.w & sac[i]);
              }
            }
          } else {
            const uint32_t msk = (1u << split) - 1u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 32 + lane;
              uint8_t* arow = abase + r * 128;
              const uint32_t lo = wc[i] & msk, hi = wc[i] & ~msk;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 ma = s_lut[(lo >> (4 * j)) & 15u], mb = s_lut[(hi >> (4 * j)) & 15u];
                *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) =
                    make_uint4((ma.x & sac[i]) | (mb.x & sbc[i]), (ma.y & sac[i]) | (mb.y & sbc[i]), (ma.z & sac[i]) | (mb.z & sbc[i]),
                               (ma.w & sac[i]) | (mb.w & sbc[i]));
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[e]);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> staging tile -> coalesced 16-byte stores (+ bias, + BatchNorm column sums) =====
    const int q = warp & 3;
    const int so = lane & 7, sq = lane >> 3;
    double acc_s[4][4], acc_q[4][4];                   // [32-column chunk][column of this lane's quad]: sum, sum of squares
#pragma unroll                                         // (statistics are offered for n <= 128 only: register budget)
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_s[i][j] = acc_q[i][j] = 0.0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait_t(&bar_tfull[s], ph, w0, tm);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row0 = t * TILE_M + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * prm.n);
      uint8_t* stg = smem_stg + q * STG_WARP;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c0 = ci * 32;
        if (c0 < prm.n) {
          uint32_t rg[32];
          tmem_ld32(taddr + c0, rg);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(stg + lane * STG_ROW + v * 16) = make_float4(
                __uint_as_float(rg[4 * v]), __uint_as_float(rg[4 * v + 1]), __uint_as_float(rg[4 * v + 2]), __uint_as_float(rg[4 * v + 3]));
          __syncwarp();
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (prm.bias) b = __ldg(reinterpret_cast<const float4*>(prm.bias + c0) + so);
          float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + sq;
            if (row0 + rr < prm.m) {
              float4 o = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + so * 16);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              *(reinterpret_cast<float4*>(prm.y + (size_t)(row0 + rr) * prm.n + c0) + so) = o;
              ps[0] += o.x; ps[1] += o.y; ps[2] += o.z; ps[3] += o.w;
              pq[0] = fmaf(o.x, o.x, pq[0]); pq[1] = fmaf(o.y, o.y, pq[1]); pq[2] = fmaf(o.z, o.z, pq[2]); pq[3] = fmaf(o.w, o.w, pq[3]);
            }
          }
          if (ci < 4 && prm.stats) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc_s[ci & 3][e] += (double)ps[e];
              acc_q[ci & 3][e] += (double)pq[e];
            }
          }
          __syncwarp();                                  // the next chunk overwrites the staging tile
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[s]);
    }
    if (prm.stats) {
      // column totals of this CTA: lanes with the same column quad (so) hold disjoint rows -> add over sq (lane bits 3,4),
      // then over the four epilogue warps in warp order through shared memory; one fp64 record per CTA
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci * 32 < prm.n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            double a = acc_s[ci][e], b = acc_q[ci][e];
            a += __shfl_xor_sync(FULL, a, 8);  b += __shfl_xor_sync(FULL, b, 8);
            a += __shfl_xor_sync(FULL, a, 16); b += __shfl_xor_sync(FULL, b, 16);
            if (sq == 0) {
              s_stat[q][0][so * 4 + e] = a;
              s_stat[q][1][so * 4 + e] = b;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (q == 0) {
            double* rec = prm.stats + (size_t)blockIdx.x * 2 * prm.n;
            const int col = lane;
            rec[ci * 32 + col] = ((s_stat[0][0][col] + s_stat[1][0][col]) + s_stat[2][0][col]) + s_stat[3][0][col];
            rec[prm.n + ci * 32 + col] = ((s_stat[0][1][col] + s_stat[1][1][col]) + s_stat[2][1][col]) + s_stat[3][1][col];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  if (tm && lane == 0) {
    prm.dbg_out[warp * 4 + 0] = w0;
    prm.dbg_out[warp * 4 + 1] = w1;
    prm.dbg_out[warp * 4 + 2] = clock64() - t_begin;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols));
  }
}

// ---- T^T[128, NB] = X^T . (diag(s) A)   (X [M, 128] fp32 via TMA, A from bits) --------------------------------------------------
// Both operands are MN-major for the MMA (the reduction index = patient row is the slow one in memory); 32-bit MN-major
// operands use the 32-byte-atom flavour of the 128-byte swizzle (dense_tc.cu: k_wgrad_tf32): a sub-tile is [rows x 128 B]
// (32 columns), the 32-byte chunk j of row r sits at chunk j ^ (r & 3); SBO = 512 B (4 rows), LBO = one sub-tile.
constexpr int AT_THREADS = 512;                  // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 expanders
constexpr int AT_ROWS = 32;                      // reduction rows per stage
constexpr int AT_SUB = AT_ROWS * KB * 4;         // 4 KB sub-tile
constexpr int AT_BSTAGES = 2;                    // ring of expanded adjacency stages (nw x 4 KB each)
constexpr int AT_GROUP = 4;                      // expander warps per adjacency stage: group gi = stage, warp j = words j, j+4, ...
constexpr int AT_MAX_XSTAGES = 10;               // ring of X stages (16 KB each): deep, so that enough HBM bytes are in flight

struct AdjTParams {
  const uint32_t* bits;
  const float* rscale[LY_MAXREL];
  BitLayout bl;
  float* partial;                      // [grid][128][dcols]
  int64_t m;
  int nb;                              // 32 * (nw + ones): expanded adjacency (+ ones) columns
  int xb;                              // 1: a dense [m, 128] matrix (second tensor map) supplies 128 more columns IN FRONT: D = X^T [Xb | A | 1]
  int dcols;                           // 128 * xb + nb
  int tmem_cols;
  int xstages;
  int ones;                            // 1: one more 32-column block whose first column is 1 -> column sums of X
};

__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(AT_SUB >> 4) << 16;               // LBO: next group of 32 columns
  d |= (uint64_t)(512 >> 4) << 32;                  // SBO: next group of 4 reduction rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                           // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}

// dynamic smem (1024-byte aligned): X ring [xstages][(4 + 4 xb) sub-tiles x 4 KB] | adjacency ring [2][(nw + ones) sub-tiles x 4 KB]
__global__ void __launch_bounds__(AT_THREADS, 1) k_adjT_tf32(const __grid_constant__ CUtensorMap map_x,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const __grid_constant__ AdjTParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_xfull[AT_MAX_XSTAGES], bar_xempty[AT_MAX_XSTAGES], bar_bfull[AT_BSTAGES], bar_bempty[AT_BSTAGES],
      bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = prm.bl.nw;
  const uint32_t a_bytes = 4u * AT_SUB;                                // 128 columns of X
  const uint32_t xs_bytes = a_bytes * (1u + (uint32_t)prm.xb);         // one X-ring stage: X (and Xb)
  const uint32_t b_bytes = (uint32_t)(nw + prm.ones) * AT_SUB;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* bbase = base + (size_t)prm.xstages * xs_bytes;
  if (prm.ones && warp == 3) {
    // constant block: element (row r, column 0) = 1 (32-byte chunk 0 of row r sits at chunk r & 3), everything else 0.
    // Rows beyond m contribute nothing: TMA zero-fills the matching X rows.
    for (int sgi = 0; sgi < AT_BSTAGES; ++sgi) {
      uint32_t* sub = reinterpret_cast<uint32_t*>(bbase + (size_t)sgi * b_bytes + (size_t)nw * AT_SUB);
      for (int i = lane; i < AT_SUB / 4; i += 32) {
        const int r = i >> 5, w = i & 31;
        sub[i] = (w == ((r & 3) << 3)) ? 0x3f800000u : 0u;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const int64_t n_tiles = (prm.m + AT_ROWS - 1) / AT_ROWS;
  const int nxs = prm.xstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < AT_MAX_XSTAGES; ++s) {
      mbar_init(&bar_xfull[s], 1);
      mbar_init(&bar_xempty[s], 1);
    }
    for (int s = 0; s < AT_BSTAGES; ++s) {
      mbar_init(&bar_bfull[s], AT_GROUP);
      mbar_init(&bar_bempty[s], 1);
    }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(prm.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 1;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&bar_xempty[s], ph);
        mbar_expect_tx(&bar_xfull[s], xs_bytes);
        uint8_t* st = base + (size_t)s * xs_bytes;
        for (int c = 0; c < 4; ++c) tma_load_2d(st + c * AT_SUB, &map_x, &bar_xfull[s], c * KB, (int)(t * AT_ROWS));
        if (prm.xb)
          for (int c = 0; c < 4; ++c) tma_load_2d(st + a_bytes + c * AT_SUB, &map_b, &bar_xfull[s], c * KB, (int)(t * AT_ROWS));
        if (++s == nxs) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // D rows = the 128 columns of X; N is split into pieces of <= 256 columns, one MMA each per 8 reduction rows:
    // [Xb (128 columns, from the X ring)] [adjacency columns 0..255] [the rest].
    // One thread runs the loop; stage / phase counters are incremental and descriptors are formed by addition.
    if (elect_one()) {
      const int n1 = prm.nb > 256 ? 256 : prm.nb, n2 = prm.nb - n1;
      const uint32_t off0 = 128u * (uint32_t)prm.xb;
      const uint32_t idesc0 = make_idesc(128) | (1u << 15) | (1u << 16);
      const uint32_t idesc1 = make_idesc(n1) | (1u << 15) | (1u << 16);
      const uint32_t idesc2 = make_idesc(n2 > 0 ? n2 : 16) | (1u << 15) | (1u << 16);
      const uint64_t dhi = make_desc_mn32(0);
      const uint32_t xa = smem_u32(base) >> 4, ba = smem_u32(bbase) >> 4;
      const uint32_t x16 = xs_bytes >> 4, a16 = a_bytes >> 4, b16 = b_bytes >> 4;
      int sx = 0, sb_ = 0;
      uint32_t phx = 0, phb = 0;
      bool first = true;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&bar_xfull[sx], phx);
        mbar_wait(&bar_bfull[sb_], phb);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = dhi + (xa + (uint32_t)sx * x16);
        const uint64_t db = dhi + (ba + (uint32_t)sb_ * b16);
#pragma unroll
        for (int j = 0; j < AT_ROWS / 8; ++j) {          // 8 reduction rows = 1024 B = 64 descriptor units
          const uint32_t acc = !(first && j == 0);
          if (prm.xb) umma_tf32(tmem_base, da + 64 * j, da + a16 + 64 * j, idesc0, acc);
          umma_tf32(tmem_base + off0, da + 64 * j, db + 64 * j, idesc1, acc);
          if (n2 > 0) umma_tf32(tmem_base + off0 + 256, da + 64 * j, db + (8 * AT_SUB >> 4) + 64 * j, idesc2, acc);
        }
        first = false;
        umma_commit(&bar_xempty[sx]);
        umma_commit(&bar_bempty[sb_]);
        if (++sx == nxs) { sx = 0; phx ^= 1; }
        if (++sb_ == AT_BSTAGES) { sb_ = 0; phb ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else if (warp >= 8) {
    // expander group gi (4 warps) owns adjacency stage gi, i.e. the tiles it = gi, gi + 2, ...; inside the group warp j expands
    // the words j, j + 4, ... of the stage's 32 rows (lane = row): two stages are being expanded at any time
    const int gi = (warp - 8) / AT_GROUP, j4 = (warp - 8) % AT_GROUP;
    constexpr int WPE = LY_MAXW / AT_GROUP;            // words per expander warp (6)
    uint32_t wcur[WPE], wnxt[WPE];
    float scur[LY_MAXREL], snxt[LY_MAXREL];
    const int64_t tstep = (int64_t)gridDim.x * AT_BSTAGES;
    auto load_row = [&](int64_t t, uint32_t (&w)[WPE], float (&sc)[LY_MAXREL]) {
      const int64_t row = t * AT_ROWS + lane;
      const bool live = t < n_tiles && row < prm.m;
#pragma unroll
      for (int i = 0; i < WPE; ++i) {
        const int k = j4 + i * AT_GROUP;
        w[i] = (live && k < nw) ? __ldg(prm.bits + (size_t)row * nw + k) : 0u;
      }
#pragma unroll
      for (int q = 0; q < LY_MAXREL; ++q) sc[q] = (live && prm.rscale[q]) ? __ldg(prm.rscale[q] + row) : 1.0f;
    };
    const int64_t t_first = blockIdx.x + (int64_t)gi * gridDim.x;
    load_row(t_first, wcur, scur);
    uint32_t u = 0;                                    // use count of this group's stage
    for (int64_t t = t_first; t < n_tiles; t += tstep, ++u) {
      load_row(t + tstep, wnxt, snxt);
      mbar_wait(&bar_bempty[gi], (u & 1) ^ 1);
      uint8_t* bst = bbase + (size_t)gi * b_bytes;
#pragma unroll
      for (int i = 0; i < WPE; ++i) {
        const int k = j4 + i * AT_GROUP;
        if (k < nw) {
          uint8_t* brow = bst + (size_t)k * AT_SUB + lane * 128;
          const uint32_t word = wcur[i];
          const int split = prm.bl.split[k];
          const uint32_t sa = __float_as_uint(pick_scale(scur, prm.bl.rel_a[k]));
          if (split >= 32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                // 32-byte chunk j (8 columns) at chunk j ^ (row & 3)
              uint8_t* dst = brow + ((j ^ (lane & 3)) << 5);
              *reinterpret_cast<uint4*>(dst) = expand4(word, 8 * j, sa);        // (ALU expansion: a shared-memory look-up table, as in
              *reinterpret_cast<uint4*>(dst + 16) = expand4(word, 8 * j + 4, sa);   //  k_layer_tf32, measured 25 % slower in this kernel)
            }
          } else {
            const uint32_t sb = __float_as_uint(pick_scale(scur, prm.bl.rel_b[k]));
            const uint32_t msk = (1u << split) - 1u;
            const uint32_t lo = word & msk, hi = word & ~msk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint8_t* dst = brow + ((j ^ (lane & 3)) << 5);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 va = expand4(lo, 8 * j + 4 * h, sa), vb = expand4(hi, 8 * j + 4 * h, sb);
                *reinterpret_cast<uint4*>(dst + 16 * h) = make_uint4(va.x | vb.x, va.y | vb.y, va.z | vb.z, va.w | vb.w);
              }
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_bfull[gi]);
#pragma unroll
      for (int i = 0; i < WPE; ++i) wcur[i] = wnxt[i];
#pragma unroll
      for (int q = 0; q < LY_MAXREL; ++q) scur[q] = snxt[q];
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = q * 32 + lane;                                   // row of D = column of X
    float* dst_row = prm.partial + ((size_t)blockIdx.x * 128 + row) * prm.dcols;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < prm.dcols; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int v = 0; v < 8; ++v)
        *(reinterpret_cast<float4*>(dst_row + c0) + v) = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                                   __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols));
  }
}

// per-CTA records [128][dcols] added in fixed order; columns [0, off0) go to out_w[128][off0] as they are (X^T Xb = a weight
// gradient dW[out, in]), columns off0.. are written transposed and scaled: out_t[c - off0][r] = cscale[c - off0] * sum
__global__ void __launch_bounds__(256) k_adjT_reduce(const float* __restrict__ partial, int n_cta, int dcols, int off0,
                                                     const float* __restrict__ cscale, float* __restrict__ out_t, float* __restrict__ out_w) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n4 = 128 * dcols / 4;
  const int i4 = blockIdx.x * 32 + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 < n4) {
    const float4* src = reinterpret_cast<const float4*>(partial) + i4;
#pragma unroll 4
    for (int c = slice; c < n_cta; c += 8) {
      const float4 v = __ldg(src + (size_t)c * n4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sh[slice][lane] = acc;
  __syncthreads();
  if (slice != 0 || i4 >= n4) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float4 v = sh[k][lane];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int i = i4 * 4, r = i / dcols, cc = i % dcols;       // dcols % 4 == 0: the 4 outputs share row r
  if (cc < off0) {
    *reinterpret_cast<float4*>(out_w + (size_t)r * off0 + cc) = acc;
    return;
  }
  const int ct = cc - off0;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
  if (cscale) sc = __ldg(reinterpret_cast<const float4*>(cscale + ct));
  out_t[(size_t)ct * 128 + r] = acc.x * sc.x;
  out_t[(size_t)(ct + 1) * 128 + r] = acc.y * sc.y;
  out_t[(size_t)(ct + 2) * 128 + r] = acc.z * sc.z;
  out_t[(size_t)(ct + 3) * 128 + r] = acc.w * sc.w;
}

// ---- small helpers ---------------------------------------------------------------------------------------------------------
// bits[row, woff_bits/32 ...] |= 1 << column, for the neighbours of every row of a by-patient CSR (thread per row: no atomics)
__global__ void __launch_bounds__(256) k_adj_bits(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                                                  int nw, int bit_off, uint32_t* __restrict__ bits) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  uint32_t* w = bits + (size_t)row * nw;
  const int b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (int j = b; j < e; ++j) {
    const int c = bit_off + __ldg(col + j);
    w[c >> 5] |= 1u << (c & 31);
  }
}

// wcat[j, 0:kx] = w[j, :] (or w[:, j] when transposed);  wcat[j, kx + off_r + t] = tab_r[t, j] * scale_r[t];  zero elsewhere
struct CatParams {
  const float* w[LY_MAXREL];           // n_w weights are summed (HeteroConv adds the lin_r products of all relations)
  const float* bias[LY_MAXREL];        // n_b biases are summed into bias_out
  float* bias_out;
  int n_w, n_b;
  const float* tab[LY_MAXREL];
  const float* scale[LY_MAXREL];
  int rows[LY_MAXREL];
  int off[LY_MAXREL];
  int n_rel;
  int n, kx, ktot, w_transposed;
};
__global__ void __launch_bounds__(256) k_cat_weights(const __grid_constant__ CatParams p, float* __restrict__ out) {
  // thread -> (k, j) with j fastest: reads of tab[t, j] / w^T are coalesced, writes are strided (70 K elements: negligible)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n * p.ktot) return;
  const int j = idx % p.n, k = idx / p.n;
  if (k == 0 && p.bias_out) {
    float b = 0.f;
#pragma unroll
    for (int i = 0; i < LY_MAXREL; ++i)
      if (i < p.n_b) b += __ldg(p.bias[i] + j);
    p.bias_out[j] = b;
  }
  float v = 0.f;
  if (k < p.kx) {
    const size_t o = p.w_transposed ? (size_t)k * p.n + j : (size_t)j * p.kx + k;
#pragma unroll
    for (int i = 0; i < LY_MAXREL; ++i)
      if (i < p.n_w) v += __ldg(p.w[i] + o);
  } else {
    const int c = k - p.kx;
#pragma unroll
    for (int r = 0; r < LY_MAXREL; ++r)
      if (r < p.n_rel && c >= p.off[r] && c < p.off[r] + p.rows[r]) {
        const int t = c - p.off[r];
        v = __ldg(p.tab[r] + (size_t)t * p.n + j) * (p.scale[r] ? __ldg(p.scale[r] + t) : 1.f);
      }
  }
  out[(size_t)j * p.ktot + k] = v;
}

// sums[2n] (fp64) = per-CTA records [n_cta][2][n] added in CTA order
__global__ void __launch_bounds__(256) k_stats_reduce(const double* __restrict__ rec, int n_cta, int n2, double* __restrict__ sums) {
  __shared__ double sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (i < n2)
    for (int c = slice; c < n_cta; c += 8) a += rec[(size_t)c * n2 + i];       // 8 interleaved slices of the CTA records ...
  sh[slice][lane] = a;
  __syncthreads();
  if (slice == 0 && i < n2) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += sh[k][lane];                                // ... combined in fixed order
    sums[i] = a;
  }
}

int fill_layout(BitLayout* bl, const b2g_bit_layout_t* h) {
  if (!h || h->nw < 1 || h->nw > LY_MAXW) {
    set_error("adjacency bit layout: nw must be in [1, %d]", LY_MAXW);
    return B2G_EINVAL;
  }
  bl->nw = h->nw;
  for (int k = 0; k < LY_MAXW; ++k) {
    const bool live = k < h->nw;
    const int ra = live ? h->rel_a[k] : 0, rb = live ? h->rel_b[k] : 0, sp = live ? h->split[k] : 32;
    if (ra < 0 || ra >= LY_MAXREL || rb < 0 || rb >= LY_MAXREL || sp < 1 || sp > 32) {
      set_error("adjacency bit layout: word %d has rel_a=%d rel_b=%d split=%d", k, ra, rb, sp);
      return B2G_EINVAL;
    }
    bl->rel_a[k] = (int8_t)ra; bl->rel_b[k] = (int8_t)rb; bl->split[k] = (int8_t)sp;
  }
  return B2G_OK;
}

inline size_t layer_smem(int n, int stages) { return (size_t)stages * (A_CHUNK + (size_t)n * KB * 4) + STG_BYTES + 1024; }
inline int layer_stages(int n) {
  int st = LY_MAX_STAGES;
  while (st > 0 && layer_smem(n, st) > 227 * 1024) --st;
  return st;
}
}  // namespace

extern "C" int b2g_adj_bits_build(const int32_t* rowptr, const int32_t* col, int64_t n_rows, int nw, int bit_off, uint32_t* bits,
                                  void* stream_) {
  B2G_CHECK_ARG(rowptr && col && bits && n_rows > 0 && nw >= 1 && nw <= LY_MAXW && bit_off >= 0 && bit_off < 32 * nw,
                "adj_bits_build: bad args");
  k_adj_bits<<<(unsigned)ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream_>>>(rowptr, col, n_rows, nw, bit_off, bits);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_layer_cat_weights(const float* const* h_ws, int n_w, int w_transposed, const float* const* h_biases, int n_b,
                                     float* bias_out, int n, int kx, const float* const* h_tabs, const float* const* h_scales,
                                     const int* h_rows, const int* h_offs, int n_rel, int ktot, float* out, void* stream_) {
  B2G_CHECK_ARG(out && n > 0 && kx >= 0 && ktot >= kx && n_rel >= 0 && n_rel <= LY_MAXREL && n_w >= 0 && n_w <= LY_MAXREL &&
                    n_b >= 0 && n_b <= LY_MAXREL && (kx == 0 || n_w > 0) && (n_b == 0 || bias_out),
                "layer_cat_weights: bad args");
  CatParams p{};
  p.n_w = n_w; p.n_b = n_b; p.bias_out = n_b > 0 ? bias_out : nullptr;
  for (int i = 0; i < n_w; ++i) {
    B2G_CHECK_ARG(h_ws[i], "layer_cat_weights: null weight %d", i);
    p.w[i] = h_ws[i];
  }
  for (int i = 0; i < n_b; ++i) {
    B2G_CHECK_ARG(h_biases[i], "layer_cat_weights: null bias %d", i);
    p.bias[i] = h_biases[i];
  }
  p.n_rel = n_rel; p.n = n; p.kx = kx; p.ktot = ktot; p.w_transposed = w_transposed;
  for (int r = 0; r < n_rel; ++r) {
    B2G_CHECK_ARG(h_tabs[r] && h_rows[r] > 0 && h_offs[r] >= 0 && kx + h_offs[r] + h_rows[r] <= ktot, "layer_cat_weights: relation %d out of range", r);
    p.tab[r] = h_tabs[r]; p.scale[r] = h_scales ? h_scales[r] : nullptr; p.rows[r] = h_rows[r]; p.off[r] = h_offs[r];
  }
  k_cat_weights<<<(unsigned)ceil_div((int64_t)n * ktot, 256), 256, 0, (cudaStream_t)stream_>>>(p, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_layer_fwd_tc_supported(int64_t m, int n, int kx, int nw) {
  if (m < 1 || n < 32 || n > 256 || (n % 32) != 0 || kx < 0 || (kx % 32) != 0 || nw < 1 || nw > LY_MAXW) return 0;
  return layer_stages(n) >= 2 ? 1 : 0;
}
extern "C" size_t b2g_layer_stats_ws_bytes(int n) { return ((size_t)sm_count() * 2 * n + 2 * n) * sizeof(double) + 256; }

extern "C" int b2g_layer_fwd_tc(const float* x, const float* wcat, const float* bias, const uint32_t* bits,
                                const b2g_bit_layout_t* h_layout, const float* const* h_rscale, int64_t m, int n, int kx, float* y,
                                double* stat_sums, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(wcat && bits && y && h_layout && b2g_layer_fwd_tc_supported(m, n, kx, h_layout->nw) && (kx == 0 || x),
                "layer_fwd_tc: unsupported shape m=%lld n=%d kx=%d", (long long)m, n, kx);
  B2G_CHECK_ARG(aligned16(x) && aligned16(wcat) && aligned16(y) && (!bias || aligned16(bias)), "layer_fwd_tc: pointers must be 16-byte aligned");
  LayerParams prm{};
  int rc = fill_layout(&prm.bl, h_layout);
  if (rc) return rc;
  const int ktot = kx + 32 * prm.bl.nw;
  CUtensorMap map_x, map_w;
  rc = make_map(&map_x, kx > 0 ? x : wcat, kx > 0 ? m : n, kx > 0 ? kx : ktot, TILE_M);     // (kx == 0: never dereferenced)
  if (rc) return rc;
  rc = make_map(&map_w, wcat, n, ktot, n);
  if (rc) return rc;
  prm.bits = bits; prm.bias = bias; prm.y = y; prm.m = m; prm.n = n; prm.kx = kx;
  for (int r = 0; r < LY_MAXREL; ++r) prm.rscale[r] = h_rscale ? h_rscale[r] : nullptr;
  int cols = 32;
  while (cols < 2 * n) cols <<= 1;
  prm.tmem_cols = cols;
  prm.stages = layer_stages(n);
  {
    const char* e = getenv("B2G_LAYER_DBG");
    prm.dbg = e ? atoi(e) : 0;
    const char* es = getenv("B2G_LAYER_STAGES");
    if (es && atoi(es) >= 2 && atoi(es) <= prm.stages) prm.stages = atoi(es);
  }
  int64_t tiles = ceil_div(m, TILE_M);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  prm.stats = nullptr;
  if (stat_sums) {
    B2G_CHECK_ARG(n <= 128, "layer_fwd_tc: column statistics are fused for n <= 128 only (n=%d)", n);
    if (!ws || ws_bytes < b2g_layer_stats_ws_bytes(n)) {
      set_error("layer_fwd_tc: workspace too small for the column statistics");
      return B2G_EWS;
    }
    prm.stats = (double*)ws;
  }
  const size_t smem = layer_smem(n, prm.stages);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_layer_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  static long long* dbg_buf = nullptr;
  if (prm.dbg & 8) {
    if (!dbg_buf) B2G_CUDA(cudaMalloc(&dbg_buf, 64 * sizeof(long long)));
    B2G_CUDA(cudaMemsetAsync(dbg_buf, 0, 64 * sizeof(long long), st));
    prm.dbg_out = dbg_buf;
  }
  k_layer_tf32<<<grid, LY_THREADS, smem, st>>>(map_x, map_w, prm);
  B2G_LAUNCH_CHECK();
  if (prm.dbg & 8) {
    long long h[64];
    B2G_CUDA(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    B2G_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[k_layer_tf32 CTA0 cycles] total %lld | TMA wait-empty %lld | MMA wait-full %lld wait-tempty %lld | epi(w4) wait-tfull %lld | "
            "exp(w8) wait-empty(adj) %lld wait-empty(x) %lld\n", h[0 * 4 + 2], h[0], h[1 * 4], h[1 * 4 + 1], h[4 * 4], h[8 * 4], h[8 * 4 + 1]);
  }
  if (stat_sums) {
    k_stats_reduce<<<(unsigned)ceil_div(2 * n, 32), 256, 0, st>>>(prm.stats, grid, 2 * n, stat_sums);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}

namespace {
inline int adjT_xstages(int nsub, int xb) {
  const size_t left = 224 * 1024 - (size_t)AT_BSTAGES * nsub * AT_SUB;
  int xs = (int)(left / ((size_t)(4 + 4 * xb) * AT_SUB));
  return xs > AT_MAX_XSTAGES ? AT_MAX_XSTAGES : xs;
}
}  // namespace
/* supported: d = 128 and 128 * with_dense + 32 * (nw + with_colsum) <= 512 TMEM columns */
extern "C" int b2g_layer_adjT_tc_supported(int64_t m, int d, int nw) {
  if (m < 1 || d != 128 || nw < 1 || nw > 16) return 0;        // D = [128, 32 nw] fp32 must fit the 512 TMEM columns
  return adjT_xstages(nw, 0) >= 2 ? 1 : 0;
}
extern "C" size_t b2g_layer_adjT_tc_ws_bytes(int nw) { return (size_t)sm_count() * 128 * (128 + 32 * (size_t)nw) * 4 + 256; }

extern "C" int b2g_layer_adjT_tc(const float* x, const uint32_t* bits, const b2g_bit_layout_t* h_layout, const float* const* h_rscale,
                                 const float* col_scale, int64_t m, int with_colsum, const float* dense_b, float* dense_out,
                                 float* out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(x && bits && out && h_layout && b2g_layer_adjT_tc_supported(m, 128, h_layout->nw), "layer_adjT_tc: unsupported shape");
  B2G_CHECK_ARG(aligned16(x) && aligned16(out) && aligned16(ws) && (!col_scale || aligned16(col_scale)), "layer_adjT_tc: unaligned pointer");
  B2G_CHECK_ARG((dense_b == nullptr) == (dense_out == nullptr) && aligned16(dense_b) && aligned16(dense_out), "layer_adjT_tc: dense_b / dense_out");
  AdjTParams prm{};
  int rc = fill_layout(&prm.bl, h_layout);
  if (rc) return rc;
  const int nw = prm.bl.nw;
  prm.ones = with_colsum ? 1 : 0;
  prm.xb = dense_b ? 1 : 0;
  const int nsub = nw + prm.ones;
  const int nb = 32 * nsub;
  prm.dcols = 128 * prm.xb + nb;
  B2G_CHECK_ARG(prm.dcols <= 512, "layer_adjT_tc: %d output columns exceed the 512 TMEM columns", prm.dcols);
  if (!ws || ws_bytes < b2g_layer_adjT_tc_ws_bytes(nsub)) {
    set_error("layer_adjT_tc: workspace too small");
    return B2G_EWS;
  }
  CUtensorMap map_x, map_b;
  rc = make_map(&map_x, x, m, 128, AT_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&map_b, dense_b ? dense_b : x, m, 128, AT_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  prm.bits = bits; prm.partial = (float*)ws; prm.m = m; prm.nb = nb;
  for (int r = 0; r < LY_MAXREL; ++r) prm.rscale[r] = h_rscale ? h_rscale[r] : nullptr;
  int cols = 32;
  while (cols < prm.dcols) cols <<= 1;
  prm.tmem_cols = cols;
  prm.xstages = adjT_xstages(nsub, prm.xb);
  B2G_CHECK_ARG(prm.xstages >= 2, "layer_adjT_tc: shared memory too small for nw=%d", nw);
  const size_t smem = (size_t)prm.xstages * (4 + 4 * prm.xb) * AT_SUB + (size_t)AT_BSTAGES * nsub * AT_SUB + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_adjT_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int64_t tiles = ceil_div(m, AT_ROWS);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  k_adjT_tf32<<<grid, AT_THREADS, smem, st>>>(map_x, map_b, prm);
  B2G_LAUNCH_CHECK();
  k_adjT_reduce<<<(unsigned)ceil_div(128 * prm.dcols / 4, 32), 256, 0, st>>>(prm.partial, grid, prm.dcols, 128 * prm.xb, col_scale, out, dense_out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

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
// Both expand the bits into operand tiles in shared memory with dedicated "expander" warps (generic-proxy stores in the
// tensor core's swizzled layout + fence.proxy.async), so the adjacency costs 4 bytes per 32 potential edges of HBM traffic.
// k_layer_tf32 expands into fp16 {0 | 1/deg} tiles by default and multiplies them with kind::f16 MMAs into the accumulator
// that the kind::tf32 MMAs of the x part feed (b2g_layer_cat_half supplies the fp16 Y rows, a power-of-two scale per output
// column); k_adjT_tf32 expands into TF32 tiles (its other operand, X, is fp32 straight from TMA).
// HBM traffic per patient row: forward 512 B (x_p) + 512 B (out_p) + bits; the reduction operands (W_root | Y) stream from
// L2 one chunk at a time next to the matching A chunk.
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {
using namespace b2g;

constexpr int LY_MAXW = 24;                      // adjacency words per patient row (<= 768 type nodes over all relations)
constexpr int LY_MAXREL = 4;
constexpr int LY_THREADS = 384;                  // warps: 0 TMA, 1 MMA, 4-7 epilogue, 2-3 and 8-11 expanders (2 also owns TMEM)
constexpr int LY_MAX_STAGES = 6;
constexpr int A_CHUNK = TILE_M * KB * 4;         // [128 rows x 128 B] = 16 KB
constexpr int STG_WARP = 32 * 128;               // epilogue staging tile of one warp: 32 rows x 32 floats, TMA SWIZZLE_128B layout
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
  int half_adj;                        // 1: adjacency chunks are 64 columns of fp16 {0 | s} fed to kind::f16 (B rows from map_h), see below
  const float* unscale;                // half_adj: y = acc * unscale[col] + bias (the B rows carry a power-of-two scale per output column)
  long long* dbg_out;                  // diagnosis only (dbg & 8): per-role wait cycles of CTA 0
  int dbg;                             // diagnosis only (B2G_LAYER_DBG): 1 = skip the bit expansion, 2 = skip the B loads, 4 = skip x loads,
                                       // 16 = half mode with kind::tf32 MMAs on the adjacency chunks (wrong results: timing of the kind change)
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

// half mode: 32 adjacency bits -> 32 fp16 entries {0 | s} = four 16-byte pieces (piece j = bits 8j .. 8j+7), pure ALU.
// v_k = w << (7 - k) carries bit 8b + k of w in the sign position of its byte b; PRMT with the replicate-sign flag (0x8) in a
// selector nibble turns that sign into a whole byte, so one PRMT of (v_2q, v_2q+1) makes the two 16-bit masks of bits
// 8j + 2q, 8j + 2q + 1.  7 shifts + 16 PRMT + 16 AND per word, no shared-memory look-up: the expander's stores never wait
// for a load (the look-up-table form serialised LDS -> STS per piece, because the table could alias the stores).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// TF32 form: entries 4j .. 4j+3 of the word as 32-bit {0 | s}: one PRMT (sign of one byte replicated into all four) + AND each
__device__ __forceinline__ void shifts8(uint32_t w, uint32_t (&v)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = w << (7 - k);
}
__device__ __forceinline__ uint4 expand4_t(const uint32_t (&v)[8], int j, uint32_t s) {
  const uint32_t sel = (0x8u | (uint32_t)(j >> 1)) * 0x1111u;
  const int k0 = (j & 1) * 4;
  return make_uint4(prmt(v[k0], 0u, sel) & s, prmt(v[k0 + 1], 0u, sel) & s, prmt(v[k0 + 2], 0u, sel) & s, prmt(v[k0 + 3], 0u, sel) & s);
}
__device__ __forceinline__ void expand_word_h(uint32_t w, uint32_t s2, uint4 (&out)[4]) {
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = w << (7 - k);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t sel = (0x8u | j) * 0x11u | (0xCu | j) * 0x1100u;
    out[j] = make_uint4(prmt(v[0], v[1], sel) & s2, prmt(v[2], v[3], sel) & s2, prmt(v[4], v[5], sel) & s2, prmt(v[6], v[7], sel) & s2);
  }
}

// dynamic smem (1024-byte aligned): stages [A chunk 16 KB | B chunk n x 128 B] | epilogue staging [4 warps][32 rows x 128 B] |
// bias[256], unscale[256] | 2 slots [128 x nw words | 4 x 128 row scales]
__global__ void __launch_bounds__(LY_THREADS, 1) k_layer_tf32(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_w,
                                                              const __grid_constant__ CUtensorMap map_h,
                                                              const __grid_constant__ CUtensorMap map_y,
                                                              const __grid_constant__ LayerParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[LY_MAX_STAGES], bar_empty[LY_MAX_STAGES], bar_tfull[2], bar_tempty[2], bar_bits[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nxc = prm.kx / KB;                         // chunks fed by TMA from x
  const int nw = prm.bl.nw;
  const bool half_adj = prm.half_adj != 0;
  const int nc = nxc + (half_adj ? (nw + 1) / 2 : nw); // chunks per tile (a half-mode adjacency chunk covers two words = 64 columns)
  const uint32_t b_bytes = (uint32_t)prm.n * KB * 4;
  const uint32_t stage_bytes = A_CHUNK + b_bytes;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* smem_stg = base + (size_t)prm.stages * stage_bytes;
  uint8_t* smem_bu = smem_stg + STG_BYTES;
  uint8_t* smem_bits = smem_bu + 2048;                 // half mode: 2 slots of [128 x nw words | LY_MAXREL x 128 row scales]
  for (int i = threadIdx.x; i < 256; i += LY_THREADS) {
    reinterpret_cast<float*>(smem_bu)[i] = (prm.bias && i < prm.n) ? __ldg(prm.bias + i) : 0.f;
    reinterpret_cast<float*>(smem_bu)[256 + i] = (prm.unscale && i < prm.n) ? __ldg(prm.unscale + i) : 1.f;
  }
  const uint32_t bits_slot = (uint32_t)TILE_M * nw * 4 + LY_MAXREL * TILE_M * 4;
  const int64_t n_tiles = (prm.m + TILE_M - 1) / TILE_M;
  const int nst = prm.stages;
  const bool tm = (prm.dbg & 8) && blockIdx.x == 0;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
  const long long t_begin = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < LY_MAX_STAGES; ++s) {
      mbar_init(&bar_full[s], 2);                      // TMA producer (expect_tx) + the stage's expander warp
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 4);
      mbar_init(&bar_bits[s], 1);
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
      int it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        for (int cp = 0; cp < nc; ++cp) {
          const int c = cp;
          mbar_wait_t(&bar_empty[s], ph, w0, tm);
          uint8_t* st = base + (size_t)s * stage_bytes;
          const bool ldb = !(prm.dbg & 2), ldx = c < nxc && !(prm.dbg & 4);
          mbar_expect_tx(&bar_full[s], (ldb ? b_bytes : 0u) + (ldx ? (uint32_t)A_CHUNK : 0u));
          if (ldb) {
            if (half_adj && c >= nxc) tma_load_2d(st + A_CHUNK, &map_h, &bar_full[s], (c - nxc) * 64, 0);
            else tma_load_2d(st + A_CHUNK, &map_w, &bar_full[s], c * KB, 0);
          }
          if (ldx) tma_load_2d(st, &map_x, &bar_full[s], c * KB, (int)(t * TILE_M));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE thread runs the whole loop; stage / phase counters are incremental and the shared-memory
    // descriptors are formed by adding the stage offset to a descriptor with a zero start address =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc(prm.n), idesc_h = make_idesc_f16(prm.n);
      const uint64_t dhi = make_desc(0);
      const uint32_t base16 = smem_u32(base) >> 4, st16 = stage_bytes >> 4, a16 = A_CHUNK >> 4;
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait_t(&bar_tempty[a], ((it >> 1) & 1) ^ 1, w1, tm);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * prm.n);
        for (int cp = 0; cp < nc; ++cp) {
          const int c = cp;
          if (c < nxc) mbar_wait_t(&bar_full[s], ph, w0, tm);
          else mbar_wait_t(&bar_full[s], ph, w2, tm);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = dhi + (base16 + (uint32_t)s * st16);
          const uint64_t db = da + a16;
          if (half_adj && c >= nxc && !(prm.dbg & 16)) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) umma_f16(d_tmem, da + 2 * k16, db + 2 * k16, idesc_h, (cp | k16) != 0);
          } else {
#pragma unroll
            for (int k8 = 0; k8 < KB / 8; ++k8) umma_tf32(d_tmem, da + 2 * k8, db + 2 * k8, idesc, (cp | k8) != 0);
          }
          umma_commit(&bar_empty[s]);
          if (cp == nc - 1) umma_commit(&bar_tfull[a]);
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 8) {
    // ===== expanders: adjacency bits -> TF32 A chunks (K-major SWIZZLE_128B: 16-byte chunk j of row r at j ^ (r & 7)) =====
    // Expander warp e is bound to pipeline stage e: it handles every chunk g = e (mod stages) -- all 128 rows of the chunk,
    // 4 rows per lane -- so up to `stages` chunks are being expanded at the same time, and it sees every phase of its stage's
    // barriers in order (a warp that skipped phases could not use parity waits).  The entries come from PRMT sign
    // replication (expand4_t / expand_word_h): pure ALU, so the stores never wait for a load.
    // For an x chunk there is nothing to expand: the warp only contributes the stage's second arrival.
    const int e = warp >= 8 ? warp - 6 : warp - 2;     // 0..5
    if (e < nst) {
      const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const uint32_t total = (uint32_t)(my_tiles * nc);
      if (half_adj) {
        // half mode: chunk k covers words 2k, 2k+1 of the row = 64 columns of fp16 {0 | fp16(scale)}: the 16-byte piece j of the
        // row (8 columns = one byte of the word pair) comes from expand_word_h (PRMT sign replication) ANDed with the row scale
        // replicated in both halves of a register; half the shared-memory stores and half the MMA operand reads of the TF32 form.
        auto h2 = [](float f) -> uint32_t {
          const __half2 v = __float2half2_rn(f);
          return *reinterpret_cast<const uint32_t*>(&v);
        };
        // The words and row scales of a whole tile (128 x nw words, contiguous in global memory, + 128 floats per relation) are
        // brought into a two-slot shared-memory buffer by bulk copies that an epilogue thread issues one tile ahead
        // (bar_bits[slot]); per-lane global loads here (24 per chunk, one 32-byte sector each) cost 0.08 ms per launch.
        // A partial last tile is read with plain loads instead (the bulk copy would run past the arrays).
        int cp = e % nc, it = e / nc;                    // position in the tile and tile counter of this warp's next chunk
        int64_t t = blockIdx.x + (int64_t)it * gridDim.x;
        uint32_t ph = 1;
        for (uint32_t g = (uint32_t)e; g < total; g += (uint32_t)nst, ph ^= 1) {
          const int c = cp;
          mbar_wait_t(&bar_empty[e], ph, w0, tm);
          if (c >= nxc && !(prm.dbg & 1)) {
            const int k0 = 2 * (c - nxc);
            uint32_t wc[4][2], sc[4][4];
            if ((t + 1) * TILE_M <= prm.m) {
              mbar_wait_t(&bar_bits[it & 1], (uint32_t)(it >> 1) & 1u, w1, tm);
              const uint32_t* sbits = reinterpret_cast<const uint32_t*>(smem_bits + (size_t)(it & 1) * bits_slot);
              const float* sscale = reinterpret_cast<const float*>(smem_bits + (size_t)(it & 1) * bits_slot + (size_t)TILE_M * nw * 4);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int k = k0 + h;
                const bool kv = k < nw;
                const int ra = kv ? prm.bl.rel_a[k] : 0, rb = kv ? prm.bl.rel_b[k] : 0;
                const bool has_a = kv && prm.rscale[ra] != nullptr, has_b = kv && prm.rscale[rb] != nullptr;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int r = i * 32 + lane;
                  wc[i][h] = kv ? sbits[r * nw + k] : 0u;
                  sc[i][2 * h] = has_a ? h2(sscale[ra * TILE_M + r]) : 0x3C003C00u;
                  sc[i][2 * h + 1] = has_b ? h2(sscale[rb * TILE_M + r]) : 0x3C003C00u;
                }
              }
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int k = k0 + h;
                const bool kv = k < nw;
                const float* ra = kv ? prm.rscale[prm.bl.rel_a[k]] : nullptr;
                const float* rb = kv ? prm.rscale[prm.bl.rel_b[k]] : nullptr;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int64_t row = t * TILE_M + i * 32 + lane;
                  const bool live = kv && row < prm.m;
                  wc[i][h] = live ? __ldg(prm.bits + (size_t)row * nw + k) : 0u;
                  sc[i][2 * h] = (live && ra) ? h2(__ldg(ra + row)) : 0x3C003C00u;
                  sc[i][2 * h + 1] = (live && rb) ? h2(__ldg(rb + row)) : 0x3C003C00u;
                }
              }
            }
            uint8_t* abase = base + (size_t)e * stage_bytes;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int split = (k0 + h < nw) ? prm.bl.split[k0 + h] : 32;
              if (split >= 32) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int r = i * 32 + lane;
                  uint8_t* arow = abase + r * 128;
                  uint4 pc[4];
                  expand_word_h(wc[i][h], sc[i][2 * h], pc);
#pragma unroll
                  for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(arow + (((4 * h + j) ^ (r & 7)) << 4)) = pc[j];
                }
              } else {
                const uint32_t msk = (1u << split) - 1u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int r = i * 32 + lane;
                  uint8_t* arow = abase + r * 128;
                  uint4 pa[4], pb[4];
                  expand_word_h(wc[i][h] & msk, sc[i][2 * h], pa);
                  expand_word_h(wc[i][h] & ~msk, sc[i][2 * h + 1], pb);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(arow + (((4 * h + j) ^ (r & 7)) << 4)) =
                        make_uint4(pa[j].x | pb[j].x, pa[j].y | pb[j].y, pa[j].z | pb[j].z, pa[j].w | pb[j].w);
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_full[e]);
          cp += nst;
          while (cp >= nc) { cp -= nc; t += gridDim.x; ++it; }
        }
      } else {
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
          const bool live = row < prm.m && !(prm.dbg & 64);
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
              uint32_t v[8];
              shifts8(wc[i], v);
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) = expand4_t(v, j, sac[i]);
            }
          } else {
            const uint32_t msk = (1u << split) - 1u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 32 + lane;
              uint8_t* arow = abase + r * 128;
              uint32_t va[8], vb[8];
              shifts8(wc[i] & msk, va);
              shifts8(wc[i] & ~msk, vb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 pa = expand4_t(va, j, sac[i]), pb = expand4_t(vb, j, sbc[i]);
                *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) = make_uint4(pa.x | pb.x, pa.y | pb.y, pa.z | pb.z, pa.w | pb.w);
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[e]);
      }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers (lane = row, 32 consecutive columns) -> y = acc * unscale + bias -> the warp's 4 KB
    // staging tile in the TMA's SWIZZLE_128B layout (conflict-free 16-byte stores) -> ONE bulk tensor store per 32 x 32 block
    // (cp.async.bulk.tensor: rows beyond m are clipped by the tensor map).  The stores leave the SM without LDS / STG work;
    // the BatchNorm column sums are read back from the staging tile, lane = column. =====
    const int q = warp & 3;
    uint8_t* stg = smem_stg + q * STG_WARP;
    const float* s_bias = reinterpret_cast<const float*>(smem_bu);
    const float* s_us = s_bias + 256;
    double acc_s[4], acc_q[4];                         // column (32 ci + lane) of this warp's rows: sum, sum of squares
#pragma unroll                                         // (statistics are offered for n <= 128 only)
    for (int i = 0; i < 4; ++i) acc_s[i] = acc_q[i] = 0.0;
    // half mode: thread (warp 4, lane 0) feeds the expanders' bits / row-scale slots: tiles 0 and 1 up front, tile it + 2 into
    // slot it & 1 as soon as tile it is complete (its MMAs have committed, so every expander is done reading the slot)
    auto issue_bits = [&](int64_t t, int slot) {
      if (t >= n_tiles || (t + 1) * TILE_M > prm.m) return;
      uint8_t* dst = smem_bits + (size_t)slot * bits_slot;
      const uint32_t wbytes = (uint32_t)TILE_M * nw * 4;
      uint32_t total = wbytes;
#pragma unroll
      for (int r = 0; r < LY_MAXREL; ++r) total += prm.rscale[r] ? TILE_M * 4u : 0u;
      mbar_expect_tx(&bar_bits[slot], total);
      bulk_load(dst, prm.bits + (size_t)t * TILE_M * nw, wbytes, &bar_bits[slot]);
#pragma unroll
      for (int r = 0; r < LY_MAXREL; ++r)
        if (prm.rscale[r]) bulk_load(dst + wbytes + r * TILE_M * 4, prm.rscale[r] + t * TILE_M, TILE_M * 4u, &bar_bits[slot]);
    };
    const bool feeder = half_adj && warp == 4 && lane == 0;
    if (feeder) {
      issue_bits(blockIdx.x, 0);
      issue_bits((int64_t)blockIdx.x + gridDim.x, 1);
    }
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait_t(&bar_tfull[s], ph, w0, tm);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (feeder) issue_bits(t + 2 * (int64_t)gridDim.x, s);
      const int64_t row0 = t * TILE_M + q * 32;
      const bool live = row0 + lane < prm.m;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * prm.n);
      uint32_t rg[32];
      tmem_ld32(taddr, rg);                            // (the load of block ci + 1 is issued as soon as block ci sits in the staging tile)
      // NOT unrolled: the kernel runs five different instruction streams at once and an unrolled epilogue (8 blocks x ~600
      // instructions) made instruction fetch the epilogue's top stall (stall_no_inst), which in turn made it the critical role
#pragma unroll 1
      for (int c0 = 0; c0 < prm.n; c0 += 32) {
        float4 b[8], us[8];                            // all look-ups BEFORE the staging stores: the compiler cannot tell that the two
#pragma unroll                                         // shared-memory regions are disjoint and would serialise load -> store otherwise
        for (int v = 0; v < 8; ++v) {                  // (same address in every lane: broadcast)
          b[v] = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * v);
          us[v] = *reinterpret_cast<const float4*>(s_us + c0 + 4 * v);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // the previous block's bulk store must have finished READING the staging tile before it is overwritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          float4 o = make_float4(fmaf(__uint_as_float(rg[4 * v]), us[v].x, b[v].x), fmaf(__uint_as_float(rg[4 * v + 1]), us[v].y, b[v].y),
                                 fmaf(__uint_as_float(rg[4 * v + 2]), us[v].z, b[v].z), fmaf(__uint_as_float(rg[4 * v + 3]), us[v].w, b[v].w));
          if (!live) o = make_float4(0.f, 0.f, 0.f, 0.f);                                 // (clipped by the store; zero for the column sums)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((v ^ (lane & 7)) << 4)) = o;
        }
        if (c0 + 32 < prm.n) tmem_ld32(taddr + c0 + 32, rg);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map_y)),
                       "r"(smem_u32(stg)), "r"(c0), "r"((int)row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (prm.stats) {
          // column c0 + lane over the warp's 32 rows (conflict-free: a row's 32 words cover all banks); four independent
          // partial sums each (one chain of 32 dependent adds costs 128 cycles of latency in the critical role)
          float pa[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float v = *reinterpret_cast<const float*>(stg + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
            pa[r & 3] += v;
            pq[r & 3] = fmaf(v, v, pq[r & 3]);
          }
          const double da = (double)((pa[0] + pa[1]) + (pa[2] + pa[3])), dq = (double)((pq[0] + pq[1]) + (pq[2] + pq[3]));
          const int ci = c0 >> 5;
#pragma unroll
          for (int i = 0; i < 4; ++i) {                // (register arrays cannot be indexed by the loop counter; selects, not branches:
            acc_s[i] += (ci == i) ? da : 0.0;          //  the compiler turned predicated adds into an indirect jump)
            acc_q[i] += (ci == i) ? dq : 0.0;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[s]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // all bulk stores of this warp complete
    __syncwarp();
    if (prm.stats) {
      // per-CTA fp64 record: the four epilogue warps' column totals added in warp order through shared memory (the bits slots
      // are free by now: every tile of this CTA has been expanded and multiplied)
      double* sst = reinterpret_cast<double*>(smem_bits);              // [4 warps][2][128]
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
        if (ci * 32 < prm.n) {
          sst[(q * 2 + 0) * 128 + ci * 32 + lane] = acc_s[ci];
          sst[(q * 2 + 1) * 128 + ci * 32 + lane] = acc_q[ci];
        }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (q == 0) {
        double* rec = prm.stats + (size_t)blockIdx.x * 2 * prm.n;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci * 32 < prm.n) {
            const int col = ci * 32 + lane;
            rec[col] = ((sst[col] + sst[2 * 128 + col]) + sst[4 * 128 + col]) + sst[6 * 128 + col];
            rec[prm.n + col] = ((sst[128 + col] + sst[3 * 128 + col]) + sst[5 * 128 + col]) + sst[7 * 128 + col];
          }
      }
    }
  }
  if (tm && lane == 0) {
    prm.dbg_out[warp * 4 + 0] = w0;
    prm.dbg_out[warp * 4 + 1] = w1;
    prm.dbg_out[warp * 4 + 2] = clock64() - t_begin;
    prm.dbg_out[warp * 4 + 3] = w2;
    if (warp == 4) { prm.dbg_out[60] = w3; prm.dbg_out[61] = w4; }      // (w1 .. w4 of the epilogue: unused since its loop is rolled)
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
constexpr int AT_BSTAGES = 3;                    // ring of expanded adjacency stages (nw x 4 KB each): up to 3, prm.bstages in use
constexpr int AT_NGROUPS = 2;                    // expander groups: group g expands the tiles g, g + 2, ... of the CTA
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
  int bstages;                         // adjacency stages in use (2 or 3)
  int ones;                            // 1: one more 32-column block whose first column is 1 -> column sums of X
  int dbg;                             // diagnosis only (B2G_ADJT_DBG): 1 = skip the bit expansion, 2 = skip the X loads, 4 = skip the bits loads
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
    for (int sgi = 0; sgi < prm.bstages; ++sgi) {
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
        const bool ldx = !(prm.dbg & 2);
        mbar_expect_tx(&bar_xfull[s], ldx ? xs_bytes : 0u);
        uint8_t* st = base + (size_t)s * xs_bytes;
        if (ldx) {
          for (int c = 0; c < 4; ++c) tma_load_2d(st + c * AT_SUB, &map_x, &bar_xfull[s], c * KB, (int)(t * AT_ROWS));
          if (prm.xb)
            for (int c = 0; c < 4; ++c) tma_load_2d(st + a_bytes + c * AT_SUB, &map_b, &bar_xfull[s], c * KB, (int)(t * AT_ROWS));
        }
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
        if (++sb_ == prm.bstages) { sb_ = 0; phb ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else if (warp >= 8) {
    // expander group gi (4 warps) owns adjacency stage gi, i.e. the tiles it = gi, gi + 2, ...; inside the group warp j expands
    // the words j, j + 4, ... of the stage's 32 rows (lane = row): two stages are being expanded at any time.  (A lane -> (row, half)
    // mapping that makes the 16-byte stores bank-conflict-free -- a lane per row puts rows r and r + 4 on the same slot -- was
    // measured 15 % SLOWER, twice: the expanders are bound by their instruction stream, not by the shared-memory pipe.)
    const int gi = (warp - 8) / AT_GROUP, j4 = (warp - 8) % AT_GROUP;
    constexpr int WPE = LY_MAXW / AT_GROUP;            // words per expander warp (6)
    uint32_t wcur[WPE], wnxt[WPE];
    float scur[LY_MAXREL], snxt[LY_MAXREL];
    const int64_t tstep = (int64_t)gridDim.x * AT_NGROUPS;
    const int nbs = prm.bstages;
    auto load_row = [&](int64_t t, uint32_t (&w)[WPE], float (&sc)[LY_MAXREL]) {
      const int64_t row = t * AT_ROWS + lane;
      const bool live = t < n_tiles && row < prm.m && !(prm.dbg & 4);
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
    // tile `it` of the CTA lives in adjacency stage it % nbs (use number it / nbs of that stage).  With three stages the
    // expansion of tile it + 2 starts when the MMAs of tile it - 1 have retired, not those of tile it: the two stages that
    // are being expanded no longer wait for the one that is being multiplied (0.247 -> 0.224 ms at the C4 shard).
    int it = gi;
    for (int64_t t = t_first; t < n_tiles; t += tstep, it += AT_NGROUPS) {
      load_row(t + tstep, wnxt, snxt);
      const int sb = it % nbs;
      mbar_wait(&bar_bempty[sb], ((uint32_t)(it / nbs) & 1u) ^ 1u);
      uint8_t* bst = bbase + (size_t)sb * b_bytes;
#pragma unroll
      for (int i = 0; i < WPE; ++i) {
        const int k = j4 + i * AT_GROUP;
        if (k < nw && !(prm.dbg & 1)) {
          uint8_t* brow = bst + (size_t)k * AT_SUB + lane * 128;
          const uint32_t word = wcur[i];
          const int split = prm.bl.split[k];
          const uint32_t sa = __float_as_uint(pick_scale(scur, prm.bl.rel_a[k]));
          if (split >= 32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                // 32-byte chunk j (8 columns) at chunk j ^ (row & 3)
              uint8_t* dst = brow + ((j ^ (lane & 3)) << 5);
              *reinterpret_cast<uint4*>(dst) = expand4(word, 8 * j, sa);        // (ALU expansion: a shared-memory look-up table
              *reinterpret_cast<uint4*>(dst + 16) = expand4(word, 8 * j + 4, sa);   //  measured 25 % slower in this kernel)
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
      if (lane == 0) mbar_arrive(&bar_bfull[sb]);
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

// half mode of k_layer_tf32: the adjacency columns of row j (= output column j of the layer) as fp16 with a power-of-two
// scale 2^e_j that puts the row's largest magnitude into [2^13, 2^14) -- fp16 keeps 11 significant bits (round to nearest;
// TF32 as the tensor core reads it keeps 11 with truncation) but only 5 exponent bits, and backward rows are gradients of
// any magnitude.  The dense columns [0, kx) are multiplied by the same 2^e_j (exact), unscale[j] = 2^-e_j undoes it in the
// epilogue.  One warp per row: no atomics, no cross-CTA dependency.
__global__ void __launch_bounds__(256) k_cat_half(float* __restrict__ wcat, int n, int kx, int ktot, int kh, __half* __restrict__ whalf,
                                                  float* __restrict__ unscale) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  float* row = wcat + (size_t)j * ktot;
  float mx = 0.f;
  for (int k = kx + lane; k < ktot; k += 32) mx = fmaxf(mx, fabsf(row[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
  int e = 0;
  if (mx > 0.f && mx <= 3.0e38f) {
    int q;
    frexpf(mx, &q);                                   // mx = f 2^q, f in [0.5, 1)
    e = 14 - q;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
  }
  const float sc = ldexpf(1.f, e);
  __half* hrow = whalf + (size_t)j * kh;
  for (int c = lane; c < kh; c += 32) hrow[c] = __float2half_rn(kx + c < ktot ? row[kx + c] * sc : 0.f);
  for (int k = lane; k < kx; k += 32) row[k] *= sc;
  if (lane == 0) unscale[j] = ldexpf(1.f, -e);
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

inline size_t layer_smem(int n, int stages, int nw) {
  size_t slots = 2 * ((size_t)TILE_M * nw * 4 + LY_MAXREL * TILE_M * 4);
  if (slots < 8192) slots = 8192;                    // (the statistics exchange at the end of the kernel reuses the slots)
  return (size_t)stages * (A_CHUNK + (size_t)n * KB * 4) + STG_BYTES + 2048 + slots + 1024;
}
inline int layer_stages(int n, int nw) {
  int st = LY_MAX_STAGES;
  while (st > 0 && layer_smem(n, st, nw) > 227 * 1024 - 1024) --st;     // (1024: the kernel's static shared memory)
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

extern "C" int b2g_layer_cat_half(float* wcat, int n, int kx, int ktot, uint16_t* whalf, float* unscale, void* stream_) {
  B2G_CHECK_ARG(wcat && whalf && unscale && n > 0 && kx >= 0 && ktot > kx && ((ktot - kx) % 32) == 0 && (ktot - kx) / 32 <= LY_MAXW,
                "layer_cat_half: bad args");
  B2G_CHECK_ARG(aligned16(whalf), "layer_cat_half: whalf must be 16-byte aligned");
  const int kh = 64 * (((ktot - kx) / 32 + 1) / 2);
  k_cat_half<<<(unsigned)ceil_div(n, 8), 256, 0, (cudaStream_t)stream_>>>(wcat, n, kx, ktot, kh, reinterpret_cast<__half*>(whalf), unscale);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

extern "C" int b2g_layer_fwd_tc_supported(int64_t m, int n, int kx, int nw) {
  if (m < 1 || n < 32 || n > 256 || (n % 32) != 0 || kx < 0 || (kx % 32) != 0 || nw < 1 || nw > LY_MAXW) return 0;
  return layer_stages(n, nw) >= 2 ? 1 : 0;
}
extern "C" size_t b2g_layer_stats_ws_bytes(int n) { return ((size_t)sm_count() * 2 * n + 2 * n) * sizeof(double) + 256; }

extern "C" int b2g_layer_fwd_tc(const float* x, const float* wcat, const uint16_t* whalf, const float* unscale, const float* bias,
                                const uint32_t* bits,
                                const b2g_bit_layout_t* h_layout, const float* const* h_rscale, int64_t m, int n, int kx, float* y,
                                double* stat_sums, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(wcat && bits && y && h_layout && b2g_layer_fwd_tc_supported(m, n, kx, h_layout->nw) && (kx == 0 || x),
                "layer_fwd_tc: unsupported shape m=%lld n=%d kx=%d", (long long)m, n, kx);
  B2G_CHECK_ARG(aligned16(x) && aligned16(wcat) && aligned16(y) && (!bias || aligned16(bias)) && aligned16(bits), "layer_fwd_tc: pointers must be 16-byte aligned");
  B2G_CHECK_ARG((whalf == nullptr) == (unscale == nullptr) && aligned16(whalf) && aligned16(unscale), "layer_fwd_tc: whalf / unscale (both or neither, 16-byte aligned)");
  LayerParams prm{};
  int rc = fill_layout(&prm.bl, h_layout);
  if (rc) return rc;
  const int ktot = kx + 32 * prm.bl.nw;
  CUtensorMap map_x, map_w, map_h;
  rc = make_map(&map_x, kx > 0 ? x : wcat, kx > 0 ? m : n, kx > 0 ? kx : ktot, TILE_M);     // (kx == 0: never dereferenced)
  if (rc) return rc;
  rc = make_map(&map_w, wcat, n, ktot, n);
  if (rc) return rc;
  map_h = map_w;                                                                            // (tf32 mode: never dereferenced)
  if (whalf) {
    rc = make_map_f16(&map_h, whalf, n, 64 * ((prm.bl.nw + 1) / 2), n);
    if (rc) return rc;
  }
  CUtensorMap map_y;
  rc = make_map(&map_y, y, m, n, 32);
  if (rc) return rc;
  prm.half_adj = whalf ? 1 : 0;
  prm.unscale = unscale;
  for (int r = 0; r < LY_MAXREL && h_rscale; ++r)
    B2G_CHECK_ARG(aligned16(h_rscale[r]), "layer_fwd_tc: row scale %d must be 16-byte aligned", r);
  prm.bits = bits; prm.bias = bias; prm.y = y; prm.m = m; prm.n = n; prm.kx = kx;
  for (int r = 0; r < LY_MAXREL; ++r) prm.rscale[r] = h_rscale ? h_rscale[r] : nullptr;
  int cols = 32;
  while (cols < 2 * n) cols <<= 1;
  prm.tmem_cols = cols;
  prm.stages = layer_stages(n, prm.bl.nw);
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
  const size_t smem = layer_smem(n, prm.stages, prm.bl.nw);
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
  k_layer_tf32<<<grid, LY_THREADS, smem, st>>>(map_x, map_w, map_h, map_y, prm);
  B2G_LAUNCH_CHECK();
  if (prm.dbg & 8) {
    long long h[64];
    B2G_CUDA(cudaMemcpyAsync(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost, st));
    B2G_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[k_layer_tf32 CTA0 cycles] total %lld | TMA wait-empty %lld | MMA wait-full(x) %lld (adj) %lld wait-tempty %lld | epi(w4) wait-tfull %lld | "
            "exp(w8) wait-empty %lld wait-bits %lld | epi(w4) tcgen05.ld+wait %lld, wait_group.read %lld, fma+sts %lld, fence+syncwarp %lld\n", h[0 * 4 + 2], h[0], h[1 * 4], h[1 * 4 + 3],
            h[1 * 4 + 1], h[4 * 4], h[8 * 4], h[8 * 4 + 1], h[4 * 4 + 1], h[4 * 4 + 3], h[60], h[61]);
  }
  if (stat_sums) {
    k_stats_reduce<<<(unsigned)ceil_div(2 * n, 32), 256, 0, st>>>(prm.stats, grid, 2 * n, stat_sums);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}

namespace {
inline int adjT_xstages(int nsub, int xb, int bstages) {
  const long long left = 224 * 1024 - (long long)bstages * nsub * AT_SUB;
  if (left <= 0) return 0;
  int xs = (int)(left / ((long long)(4 + 4 * xb) * AT_SUB));
  return xs > AT_MAX_XSTAGES ? AT_MAX_XSTAGES : xs;
}
// three adjacency stages when at least 4 X stages (64 KB of loads in flight) still fit next to them
inline int adjT_bstages(int nsub, int xb) { return adjT_xstages(nsub, xb, 3) >= 4 ? 3 : 2; }
}  // namespace
/* supported: d = 128 and 128 * with_dense + 32 * (nw + with_colsum) <= 512 TMEM columns */
extern "C" int b2g_layer_adjT_tc_supported(int64_t m, int d, int nw) {
  if (m < 1 || d != 128 || nw < 1 || nw > 16) return 0;        // D = [128, 32 nw] fp32 must fit the 512 TMEM columns
  return adjT_xstages(nw, 0, 2) >= 2 ? 1 : 0;
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
  prm.bstages = adjT_bstages(nsub, prm.xb);
  {
    const char* eb = getenv("B2G_ADJT_BSTAGES");
    if (eb && (atoi(eb) == 2 || (atoi(eb) == 3 && adjT_xstages(nsub, prm.xb, 3) >= 2))) prm.bstages = atoi(eb);
  }
  prm.xstages = adjT_xstages(nsub, prm.xb, prm.bstages);
  {
    const char* e = getenv("B2G_ADJT_DBG");
    prm.dbg = e ? atoi(e) : 0;
  }
  B2G_CHECK_ARG(prm.xstages >= 2, "layer_adjT_tc: shared memory too small for nw=%d", nw);
  const size_t smem = (size_t)prm.xstages * (4 + 4 * prm.xb) * AT_SUB + (size_t)prm.bstages * nsub * AT_SUB + 1024;
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

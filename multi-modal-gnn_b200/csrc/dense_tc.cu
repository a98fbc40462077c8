// tcgen05 (5th-gen tensor core) dense layer for the large-M linears:  Y[M,N] = X[M,K] W[N,K]^T (+ bias)
// (nn.Linear / SAGEConv.lin_r on the patient rows, model.py:93-103,125-131).
//
// fp32 operands are fed to the tensor cores as TF32 (kind::tf32: the MMA reads the fp32 words in shared memory
// and ignores the low 13 mantissa bits), accumulation is fp32 in TMEM, output is fp32.  No conversion pass and no
// extra HBM traffic: the kernel is a stream over X and Y (1 KB per row for d = 128), i.e. HBM / L2 bound.
//
// Persistent, warp-specialised CTA (256 threads), one 128-row tile of X at a time:
//   warp 0   TMA producer: W once (resident for the CTA's lifetime), then X tiles into a 2-stage ring
//            (cp.async.bulk.tensor, SWIZZLE_128B boxes of 32 floats x 128 rows)
//   warp 1   MMA issuer: one elected lane issues K/8 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N<=256, K=8 each)
//            per tile into one of two TMEM accumulators; tcgen05.commit releases the smem stage and
//            publishes the accumulator
//   warp 2   TMEM allocator / deallocator
//   warps 4-7 epilogue: tcgen05.ld 32x32b (lane = row) -> registers -> a padded per-warp staging tile in shared memory ->
//            re-read with 8 lanes per 128-byte row segment -> + bias (+ previous Y when accumulating) -> 128-bit stores
//            that cover 4 whole row segments per instruction (a thread-per-row store touches 32 lines per instruction
//            and made the L1 tag stage, not HBM, the bound of this kernel)
#include "tc_common.cuh"

namespace {
using namespace b2g;

constexpr int TC_THREADS = 256;
constexpr int SUB_BYTES = TILE_M * KB * 4;      // one [128 rows x 128 B] sub-tile of X = 16 KB
constexpr int STG_ROW = 128 + 16;               // staging row: 32 floats + 16 B pad (conflict-free 16-byte accesses)
constexpr int STG_WARP = 32 * STG_ROW;          // per epilogue warp: 32 rows x 32 columns
constexpr int STG_BYTES = 4 * STG_WARP;

struct TcParams {
  const float* bias;   // [N] or null
  float* y;            // [M, N]
  int64_t m;
  int n, k;
  int accumulate;
  int tmem_cols;       // 2 * N rounded to a power of two >= 32
  int stages;          // X ring depth: 2 when it fits in shared memory, else 1
  int staged;          // epilogue through the shared-memory staging tile (coalesced stores) when it fits, else direct
  double* stats;       // null, or per-CTA fp64 column records [grid][2][n] {sum y, sum y^2} (BatchNorm statistics; n <= 128, staged)
  float* inv_norm;     // null, or [M]: the rows of y are L2-normalised in the epilogue, 1 / max(||row||, eps) is stored here
  float l2_eps;        //   (F.normalize, model.py:232; needs n <= 128 so that a row lives in one tile, staged, no accumulate)
};

// dynamic smem layout (1024-byte aligned): W sub-tiles [K/32][N rows x 128 B] | X stages [2][K/32][128 rows x 128 B] |
// epilogue staging [4 warps][32 rows x 144 B]
// EXT & 1: BatchNorm column statistics in the epilogue; EXT & 2: row L2 normalisation in the epilogue; EXT = 0: the plain
// kernel (separate instantiations: the statistics' accumulator registers cost the plain path 25 %)
template <int EXT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_linear_tf32(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w, TcParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double s_stat[4][2][32];
  __shared__ float s_bias[128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = prm.k / KB;
  const uint32_t w_bytes = (uint32_t)prm.n * prm.k * 4;
  const uint32_t x_bytes = (uint32_t)TILE_M * prm.k * 4;
  // SWIZZLE_128B tiles must start on a 1024-byte boundary of the shared window
  uint8_t* smem_w = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  uint8_t* smem_x = smem_w + w_bytes;
  uint8_t* smem_stg = smem_x + (size_t)prm.stages * x_bytes;
  const int64_t n_tiles = (prm.m + TILE_M - 1) / TILE_M;

  if ((EXT & 2) && threadIdx.x < 128) s_bias[threadIdx.x] = (prm.bias && (int)threadIdx.x < prm.n) ? prm.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], 4);      // one arrive per epilogue warp
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
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(&bar_w, w_bytes);
      for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(smem_w + (size_t)kb * prm.n * KB * 4, &map_w, &bar_w, kb * KB, 0);
      int it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % prm.stages;
        const uint32_t ph = (it / prm.stages) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);              // first use of each stage passes immediately
        mbar_expect_tx(&bar_full[s], x_bytes);
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_2d(smem_x + (size_t)s * x_bytes + (size_t)kb * SUB_BYTES, &map_x, &bar_full[s], kb * KB, (int)(t * TILE_M));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc(prm.n);
    mbar_wait(&bar_w, 0);
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it % prm.stages;                   // smem stage
      const uint32_t ph = (it / prm.stages) & 1;
      const int a = it & 1;                            // TMEM accumulator
      const uint32_t pa = (it >> 1) & 1;
      mbar_wait(&bar_tempty[a], pa ^ 1);               // epilogue has drained this accumulator
      mbar_wait(&bar_full[s], ph);                     // X tile landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * prm.n);
        const uint32_t xa = smem_u32(smem_x + (size_t)s * x_bytes);
        const uint32_t wa = smem_u32(smem_w);
        for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
          for (int k8 = 0; k8 < KB / 8; ++k8) {
            uint64_t da = make_desc(xa + kb * SUB_BYTES + k8 * 32);
            uint64_t db = make_desc(wa + kb * prm.n * KB * 4 + k8 * 32);
            umma_tf32(d_tmem, da, db, idesc, (kb | k8) != 0);
          }
        }
        umma_commit(&bar_empty[s]);                    // smem stage may be refilled once these MMAs retire
        umma_commit(&bar_tfull[a]);                    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    float acc_s[4][4], acc_q[4][4];                    // BatchNorm statistics: per-thread fp32 partials, combined in fp64
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_s[i][j] = acc_q[i][j] = 0.f;
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&bar_tfull[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row0 = t * TILE_M + q * 32;          // first row of this warp's 32-row slab
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * prm.n);
      uint8_t* stg = smem_stg + q * STG_WARP;
      const int so = lane & 7, sq = lane >> 3;             // store phase: 16-byte chunk so of row 4j + sq
      float inv_row = 1.f;                                 // fused F.normalize: this lane's row (TMEM lane = row) is read twice
      if (EXT & 2) {
        float ss = 0.f;
        for (int c0 = 0; c0 < prm.n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int v = 0; v < 32; ++v) {
            const float y = __uint_as_float(r[v]) + s_bias[c0 + v];
            ss = fmaf(y, y, ss);
          }
        }
        inv_row = 1.f / fmaxf(sqrtf(ss), prm.l2_eps);
        if (row0 + lane < prm.m) prm.inv_norm[row0 + lane] = inv_row;
      }
      // accumulate mode (out += ...): the previous values of the NEXT 32-column chunk are requested before the current
      // chunk is processed, so that two chunks of loads (16 x 16 B per lane) are in flight -- with one chunk the epilogue
      // warps' memory-level parallelism, not HBM, bounded the kernel at ~2.2 TB/s on HBM-resident outputs
      float4 oldv[8], nxtv[8];
      auto load_old = [&](int c0, float4 (&buf)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = 4 * j + sq;
          buf[j] = (row0 + rr < prm.m) ? *(reinterpret_cast<const float4*>(prm.y + (size_t)(row0 + rr) * prm.n + c0) + so)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      const bool acc_staged = prm.accumulate && prm.staged;
      if (acc_staged) load_old(0, oldv);
#pragma unroll 1
      for (int ci = 0; ci < 8; ++ci) {                     // (rolled: an unrolled epilogue made instruction fetch its top stall)
        const int c0 = ci * 32;
        if (c0 >= prm.n) break;
        if (acc_staged && c0 + 32 < prm.n) load_old(c0 + 32, nxtv);
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!prm.staged) {                                 // no room for the staging tile: thread-per-row stores
          if (row0 + lane < prm.m) {
            float* dst = prm.y + (size_t)(row0 + lane) * prm.n + c0;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                     __uint_as_float(r[4 * v + 3]));
              if (prm.bias) {
                float4 b = __ldg(reinterpret_cast<const float4*>(prm.bias + c0) + v);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (prm.accumulate) {
                float4 p = *(reinterpret_cast<const float4*>(dst) + v);
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
              }
              *(reinterpret_cast<float4*>(dst) + v) = o;
            }
          }
          continue;
        }
        if (EXT & 2) {                                     // (y + bias) * 1/||row||: the bias is added here, not in the store phase
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(stg + lane * STG_ROW + v * 16) =
                make_float4((__uint_as_float(r[4 * v]) + s_bias[c0 + 4 * v]) * inv_row, (__uint_as_float(r[4 * v + 1]) + s_bias[c0 + 4 * v + 1]) * inv_row,
                            (__uint_as_float(r[4 * v + 2]) + s_bias[c0 + 4 * v + 2]) * inv_row, (__uint_as_float(r[4 * v + 3]) + s_bias[c0 + 4 * v + 3]) * inv_row);
        } else {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(stg + lane * STG_ROW + v * 16) =
                make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        }
        __syncwarp();
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prm.bias && !(EXT & 2)) b = __ldg(reinterpret_cast<const float4*>(prm.bias + c0) + so);
        float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
        if (row0 + 32 <= prm.m) {
          // all 32 rows live (every tile but the last): no per-row branch, so the 8 staging loads are issued back to back
          // (a branch per row serialised load -> store -> load)
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = *reinterpret_cast<const float4*>(stg + (4 * j + sq) * STG_ROW + so * 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j].x += b.x; o[j].y += b.y; o[j].z += b.z; o[j].w += b.w;
            if (prm.accumulate) {
              const float4 p = oldv[j];
              o[j].x += p.x; o[j].y += p.y; o[j].z += p.z; o[j].w += p.w;
            }
            *(reinterpret_cast<float4*>(prm.y + (size_t)(row0 + 4 * j + sq) * prm.n + c0) + so) = o[j];
            if (EXT & 1) {
              ps[0] += o[j].x; ps[1] += o[j].y; ps[2] += o[j].z; ps[3] += o[j].w;
              pq[0] = fmaf(o[j].x, o[j].x, pq[0]); pq[1] = fmaf(o[j].y, o[j].y, pq[1]); pq[2] = fmaf(o[j].z, o[j].z, pq[2]);
              pq[3] = fmaf(o[j].w, o[j].w, pq[3]);
            }
          }
        } else
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rr = 4 * j + sq;
          if (row0 + rr < prm.m) {
            float4 o = *reinterpret_cast<const float4*>(stg + rr * STG_ROW + so * 16);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            float4* dst = reinterpret_cast<float4*>(prm.y + (size_t)(row0 + rr) * prm.n + c0) + so;
            if (prm.accumulate) {
              const float4 p = oldv[j];
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *dst = o;
            if (EXT & 1) {
              ps[0] += o.x; ps[1] += o.y; ps[2] += o.z; ps[3] += o.w;
              pq[0] = fmaf(o.x, o.x, pq[0]); pq[1] = fmaf(o.y, o.y, pq[1]); pq[2] = fmaf(o.z, o.z, pq[2]); pq[3] = fmaf(o.w, o.w, pq[3]);
            }
          }
        }
        if (EXT & 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i)                        // (register arrays cannot be indexed by the loop counter: predicated adds)
            if (ci == i) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc_s[i][e] += ps[e];
                acc_q[i][e] += pq[e];
              }
            }
        }
        if (acc_staged) {
#pragma unroll
          for (int j = 0; j < 8; ++j) oldv[j] = nxtv[j];
        }
        __syncwarp();                                      // the next chunk overwrites the staging tile
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[s]);
    }
    if (EXT & 1) {
      // column totals of this CTA: lanes with the same column quad hold disjoint rows -> add over lane bits 3,4, then over the
      // four epilogue warps in warp order through shared memory; one fp64 record per CTA (reduced in CTA order afterwards)
      const int so = lane & 7, sq = lane >> 3;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci * 32 < prm.n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            double a = (double)acc_s[ci][e], b = (double)acc_q[ci][e];
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
            rec[ci * 32 + lane] = ((s_stat[0][0][lane] + s_stat[1][0][lane]) + s_stat[2][0][lane]) + s_stat[3][0][lane];
            rec[prm.n + ci * 32 + lane] = ((s_stat[0][1][lane] + s_stat[1][1][lane]) + s_stat[2][1][lane]) + s_stat[3][1][lane];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(prm.tmem_cols));
  }
}

// sums[2n] (fp64) = per-CTA records [n_cta][2][n]: 8 interleaved slices of the records, combined in fixed order
__global__ void __launch_bounds__(256) k_linear_stats_reduce(const double* __restrict__ rec, int n_cta, int n2, double* __restrict__ sums) {
  __shared__ double sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (i < n2)
    for (int c = slice; c < n_cta; c += 8) a += rec[(size_t)c * n2 + i];
  sh[slice][lane] = a;
  __syncthreads();
  if (slice == 0 && i < n2) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += sh[k][lane];
    sums[i] = a;
  }
}

// ---- weight gradient on tcgen05:  D[128, NB] = A^T B  with A [M, 128], B [M, NB] row-major (reduction over the M rows) ----
// Both operands are "MN-major" for the MMA (the reduction index is the slow one in memory).  TMA lands [64 rows x 128 B]
// sub-tiles (SWIZZLE_128B); in the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) (units of 16 B) this
// is: 8 reduction rows 128 B apart form one swizzle atom, SBO = 1024 B to the next 8 rows, LBO = one sub-tile (8 KB) to
// the next 32 columns.  One MMA consumes 8 reduction rows (K = 8 for tf32): start address advances by 1024 B.
constexpr int WG_ROWS = 64;                         // reduction rows per stage
constexpr int WG_SUB = WG_ROWS * KB * 4;            // 8 KB sub-tile
constexpr int WG_STAGES = 3;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  // 32-bit operands in MN-major form need the 32-byte-atom flavour of the 128-byte swizzle (cute
  // Layout_MN_SW128_32B_Atom, Swizzle<2,5,2> on byte addresses: 32 B chunks XOR (row mod 4); TMA mode
  // SWIZZLE_128B_ATOM_32B): atom = 4 reduction rows x 128 B, SBO = 512 B to the next 4 rows, LBO = next 32 columns.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(WG_SUB >> 4) << 16;               // LBO: next group of 32 columns
  d |= (uint64_t)(512 >> 4) << 32;                  // SBO: next group of 4 reduction rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                           // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}

struct WgParams {
  float* partial;      // [grid][128][nb]
  int64_t m;
  int nb;              // columns of B (N of the MMA), multiple of 32, <= 256
  int tmem_cols;
  int stages;          // <= WG_STAGES, as many as fit in shared memory
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_wgrad_tf32(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_b, WgParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[WG_STAGES], bar_empty[WG_STAGES], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = 4u * WG_SUB;                              // 128 columns of A
  const uint32_t b_bytes = (uint32_t)(prm.nb / KB) * WG_SUB;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int64_t n_tiles = (prm.m + WG_ROWS - 1) / WG_ROWS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
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
      int it = 0;
      for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % prm.stages;
        const uint32_t ph = (it / prm.stages) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_expect_tx(&bar_full[s], stage_bytes);
        uint8_t* st = base + (size_t)s * stage_bytes;
        for (int c = 0; c < 4; ++c) tma_load_2d(st + c * WG_SUB, &map_a, &bar_full[s], c * KB, (int)(t * WG_ROWS));
        for (int c = 0; c < prm.nb / KB; ++c) tma_load_2d(st + a_bytes + c * WG_SUB, &map_b, &bar_full[s], c * KB, (int)(t * WG_ROWS));
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: as make_idesc, plus MN-major A (bit 15) and B (bit 16)
    const uint32_t idesc = make_idesc(prm.nb) | (1u << 15) | (1u << 16);
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int s = it % prm.stages;
      const uint32_t ph = (it / prm.stages) & 1;
      mbar_wait(&bar_full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t sa = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int j = 0; j < WG_ROWS / 8; ++j)
          umma_tf32(tmem_base, make_desc_mn(sa + j * 1024), make_desc_mn(sb + j * 1024), idesc, (it | j) != 0);
        umma_commit(&bar_empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar_done);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = q * 32 + lane;                                   // row of D = column of A
    float* dst_row = prm.partial + ((size_t)blockIdx.x * 128 + row) * prm.nb;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < prm.nb; c0 += 32) {
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

// out = sum over CTAs of partial[c] (fixed order); optionally transposed: partial[c] is [128][nb] = dW^T.
// A thread owns 4 consecutive outputs (16-byte, fully coalesced loads: consecutive lanes read consecutive float4 of the
// same partial record) for one of 8 slices of the CTA records; the 8 slices are combined in order through shared memory.
__global__ void __launch_bounds__(256) k_wgrad_tc_reduce(const float* __restrict__ partial, int n_cta, int rows, int cols, int transpose,
                                                         float* __restrict__ out) {
  __shared__ float4 sh[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n4 = rows * cols / 4;
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
  if (transpose) {
    const int i = i4 * 4, r = i / cols, cc = i % cols;       // cols % 4 == 0: the 4 outputs share row r
    out[(size_t)cc * rows + r] = acc.x;
    out[(size_t)(cc + 1) * rows + r] = acc.y;
    out[(size_t)(cc + 2) * rows + r] = acc.z;
    out[(size_t)(cc + 3) * rows + r] = acc.w;
  } else {
    reinterpret_cast<float4*>(out)[i4] = acc;
  }
}

}  // namespace

namespace {
__global__ void k_transpose(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)(r0 + i) * cols + c];
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8)
    if (c0 + i < cols && r < rows) out[(size_t)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}
}  // namespace

/* out[cols, rows] = in[rows, cols]^T (weights only: a few hundred KB) */
extern "C" int b2g_transpose(const float* in, int rows, int cols, float* out, void* stream_) {
  B2G_CHECK_ARG(in && out && rows > 0 && cols > 0, "transpose: bad args");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
  k_transpose<<<grid, block, 0, (cudaStream_t)stream_>>>(in, rows, cols, out);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

namespace {
inline size_t tc_smem(int n, int k, int st, bool staged) {
  return (size_t)n * k * 4 + (size_t)st * TILE_M * k * 4 + (staged ? STG_BYTES : 0) + 1024;
}
// 2 X stages if they fit next to the resident W, else 1, else unsupported (0); the epilogue staging tile is dropped before a stage is
inline int tc_stages(int n, int k, bool* staged = nullptr) {
  for (int st = 2; st >= 1; --st)
    for (int sg = 1; sg >= 0; --sg)
      if (tc_smem(n, k, st, sg != 0) <= 227 * 1024) {
        if (staged) *staged = sg != 0;
        return st;
      }
  return 0;
}
}  // namespace

extern "C" int b2g_linear_fwd_tc_supported(int64_t m, int n, int k) {
  if (m < 1 || n < 32 || n > 256 || (n % 32) != 0 || k < 32 || (k % 32) != 0) return 0;   // epilogue works in 32-column chunks
  return tc_stages(n, k) > 0 ? 1 : 0;
}

static int linear_fwd_tc_impl(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y, int accumulate,
                              double* stat_sums, void* ws, size_t ws_bytes, float* inv_norm, float l2_eps, void* stream_);

extern "C" int b2g_linear_fwd_tc(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y, int accumulate,
                                 void* stream_) {
  return linear_fwd_tc_impl(x, w, bias, m, n, k, y, accumulate, nullptr, nullptr, 0, nullptr, 0.f, stream_);
}

extern "C" size_t b2g_linear_stats_ws_bytes(int n) { return ((size_t)sm_count() * 2 * n) * sizeof(double) + 256; }

/* b2g_linear_fwd_tc with an extended epilogue (n <= 128):
 *   stat_sums != NULL: fp64 {sum y, sum y^2} per column (the statistics of the nn.BatchNorm1d that follows, model.py:93-101),
 *                      ws: b2g_linear_stats_ws_bytes(n);
 *   inv_norm  != NULL: the rows of y are L2-normalised, y /= max(||y||_2, l2_eps) (F.normalize after the last MLP linear,
 *                      model.py:103-105,232), inv_norm[m] receives the reciprocal norms (saved for backward). */
extern "C" int b2g_linear_fwd_tc_ex(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y,
                                    double* stat_sums, void* ws, size_t ws_bytes, float* inv_norm, float l2_eps, void* stream_) {
  return linear_fwd_tc_impl(x, w, bias, m, n, k, y, 0, stat_sums, ws, ws_bytes, inv_norm, l2_eps, stream_);
}

static int linear_fwd_tc_impl(const float* x, const float* w, const float* bias, int64_t m, int n, int k, float* y, int accumulate,
                              double* stat_sums, void* ws, size_t ws_bytes, float* inv_norm, float l2_eps, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(x && w && y && b2g_linear_fwd_tc_supported(m, n, k), "linear_fwd_tc: unsupported shape m=%lld n=%d k=%d", (long long)m, n, k);
  B2G_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y) && (!bias || aligned16(bias)), "linear_fwd_tc: pointers must be 16-byte aligned");
  CUtensorMap map_x, map_w;
  int rc = make_map(&map_x, x, m, k, TILE_M);
  if (rc) return rc;
  rc = make_map(&map_w, w, n, k, n);
  if (rc) return rc;
  TcParams prm;
  prm.bias = bias; prm.y = y; prm.m = m; prm.n = n; prm.k = k; prm.accumulate = accumulate;
  int cols = 32;
  while (cols < 2 * n) cols <<= 1;
  prm.tmem_cols = cols;
  bool staged = false;
  prm.stages = tc_stages(n, k, &staged);
  prm.staged = staged ? 1 : 0;
  prm.stats = nullptr; prm.inv_norm = inv_norm; prm.l2_eps = l2_eps;
  if (stat_sums || inv_norm) {
    B2G_CHECK_ARG(n <= 128 && staged && !accumulate, "linear_fwd_tc_ex: the extended epilogue needs n <= 128 (n=%d, k=%d)", n, k);
    if (stat_sums) {
      if (!ws || ws_bytes < b2g_linear_stats_ws_bytes(n)) {
        set_error("linear_fwd_tc_ex: workspace too small");
        return B2G_EWS;
      }
      prm.stats = (double*)ws;
    }
  }
  const size_t smem = tc_smem(n, k, prm.stages, staged);
  B2G_CHECK_ARG(!(stat_sums && inv_norm), "linear_fwd_tc_ex: statistics and row normalisation are separate variants");
  const int ext = stat_sums ? 1 : (inv_norm ? 2 : 0);
  static size_t smem_set[3] = {0, 0, 0};
  auto kern = ext == 1 ? k_linear_tf32<1> : (ext == 2 ? k_linear_tf32<2> : k_linear_tf32<0>);
  if (smem > smem_set[ext]) {
    B2G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[ext] = smem;
  }
  int64_t tiles = ceil_div(m, TILE_M);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, TC_THREADS, smem, st>>>(map_x, map_w, prm);
  B2G_LAUNCH_CHECK();
  if (stat_sums) {
    k_linear_stats_reduce<<<(unsigned)ceil_div(2 * n, 32), 256, 0, st>>>(prm.stats, grid, 2 * n, stat_sums);
    B2G_LAUNCH_CHECK();
  }
  return B2G_OK;
}

/* dW[N,K] = dy[M,N]^T x[M,K] on tcgen05 (TF32 operands, fp32 accumulation in TMEM across the whole M range of a CTA;
 * per-CTA partials are added in fixed order).  Supported when one of N, K is 128 and the other a multiple of 32 <= 256. */
extern "C" int b2g_linear_bwd_weight_tc_supported(int64_t m, int n, int k) {
  if (m < 1) return 0;
  if (n == 128 && k % 32 == 0 && k >= 32 && k <= 256) return 1;
  if (k == 128 && n % 32 == 0 && n >= 32 && n <= 256) return 1;
  return 0;
}
extern "C" size_t b2g_linear_bwd_weight_tc_ws_bytes(int64_t m, int n, int k) {
  (void)m;
  return (size_t)sm_count() * n * k * 4 + 256;
}
extern "C" int b2g_linear_bwd_weight_tc(const float* dy, const float* x, int64_t m, int n, int k, float* dw, void* ws, size_t ws_bytes,
                                        void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  B2G_CHECK_ARG(dy && x && dw && b2g_linear_bwd_weight_tc_supported(m, n, k), "linear_bwd_weight_tc: unsupported shape m=%lld n=%d k=%d",
                (long long)m, n, k);
  B2G_CHECK_ARG(aligned16(dy) && aligned16(x) && aligned16(dw) && aligned16(ws), "linear_bwd_weight_tc: unaligned pointer");
  if (!ws || ws_bytes < b2g_linear_bwd_weight_tc_ws_bytes(m, n, k)) {
    set_error("linear_bwd_weight_tc: workspace too small");
    return B2G_EWS;
  }
  const bool swap = (n != 128);                 // the 128-wide matrix is the MMA's A (its columns become D's 128 rows)
  const float* a = swap ? x : dy;
  const float* b = swap ? dy : x;
  const int nb = swap ? n : k;
  CUtensorMap map_a, map_b;
  int rc = make_map(&map_a, a, m, 128, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = make_map(&map_b, b, m, nb, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  WgParams prm;
  prm.partial = (float*)ws; prm.m = m; prm.nb = nb;
  int cols = 32;
  while (cols < nb) cols <<= 1;
  prm.tmem_cols = cols;
  const size_t stage_bytes = (size_t)(4 + nb / KB) * WG_SUB;
  prm.stages = (int)((226 * 1024) / stage_bytes);
  if (prm.stages > WG_STAGES) prm.stages = WG_STAGES;
  const size_t smem = (size_t)prm.stages * stage_bytes + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    B2G_CUDA(cudaFuncSetAttribute(k_wgrad_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int64_t tiles = ceil_div(m, WG_ROWS);
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  k_wgrad_tf32<<<grid, TC_THREADS, smem, st>>>(map_a, map_b, prm);
  B2G_LAUNCH_CHECK();
  // partial[c] is [128][nb]: equals dW[N,K] when !swap (rows = n), dW^T when swap (rows = k, cols = n)
  k_wgrad_tc_reduce<<<(unsigned)ceil_div(128 * nb / 4, 32), 256, 0, st>>>(prm.partial, grid, 128, nb, swap ? 1 : 0, dw);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

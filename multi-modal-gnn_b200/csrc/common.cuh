// Shared device / host helpers for libb2g (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b2g.h"

namespace b2g {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

#define B2G_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      b2g::set_error(__VA_ARGS__);              \
      return B2G_EINVAL;                        \
    }                                           \
  } while (0)

#define B2G_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      b2g::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return B2G_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

#define B2G_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    b2g::count_launch();                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      b2g::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return B2G_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// streaming (read-once) 128-bit load / store: keep L1 for the reused small tables
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

// Philox4x32-7 (Salmon et al. 2011: 7 rounds are the fewest that pass BigCrush; cuRAND's default of 10 adds margin that a
// dropout mask does not need) -- counter-based, so forward and backward regenerate the same mask.  The decoder kernels spend
// more than half of their instructions here (12 calls per pair).
constexpr int PHILOX_ROUNDS = 7;
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < PHILOX_ROUNDS; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// CUDA-graph friendly seeding: when bit 63 of the stream id is set, `seed` is a DEVICE POINTER to the 64-bit seed, so
// a captured graph can be replayed with a fresh seed written by the host before each replay.
constexpr uint64_t SEED_IS_POINTER = 1ull << 63;
__device__ __forceinline__ void resolve_seed(uint64_t& seed, uint64_t& sid) {
  if (sid & SEED_IS_POINTER) {
    seed = *reinterpret_cast<const uint64_t*>(seed);
    sid &= ~SEED_IS_POINTER;
  }
}

// Dropout keep-scales.  One Philox4x32-7 call yields 128 random bits = eight 16-bit uniforms; element e of dropout
// stream `sid` uses call counter e / 8 and 16-bit lane e % 8, and is dropped when its uniform is below
// floor(p * 65536) (keep probability quantised to 2^-16: relative error <= 1.6e-5 at p = 0.2).
__device__ __forceinline__ void dropout_scale8(uint64_t seed, uint64_t sid, uint64_t oct, float p, float (&m)[8]) {
  Philox ph(seed);
  uint4 r = ph(oct, sid);
  const float inv = 1.0f / (1.0f - p);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = ((w[i] & 0xffffu) >= thr) ? inv : 0.f;
    m[2 * i + 1] = ((w[i] >> 16) >= thr) ? inv : 0.f;
  }
}
// keep-scale for the 4 consecutive elements 4*quad .. 4*quad+3 (half of one Philox call)
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint64_t sid, uint64_t quad, float p) {
  Philox ph(seed);
  uint4 r = ph(quad >> 1, sid);
  const float inv = 1.0f / (1.0f - p);
  const uint32_t thr = (uint32_t)(p * 65536.0f);
  const uint32_t a = (quad & 1) ? r.z : r.x, b = (quad & 1) ? r.w : r.y;
  float4 m;
  m.x = ((a & 0xffffu) >= thr) ? inv : 0.f;
  m.y = ((a >> 16) >= thr) ? inv : 0.f;
  m.z = ((b & 0xffffu) >= thr) ? inv : 0.f;
  m.w = ((b >> 16) >= thr) ? inv : 0.f;
  return m;
}
// A row slice held by one lane: D/32 floats. D >= 128 -> float4 pieces at columns j*128 + lane*4 (every load
// instruction is one fully coalesced 512 B row segment); D == 64 -> float2; D == 32 -> float.
template <int D>
struct RowVec {
  static constexpr int N = D / 32;
  float v[N];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
    if constexpr (D >= 128) {
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        float4 t = __ldg(reinterpret_cast<const float4*>(row + j * 128) + lane);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    } else if constexpr (D == 64) {
      float2 t = __ldg(reinterpret_cast<const float2*>(row) + lane);
      v[0] = t.x; v[1] = t.y;
    } else {
      v[0] = __ldg(row + lane);
    }
  }
  __device__ __forceinline__ void load_rw(const float* row, int lane) {  // plain (non read-only) load
    if constexpr (D >= 128) {
#pragma unroll
      for (int j = 0; j < D / 128; ++j) {
        float4 t = *(reinterpret_cast<const float4*>(row + j * 128) + lane);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    } else if constexpr (D == 64) {
      float2 t = *(reinterpret_cast<const float2*>(row) + lane);
      v[0] = t.x; v[1] = t.y;
    } else {
      v[0] = row[lane];
    }
  }
  __device__ __forceinline__ void store(float* row, int lane) const {
    if constexpr (D >= 128) {
#pragma unroll
      for (int j = 0; j < D / 128; ++j)
        *(reinterpret_cast<float4*>(row + j * 128) + lane) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if constexpr (D == 64) {
      *(reinterpret_cast<float2*>(row) + lane) = make_float2(v[0], v[1]);
    } else {
      row[lane] = v[0];
    }
  }
  __device__ __forceinline__ void fma(float s, const RowVec& o) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = fmaf(s, o.v[i], v[i]);
  }
  __device__ __forceinline__ void add(const RowVec& o) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += o.v[i];
  }
};

#endif  // __CUDACC__

}  // namespace b2g

#define DISPATCH_D(d, ...)                                     \
  switch (d) {                                                 \
    case 32: { constexpr int D = 32; __VA_ARGS__; } break;     \
    case 64: { constexpr int D = 64; __VA_ARGS__; } break;     \
    case 128: { constexpr int D = 128; __VA_ARGS__; } break;   \
    case 256: { constexpr int D = 256; __VA_ARGS__; } break;   \
    default:                                                   \
      b2g::set_error("unsupported feature width d=%d (32/64/128/256)", d); \
      return B2G_EINVAL;                                       \
  }

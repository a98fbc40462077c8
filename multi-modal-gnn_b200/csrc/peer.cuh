// Peer-memory exchange over NVLink / NVSwitch for the patient-partitioned multi-GPU mode.
//
// Every rank owns one "symmetric region" (plain cudaMalloc, exported with CUDA IPC and mapped by every other rank of the
// node), so a kernel can store flags into and load payload from any peer directly.  The step's many tiny all-reduces
// (BatchNorm statistics 2 KB, per-layer type sums 235 KB, replicated->local gradients) are latency-bound, so they use a
// ONE-SHOT scheme per CTA-sized slice, inside the kernel that produced the payload where there is one:
//
//     write my slice into my region (parity = sequence number & 1)
//     fence.sys; store the sequence number into flag[slot][my rank] of EVERY rank's region      (signal)
//     spin until flag[slot][r] >= sequence number for every r in my own region                  (wait)
//     out = sum_r slice_r, r = 0 .. world-1   (volatile loads from the peers' regions, FIXED rank order:
//                                               every rank gets bit-identical sums)
//
// Sequence numbers live in device memory and are advanced by the kernel itself, so a captured CUDA graph can be
// replayed.  Two parities make the scheme safe without a second barrier: nobody can reach call n+2 on a slot (and
// overwrite the parity of call n) before everybody has signalled call n+1, i.e. finished reading call n.
// A spin that does not complete within the time-out (default ~30 s: ranks can be skewed by seconds while they build their graph
// indices or load modules; b2g_comm_set_timeout / B2G_PEER_TIMEOUT_S) sets the error flag and gives up instead of hanging the
// GPU; once the flag is set every later wait returns at once, so a dead peer costs one time-out, not one per exchange.  Results
// after a time-out are garbage: the host checks the flag at its next synchronisation point (Trainer.train_epoch / validate,
// bench.py) and raises.
#pragma once
#include "common.cuh"

namespace b2g {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_AR_SLOTS = 256;                    // slices of a standalone all-reduce (one CTA each)
constexpr int PEER_SLOTS = PEER_AR_SLOTS + 32;        // + slots reserved for kernels with a fused exchange
constexpr int PEER_SLOT_BN = PEER_AR_SLOTS;           // BatchNorm statistics (forward and backward)
constexpr int PEER_SLICE_BYTES = 16384;
constexpr size_t PEER_FLAG_BYTES = (size_t)PEER_SLOTS * PEER_MAX_WORLD * sizeof(uint32_t);     // 9216
constexpr size_t PEER_DATA_OFF = 16384;
constexpr size_t PEER_REGION_BYTES = PEER_DATA_OFF + (size_t)2 * PEER_SLOTS * PEER_SLICE_BYTES;

struct PeerCtx {
  int rank, world;
  uint8_t* base[PEER_MAX_WORLD];   // base[r]: rank r's symmetric region as mapped in THIS process (base[rank] = own)
  uint32_t* seq;                   // [PEER_SLOTS], local
  int* error;                      // local; set when a wait timed out
  int max_spins;                   // polls before a wait gives up (2^25 ~ 30 s; b2g_comm_set_timeout / B2G_PEER_TIMEOUT_S)
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_volatile_d2(const double2* p) {
  double2 v;
  asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// byte offset of slot `slot`'s slice of parity `par` inside a region
__device__ __forceinline__ size_t peer_slice_off(int slot, uint32_t par) {
  return PEER_DATA_OFF + ((size_t)par * PEER_SLOTS + slot) * PEER_SLICE_BYTES;
}

// Sequence number of the call about to run on `slot` (all threads of the CTA read the same value).
__device__ __forceinline__ uint32_t peer_next_seq(const PeerCtx& c, int slot) { return c.seq[slot] + 1u; }

// Cross-rank rendezvous of one CTA per rank on `slot`.  Call with ALL threads of the CTA, after the CTA's threads wrote
// their payload into the own slice; when it returns every rank's slice of this parity is complete and visible.
__device__ __forceinline__ void peer_signal_wait(const PeerCtx& c, int slot, uint32_t seq) {
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int r = threadIdx.x;
    __threadfence_system();                                   // the CTA's payload stores (ordered by the barrier) first
    st_volatile_u32(reinterpret_cast<uint32_t*>(c.base[r]) + (size_t)slot * PEER_MAX_WORLD + c.rank, seq);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(c.base[c.rank]) + (size_t)slot * PEER_MAX_WORLD + r;
    bool ok = false;
    const bool dead = *reinterpret_cast<volatile int*>(c.error) != 0;      // an earlier wait already timed out
    for (int it = 0; it < (dead ? 1 : c.max_spins); ++it) {
      if ((int32_t)(ld_volatile_u32(f) - seq) >= 0) {
        ok = true;
        break;
      }
      if (it > 4096) __nanosleep(it > (1 << 16) ? 1000 : 200);             // ~30 s in total
    }
    if (!ok) *c.error = 1;
    __threadfence_system();
  }
  __syncthreads();
}

// Record that the call `seq` on `slot` is done (one thread, after the CTA's last use of seq).
__device__ __forceinline__ void peer_commit_seq(const PeerCtx& c, int slot, uint32_t seq) { c.seq[slot] = seq; }
#endif

}  // namespace b2g

// Peer-memory communicator: CUDA-IPC setup of the symmetric regions and the one-shot all-reduce kernels (peer.cuh).
// Replaces the NCCL launches of the latency-bound exchanges of the patient-partitioned mode; NCCL (torch.distributed)
// stays the plumbing for setup-time integer all-reduces and anything larger than the slices hold.
#include <stdlib.h>
#include <string.h>
#include "peer.cuh"

struct b2g_comm {
  b2g::PeerCtx ctx;
  void* local_base;
  void* opened[b2g::PEER_MAX_WORLD];
  uint32_t* seq;
  int* error;
};

namespace {
using namespace b2g;

constexpr int AR_THREADS = 256;

// out[i] = sum over ranks of in_r[i]; one CTA per 16 KB slice; T4 = float4 or double2 (16-byte units)
template <typename T4>
__device__ __forceinline__ T4 ldv16(const T4* p);
template <>
__device__ __forceinline__ float4 ldv16<float4>(const float4* p) { return ld_volatile_f4(p); }
template <>
__device__ __forceinline__ double2 ldv16<double2>(const double2* p) { return ld_volatile_d2(p); }
__device__ __forceinline__ void add16(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void add16(double2& a, const double2& b) { a.x += b.x; a.y += b.y; }

template <typename T4>
__global__ void __launch_bounds__(AR_THREADS) k_peer_allreduce(PeerCtx c, const T4* __restrict__ in, T4* __restrict__ out, int64_t n16) {
  const int slot = blockIdx.x;
  const uint32_t seq = peer_next_seq(c, slot);
  const size_t off = peer_slice_off(slot, seq & 1u);
  constexpr int PER = PEER_SLICE_BYTES / 16;                 // 16-byte units per slice
  const int64_t i0 = (int64_t)slot * PER;
  const int cnt = (int)((n16 - i0) < PER ? (n16 - i0) : PER);
  T4* mine = reinterpret_cast<T4*>(c.base[c.rank] + off);
  for (int i = threadIdx.x; i < cnt; i += AR_THREADS) mine[i] = in[i0 + i];
  peer_signal_wait(c, slot, seq);
  for (int i = threadIdx.x; i < cnt; i += AR_THREADS) {
    T4 acc = ldv16<T4>(reinterpret_cast<const T4*>(c.base[0] + off) + i);
    for (int r = 1; r < c.world; ++r) add16(acc, ldv16<T4>(reinterpret_cast<const T4*>(c.base[r] + off) + i));
    out[i0 + i] = acc;
  }
  if (threadIdx.x == 0) peer_commit_seq(c, slot, seq);
}
}  // namespace

extern "C" int b2g_comm_set_timeout(b2g_comm* c, double seconds);
extern "C" size_t b2g_comm_region_bytes(void) { return b2g::PEER_REGION_BYTES; }
extern "C" size_t b2g_comm_max_bytes(void) { return (size_t)b2g::PEER_AR_SLOTS * b2g::PEER_SLICE_BYTES; }

/* SYNC.  Allocates and zeroes this rank's symmetric region and returns its CUDA-IPC handle (64 bytes) for the peers. */
extern "C" int b2g_comm_local_alloc(void** region, unsigned char* h_handle64) {
  B2G_CHECK_ARG(region && h_handle64, "comm_local_alloc: null pointer");
  void* p = nullptr;
  B2G_CUDA(cudaMalloc(&p, b2g::PEER_REGION_BYTES));
  B2G_CUDA(cudaMemset(p, 0, b2g::PEER_REGION_BYTES));
  B2G_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  B2G_CUDA(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(h_handle64, &h, 64);
  *region = p;
  return B2G_OK;
}

/* SYNC.  h_handles: world x 64 bytes, rank order (the own entry is ignored).  Call after EVERY rank has finished
 * b2g_comm_local_alloc (the handle exchange is that rendezvous). */
extern "C" int b2g_comm_create(int rank, int world, void* local_region, const unsigned char* h_handles, b2g_comm** out) {
  B2G_CHECK_ARG(out && local_region && h_handles && world >= 1 && world <= b2g::PEER_MAX_WORLD && rank >= 0 && rank < world,
                "comm_create: bad args (rank=%d world=%d, at most %d ranks of one node)", rank, world, b2g::PEER_MAX_WORLD);
  b2g_comm* c = new b2g_comm();
  memset(c, 0, sizeof(*c));
  c->local_base = local_region;
  c->ctx.rank = rank;
  c->ctx.world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->ctx.base[r] = (uint8_t*)local_region;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handles + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      b2g::set_error("comm_create: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
      (void)cudaGetLastError();                       // do not leave the error for the next launch check to trip over
      for (int q = 0; q < r; ++q)
        if (c->opened[q]) cudaIpcCloseMemHandle(c->opened[q]);
      delete c;
      return B2G_ECUDA;
    }
    c->opened[r] = p;
    c->ctx.base[r] = (uint8_t*)p;
  }
  B2G_CUDA(cudaMalloc((void**)&c->seq, b2g::PEER_SLOTS * sizeof(uint32_t) + sizeof(int)));
  B2G_CUDA(cudaMemset(c->seq, 0, b2g::PEER_SLOTS * sizeof(uint32_t) + sizeof(int)));
  B2G_CUDA(cudaDeviceSynchronize());
  c->error = (int*)(c->seq + b2g::PEER_SLOTS);
  c->ctx.seq = c->seq;
  c->ctx.error = c->error;
  c->ctx.max_spins = 1 << 25;
  if (const char* e = getenv("B2G_PEER_TIMEOUT_S")) b2g_comm_set_timeout(c, atof(e));
  *out = c;
  return B2G_OK;
}

/* Seconds a rendezvous may wait for a peer before it gives up and raises the error flag (default ~30). */
extern "C" int b2g_comm_set_timeout(b2g_comm* c, double seconds) {
  B2G_CHECK_ARG(c && seconds > 0, "comm_set_timeout: bad args");
  double spins = seconds * ((double)(1 << 25) / 30.0);
  if (spins < 8192) spins = 8192;
  if (spins > 2.0e9) spins = 2.0e9;
  c->ctx.max_spins = (int)spins;
  return B2G_OK;
}

extern "C" int b2g_comm_destroy(b2g_comm* c) {
  if (!c) return B2G_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->ctx.world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  if (c->seq) cudaFree(c->seq);
  if (c->local_base) cudaFree(c->local_base);
  delete c;
  return B2G_OK;
}

/* SYNC.  1 if any wait of this communicator has timed out since creation (results after that are undefined). */
extern "C" int b2g_comm_error(b2g_comm* c) {
  if (!c) return 0;
  int e = 0;
  if (cudaMemcpy(&e, c->error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  return e;
}

/* internal: the device-side context for kernels with a fused exchange (norm.cu) */
extern "C" const void* b2g_comm_ctx(const b2g_comm* c) { return c ? (const void*)&c->ctx : nullptr; }

static int comm_allreduce(b2g_comm* c, const void* in, void* out, int64_t n, int elem, void* stream_) {
  B2G_CHECK_ARG(c && (n == 0 || (in && out)) && n >= 0, "comm_allreduce: bad args");
  if (n == 0) return B2G_OK;
  const int64_t bytes = n * elem;
  B2G_CHECK_ARG(bytes % 16 == 0 && aligned16(in) && aligned16(out), "comm_allreduce: payload must be a multiple of 16 bytes, 16-byte aligned");
  B2G_CHECK_ARG((size_t)bytes <= b2g_comm_max_bytes(), "comm_allreduce: %lld bytes exceed the %zu-byte one-shot limit", (long long)bytes,
                b2g_comm_max_bytes());
  const int64_t n16 = bytes / 16;
  const unsigned grid = (unsigned)b2g::ceil_div(bytes, b2g::PEER_SLICE_BYTES);
  if (elem == 4)
    k_peer_allreduce<float4><<<grid, AR_THREADS, 0, (cudaStream_t)stream_>>>(c->ctx, (const float4*)in, (float4*)out, n16);
  else
    k_peer_allreduce<double2><<<grid, AR_THREADS, 0, (cudaStream_t)stream_>>>(c->ctx, (const double2*)in, (double2*)out, n16);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

/* out[i] = sum over the ranks of in[i] (in == out allowed); every rank calls with the same n, in the same order.
 * The sum runs in rank order on every rank: all ranks receive bit-identical results. */
extern "C" int b2g_comm_allreduce_f32(b2g_comm* c, const float* in, float* out, int64_t n, void* stream) {
  return comm_allreduce(c, in, out, n, 4, stream);
}
extern "C" int b2g_comm_allreduce_f64(b2g_comm* c, const double* in, double* out, int64_t n, void* stream) {
  return comm_allreduce(c, in, out, n, 8, stream);
}

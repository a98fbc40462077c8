// Multi-tensor Adam step: torch.optim.Adam(params, lr, weight_decay) as the reference builds it
// (train.py:255-260 -- betas (0.9, 0.999), eps 1e-8, L2 weight decay added to the gradient, no amsgrad) for ALL parameter
// tensors in ONE launch.  torch's foreach implementation issues ~10 multi-tensor launches plus scalar bookkeeping per step
// (~130 us of launch-bound work at 484k parameters); here the host passes a table of tensors and a table of 4096-element
// chunks, and every CTA updates one chunk with 128-bit accesses.  The arithmetic follows torch's operation order
// (lerp for exp_avg, mul + addcmul for exp_avg_sq, sqrt / bias_correction2_sqrt + eps, addcdiv) in fp32.
#include "common.cuh"

namespace {
using namespace b2g;

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = 4096;   // elements per CTA

__global__ void __launch_bounds__(ADAM_THREADS) k_adam_multi(const b2g_adam_tensor_t* __restrict__ tensors, const int2* __restrict__ chunks,
                                                             float step_size, float bc2_sqrt, float w1, float beta2, float w2, float eps,
                                                             float weight_decay) {
  const int2 ck = chunks[blockIdx.x];                    // (tensor id, first element)
  const b2g_adam_tensor_t t = tensors[ck.x];
  const int64_t end = (int64_t)ck.y + ADAM_CHUNK < t.numel ? (int64_t)ck.y + ADAM_CHUNK : t.numel;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) | reinterpret_cast<uintptr_t>(t.exp_avg) |
                     reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15u) == 0;
  auto upd = [&](float& p, float g, float& m, float& v) {
    g = fmaf(weight_decay, p, g);                        // grad.add(param, alpha=weight_decay)
    m = fmaf(g - m, w1, m);                              // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(g * g, w2, v * beta2);                      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
  };
  if (vec) {
    for (int64_t i = (int64_t)ck.y + threadIdx.x * 4; i < end; i += ADAM_THREADS * 4) {
      if (i + 3 < end) {
        float4 p = *reinterpret_cast<float4*>(t.param + i), g = *reinterpret_cast<const float4*>(t.grad + i);
        float4 m = *reinterpret_cast<float4*>(t.exp_avg + i), v = *reinterpret_cast<float4*>(t.exp_avg_sq + i);
        upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(t.param + i) = p;
        *reinterpret_cast<float4*>(t.exp_avg + i) = m;
        *reinterpret_cast<float4*>(t.exp_avg_sq + i) = v;
      } else {
        for (int64_t j = i; j < end; ++j) upd(t.param[j], t.grad[j], t.exp_avg[j], t.exp_avg_sq[j]);
      }
    }
  } else {
    for (int64_t i = (int64_t)ck.y + threadIdx.x; i < end; i += ADAM_THREADS) upd(t.param[i], t.grad[i], t.exp_avg[i], t.exp_avg_sq[i]);
  }
}
}  // namespace

extern "C" int b2g_adam_chunk_elems(void) { return ADAM_CHUNK; }

/* d_tensors: n_tensors descriptors, d_chunks: n_chunks (tensor id, first element) pairs -- both in DEVICE memory, built by
 * the caller once per set of parameter / gradient tensors.  One launch updates every parameter.  The scalars are computed
 * by the host in double precision exactly as torch.optim.Adam does for step t:
 *   step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t), one_minus_beta1 = 1 - beta1, one_minus_beta2 = 1 - beta2 */
extern "C" int b2g_adam_step(const b2g_adam_tensor_t* d_tensors, int n_tensors, const int32_t* d_chunks, int n_chunks, double step_size,
                             double bc2_sqrt, double one_minus_beta1, double beta2, double one_minus_beta2, double eps, double weight_decay,
                             void* stream_) {
  B2G_CHECK_ARG(n_tensors >= 0 && n_chunks >= 0 && (n_chunks == 0 || (d_tensors && d_chunks && n_tensors > 0)), "adam_step: bad args");
  if (n_chunks == 0) return B2G_OK;
  k_adam_multi<<<(unsigned)n_chunks, ADAM_THREADS, 0, (cudaStream_t)stream_>>>(
      d_tensors, reinterpret_cast<const int2*>(d_chunks), (float)step_size, (float)bc2_sqrt, (float)one_minus_beta1, (float)beta2,
      (float)one_minus_beta2, (float)eps, (float)weight_decay);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

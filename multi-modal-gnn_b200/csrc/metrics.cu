// On-device evaluation metrics: the step right after the hot path in the reference's
// evaluate.py -- per-lab +-3 sigma winsorisation of the residuals (evaluate.py:417-440) followed by MAE / RMSE / R^2 / MAPE
// overall and per lab (evaluate.py:36-82, 88-139).  The reference loops over labs in numpy on the host; here one CTA per lab
// walks that lab's pairs through a by-lab CSR (two passes: moments, then clipped sums), fp64 accumulation, block reductions in
// fixed order -> deterministic.  The host combines the per-lab records (a few hundred rows) into the overall figures.
#include "common.cuh"

namespace {
using namespace b2g;

constexpr int EV_THREADS = 256;
constexpr int EV_FIELDS = 10;   // n, mean_r, std_r, n_capped, sum|rc|, sum rc^2, sum t, sum t^2, sum |rc/t| (t != 0), n(t != 0)

__device__ __forceinline__ double block_sum(double v, double* sh) {   // all threads get the total; fixed order
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int w = 0; w < EV_THREADS / 32; ++w) t += sh[w];
  return t;
}

__global__ void __launch_bounds__(EV_THREADS) k_eval_per_lab(const float* __restrict__ pred, const float* __restrict__ target,
                                                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ pair_of,
                                                             int winsorize, float n_sigma, double* __restrict__ out,
                                                             float* __restrict__ pred_w) {
  __shared__ double sh[EV_THREADS / 32];
  const int lab = blockIdx.x;
  const int b = rowptr[lab], e = rowptr[lab + 1];
  const int n = e - b;
  double s1 = 0, s2 = 0;
  for (int j = b + threadIdx.x; j < e; j += EV_THREADS) {
    const int p = pair_of[j];
    const double r = (double)(pred[p] - target[p]);       // the reference's residual is a float32 difference
    s1 += r;
    s2 += r * r;
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  const double mean = n > 0 ? s1 / n : 0.0;
  double var = n > 0 ? s2 / n - mean * mean : 0.0;          // np.std: population standard deviation
  if (var < 0) var = 0;
  const double sd = sqrt(var);
  const bool clip = winsorize && n > 1;                      // evaluate.py:424 `if len(lab_residuals) > 1`
  const float lo = (float)(mean - (double)n_sigma * sd), hi = (float)(mean + (double)n_sigma * sd);
  double a_abs = 0, a_sq = 0, a_t = 0, a_t2 = 0, a_ape = 0, a_nz = 0, a_cap = 0;
  for (int j = b + threadIdx.x; j < e; j += EV_THREADS) {
    const int p = pair_of[j];
    const float t = target[p];
    float r = pred[p] - t;
    if (clip) {
      const float rc = fminf(fmaxf(r, lo), hi);
      if (rc != r) a_cap += 1.0;
      r = rc;
    }
    const float pw = t + r;                                  // predictions_np[mask] = targets + capped residuals (float32)
    if (pred_w) pred_w[p] = clip ? pw : pred[p];
    const double d = (double)(clip ? pw : pred[p]) - (double)t;
    a_abs += fabs(d);
    a_sq += d * d;
    a_t += (double)t;
    a_t2 += (double)t * (double)t;
    if (t != 0.f) {
      a_ape += fabs(d / (double)t);
      a_nz += 1.0;
    }
  }
  a_abs = block_sum(a_abs, sh); a_sq = block_sum(a_sq, sh); a_t = block_sum(a_t, sh); a_t2 = block_sum(a_t2, sh);
  a_ape = block_sum(a_ape, sh); a_nz = block_sum(a_nz, sh); a_cap = block_sum(a_cap, sh);
  if (threadIdx.x == 0) {
    double* o = out + (size_t)lab * EV_FIELDS;
    o[0] = n; o[1] = mean; o[2] = sd; o[3] = a_cap; o[4] = a_abs; o[5] = a_sq; o[6] = a_t; o[7] = a_t2; o[8] = a_ape; o[9] = a_nz;
  }
}
}  // namespace

extern "C" int b2g_eval_fields(void) { return EV_FIELDS; }

/* Per-lab evaluation record out[n_lab][10] (double): n, mean residual, residual std, number of winsorised residuals,
 * sum |r|, sum r^2, sum t, sum t^2, sum |r / t| over t != 0, count of t != 0 -- with r the (optionally winsorised) residual
 * pred - target of the lab's pairs.  rowptr / pair_of: CSR of the pair list keyed by lab (b2g_csr_build with val = pair id).
 * pred_w (optional, [M]): the winsorised predictions, as evaluate.py:434-437 writes them back. */
extern "C" int b2g_eval_per_lab(const float* pred, const float* target, const int32_t* rowptr, const int32_t* pair_of, int n_lab,
                                int winsorize, float n_sigma, double* out, float* pred_w, void* stream_) {
  B2G_CHECK_ARG(n_lab >= 0 && (n_lab == 0 || (pred && target && rowptr && pair_of && out)) && n_sigma >= 0.f, "eval_per_lab: bad args");
  if (n_lab == 0) return B2G_OK;
  k_eval_per_lab<<<(unsigned)n_lab, EV_THREADS, 0, (cudaStream_t)stream_>>>(pred, target, rowptr, pair_of, winsorize, n_sigma, out, pred_w);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

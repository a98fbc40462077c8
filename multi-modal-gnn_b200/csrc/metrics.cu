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

constexpr int EV_BINS = 3;      // patient-degree strata of evaluate.py:268-272: 1-5, 6-15, 16+ observed labs
constexpr int EV_BIN_FIELDS = 7;   // n, sum|r|, sum r^2, sum t, sum t^2, sum |r/t| (t != 0), n(t != 0)

__global__ void __launch_bounds__(EV_THREADS) k_eval_per_lab(const float* __restrict__ pred, const float* __restrict__ target,
                                                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ pair_of,
                                                             int winsorize, float n_sigma, double* __restrict__ out,
                                                             float* __restrict__ pred_w, const int64_t* __restrict__ patient_idx,
                                                             const int64_t* __restrict__ degree, double* __restrict__ out_bins) {
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
  double bins[EV_BINS][EV_BIN_FIELDS];
#pragma unroll
  for (int g = 0; g < EV_BINS; ++g)
#pragma unroll
    for (int f = 0; f < EV_BIN_FIELDS; ++f) bins[g][f] = 0.0;
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
    if (out_bins) {                                          // evaluate.py:237-287: strata of the pair's patient degree
      const int64_t dg = degree[patient_idx[p]];
      const int g = dg >= 16 ? 2 : (dg >= 6 ? 1 : (dg >= 1 ? 0 : -1));
#pragma unroll
      for (int q = 0; q < EV_BINS; ++q) {
        const double on = (q == g) ? 1.0 : 0.0;
        bins[q][0] += on; bins[q][1] += on * fabs(d); bins[q][2] += on * d * d; bins[q][3] += on * (double)t;
        bins[q][4] += on * (double)t * (double)t;
        if (t != 0.f) { bins[q][5] += on * fabs(d / (double)t); bins[q][6] += on; }
      }
    }
  }
  if (out_bins) {
#pragma unroll
    for (int g = 0; g < EV_BINS; ++g)
#pragma unroll
      for (int f = 0; f < EV_BIN_FIELDS; ++f) {
        const double v = block_sum(bins[g][f], sh);
        if (threadIdx.x == 0) out_bins[((size_t)lab * EV_BINS + g) * EV_BIN_FIELDS + f] = v;
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
  k_eval_per_lab<<<(unsigned)n_lab, EV_THREADS, 0, (cudaStream_t)stream_>>>(pred, target, rowptr, pair_of, winsorize, n_sigma, out, pred_w,
                                                                            nullptr, nullptr, nullptr);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

/* The same pass with the patient-degree strata of evaluate.py:237-287 (stratify_by_patient_degree): out_bins[n_lab][3][7]
 * (double) = per lab and per degree group {1-5, 6-15, 16+ observed labs of the pair's patient}: n, sum |r|, sum r^2, sum t,
 * sum t^2, sum |r / t| over t != 0, count of t != 0, on the winsorised residuals (the reference stratifies after the outlier
 * guard, evaluate.py:525-530).  degree[N_patient] is torch.bincount(has_lab.edge_index[0]) as int64 (b2g_csr_degrees).
 * The lab-frequency strata (evaluate.py:290-341) are unions of labs, i.e. sums of rows of `out`. */
extern "C" int b2g_eval_per_lab_strata(const float* pred, const float* target, const int32_t* rowptr, const int32_t* pair_of,
                                       const int64_t* patient_idx, const int64_t* degree, int n_lab, int winsorize, float n_sigma,
                                       double* out, double* out_bins, float* pred_w, void* stream_) {
  B2G_CHECK_ARG(n_lab >= 0 && (n_lab == 0 || (pred && target && rowptr && pair_of && out && patient_idx && degree && out_bins)) && n_sigma >= 0.f,
                "eval_per_lab_strata: bad args");
  if (n_lab == 0) return B2G_OK;
  k_eval_per_lab<<<(unsigned)n_lab, EV_THREADS, 0, (cudaStream_t)stream_>>>(pred, target, rowptr, pair_of, winsorize, n_sigma, out, pred_w,
                                                                            patient_idx, degree, out_bins);
  B2G_LAUNCH_CHECK();
  return B2G_OK;
}

// Evaluation statistics on the device: the autocorrelation score of timeGAN/evaluation.py.
//
// Reference: autocorr_seq (evaluation.py:63-71) called for every (window, channel) by statistical_similarity
// (evaluation.py:126-131):  0 if std(x) < 1e-8, else the MEAN over lag = 1..maxlag (lag < T) of the Pearson
// correlation of x[:-lag] and x[lag:] -- each slice with its OWN mean and variance (np.corrcoef), in float64.
// In numpy that is N*C*maxlag corrcoef calls (minutes for a few thousand windows); here one CTA owns one
// (window, channel) series in shared memory, its warps take the lags round-robin, every lane accumulates the five
// running sums in fp64 and the per-lag correlations are averaged in a fixed order.
#include "common.cuh"
#include "kernels.h"
#include "losses.h"

namespace {

constexpr int AC_THREADS = 128;

__global__ void __launch_bounds__(AC_THREADS) acf_score_kernel(const float* __restrict__ x, int T, int C, int maxlag,
                                                               double* __restrict__ out) {
  extern __shared__ float xs[];                 // [T]
  __shared__ double part[AC_THREADS / 32];
  __shared__ double stat[2];
  const int n = blockIdx.x, c = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* src = x + (size_t)n * T * C + c;
  double s = 0.0, ss = 0.0;
  for (int t = tid; t < T; t += AC_THREADS) {
    const float v = src[(size_t)t * C];
    xs[t] = v;
    s += v;
    ss += (double)v * v;
  }
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  if (lane == 0) part[warp] = s;
  __syncthreads();
  if (tid == 0) { double a = 0; for (int w = 0; w < AC_THREADS / 32; ++w) a += part[w]; stat[0] = a; }
  __syncthreads();
  if (lane == 0) part[warp] = ss;
  __syncthreads();
  if (tid == 0) { double a = 0; for (int w = 0; w < AC_THREADS / 32; ++w) a += part[w]; stat[1] = a; }
  __syncthreads();
  const double mean = stat[0] / T;
  const double var = stat[1] / T - mean * mean;          // np.std: population variance
  const int nlag = min(maxlag, T - 1);
  if (!(var >= 1e-16) || nlag <= 0) {                    // std < 1e-8 (or nothing to correlate): score 0
    if (tid == 0) out[(size_t)n * C + c] = 0.0;
    return;
  }
  double acc = 0.0;                                      // this warp's sum of correlations
  for (int lag = 1 + warp; lag <= nlag; lag += AC_THREADS / 32) {
    const int m = T - lag;
    double sa = 0, sb = 0, sab = 0, saa = 0, sbb = 0;
    for (int i = lane; i < m; i += 32) {
      const double a = xs[i], b = xs[i + lag];
      sa += a; sb += b; sab += a * b; saa += a * a; sbb += b * b;
    }
    sa = warp_sum_d(sa); sb = warp_sum_d(sb); sab = warp_sum_d(sab); saa = warp_sum_d(saa); sbb = warp_sum_d(sbb);
    const double cov = sab - sa * sb / m, va = saa - sa * sa / m, vb = sbb - sb * sb / m;
    acc += cov / sqrt(va * vb);                          // NaN for a constant slice, like np.corrcoef
  }
  __syncthreads();
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double a = 0;
    for (int w = 0; w < AC_THREADS / 32; ++w) a += part[w];
    out[(size_t)n * C + c] = a / nlag;
  }
}

}  // namespace

int tg_acf_score_impl(cudaStream_t st, const float* x, int N, int T, int C, int maxlag, double* out) {
  TG_REQUIRE(x && out, TG_ERR_ARG, "acf_score: null pointer");
  TG_REQUIRE(N > 0 && T > 0 && C > 0 && C <= 65535 && maxlag >= 0, TG_ERR_SHAPE, "acf_score: bad shape N=%d T=%d C=%d", N, T, C);
  const size_t smem = (size_t)T * sizeof(float);
  TG_REQUIRE(smem <= 48 * 1024, TG_ERR_UNSUPPORTED, "acf_score: window of %d samples does not fit shared memory", T);
  acf_score_kernel<<<dim3(N, C), AC_THREADS, smem, st>>>(x, T, C, maxlag, out);
  return tg_check_launch("acf_score");
}

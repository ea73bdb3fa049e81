// Fused loss kernels of the TimeGAN training step (north_star kernel family (4)); all HBM-bound.
//
// Reference call sites (timeGAN/train_timegan.py):
//   recon_loss           tt:72-74     10*sqrt(mean((x-x~)^2)+1e-8)        -> sqdiff_sum + scaled_diff
//   phase-2 sup MSE      tt:156-158   mean((S(h[:,:-1]) - h[:,1:])^2)      -> sqdiff_sum + scaled_diff
//   sup_loss_fake        tt:79-80     mean((h[:,1:]-h[:,:-1])^2)           -> diff1_sum + diff1_grad
//   batch_cov(_with_grad) tt:82-101   centred Gram / (BT-1)                -> center + wgrad GEMM (gemm_ffma.cu)
//   acf_loss_torch       tt:103-126   lagged products of z-scored series   -> acf_fwd / acf_bwd / acf_bwd_final
// Every reduction is two-stage with a fixed summation order (deterministic, no float atomics); under
// data parallelism the host all-reduces the small sufficient statistics between the stages.
#include "common.cuh"
#include "kernels.h"
#include "losses.h"

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_BLOCKS = 592;  // 4 x 148 SMs

__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double sm[RED_THREADS / 32];
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < RED_THREADS / 32) ? sm[threadIdx.x] : 0.0;
    t = warp_sum_d(t);
  }
  return t;  // valid in thread 0
}

// part[blk] = sum over the block's grid-stride slice of (a-b)^2 ; float4 path when aligned
__global__ void __launch_bounds__(RED_THREADS) sqdiff_partial_kernel(const float* __restrict__ a,
                                                                      const float* __restrict__ b, long long n,
                                                                      int vec, double* part) {
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n / 4;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (long long k = i; k < n4; k += stride) {
      float4 x = a4[k], y = b4[k];
      float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      s += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
    }
    for (long long k = n4 * 4 + i; k < n; k += stride) { float d = a[k] - b[k]; s += (double)(d * d); }
  } else {
    for (long long k = i; k < n; k += stride) { float d = a[k] - b[k]; s += (double)(d * d); }
  }
  s = block_sum_d(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(RED_THREADS) final_sum_kernel(const double* part, int nparts, float* out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += RED_THREADS) s += part[i];
  s = block_sum_d(s);
  if (threadIdx.x == 0) out[0] = (float)s;
}

// out (+)= coef[0] * (a - b)
__global__ void scaled_diff_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   const float* __restrict__ coef, float* __restrict__ out, long long n, int accumulate) {
  const float c = coef[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    float v = c * (a[k] - b[k]);
    out[k] = accumulate ? out[k] + v : v;
  }
}

// first-difference sum of squares over (B,T,H): sum_{t>=1} (h_t - h_{t-1})^2
__global__ void __launch_bounds__(RED_THREADS) diff1_partial_kernel(const float* __restrict__ h, int B, int T, int H,
                                                                     double* part) {
  double s = 0.0;
  const long long n = (long long)B * T * H, TH = (long long)T * H;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    if ((k % TH) >= H) { float d = h[k] - h[k - H]; s += (double)(d * d); }
  }
  s = block_sum_d(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// out (+)= coef * ( (h_t - h_{t-1})[t>=1] - (h_{t+1} - h_t)[t<T-1] )
__global__ void diff1_grad_kernel(const float* __restrict__ h, const float* __restrict__ coef, float* __restrict__ out,
                                  int B, int T, int H, int accumulate) {
  const float c = coef[0];
  const long long n = (long long)B * T * H, TH = (long long)T * H;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const long long r = k % TH;
    float g = 0.f;
    const float hv = h[k];
    if (r >= H) g += hv - h[k - H];
    if (r < TH - H) g -= h[k + H] - hv;
    g *= c;
    out[k] = accumulate ? out[k] + g : g;
  }
}

// out[r][c] = (x[r][c] - mean[c]) * scale[c]   (scale == nullptr -> 1)
__global__ void center_scale_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ scale, float* __restrict__ out, long long rows, int C) {
  const long long n = rows * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int c = (int)(k % C);
    float v = x[k] - mean[c];
    if (scale) v *= scale[c];
    out[k] = v;
  }
}

// One CTA per sequence.  xz: (B,T,C) z-scored.  part[b][l-1][c] = sum_{t<T-l} xz[b,t,c]*xz[b,t+l,c]
__global__ void __launch_bounds__(256) acf_fwd_kernel(const float* __restrict__ xz, int T, int C, int L, float* part) {
  extern __shared__ float sx[];  // [T][C]
  const int b = blockIdx.x;
  const float* src = xz + (size_t)b * T * C;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) sx[i] = src[i];
  __syncthreads();
  for (int p = threadIdx.x; p < L * C; p += blockDim.x) {
    const int l = p / C + 1, c = p % C;
    float s0 = 0.f, s1 = 0.f;
    int t = 0;
    for (; t + 1 < T - l; t += 2) {
      s0 = fmaf(sx[t * C + c], sx[(t + l) * C + c], s0);
      s1 = fmaf(sx[(t + 1) * C + c], sx[(t + 1 + l) * C + c], s1);
    }
    if (t < T - l) s0 = fmaf(sx[t * C + c], sx[(t + l) * C + c], s0);
    part[((size_t)b * L + (l - 1)) * C + c] = s0 + s1;
  }
}

// gz[b,t,c] = sum_l S[l-1][c] * (xz[b,t+l,c] + xz[b,t-l,c]);  stat[b][0][c] = sum_t gz, stat[b][1][c] = sum_t gz*xz
__global__ void __launch_bounds__(256) acf_bwd_kernel(const float* __restrict__ xz, const float* __restrict__ S, int T,
                                                       int C, int L, float* __restrict__ gz, float* __restrict__ stat) {
  extern __shared__ float sm[];
  float* sx = sm;            // [T][C]
  float* sS = sm + T * C;    // [L][C]
  float* red = sS + L * C;   // [2][C] accumulators (float atomics avoided: fixed-order reduction below)
  const int b = blockIdx.x;
  const float* src = xz + (size_t)b * T * C;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) sx[i] = src[i];
  for (int i = threadIdx.x; i < L * C; i += blockDim.x) sS[i] = S[i];
  __syncthreads();
  float* dst = gz + (size_t)b * T * C;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    float g = 0.f;
    for (int l = 1; l <= L; ++l) {
      float v = 0.f;
      if (t + l < T) v += sx[(t + l) * C + c];
      if (t - l >= 0) v += sx[(t - l) * C + c];
      g = fmaf(sS[(l - 1) * C + c], v, g);
    }
    dst[i] = g;
  }
  __syncthreads();
  // per-channel sums over t, fixed order: thread c walks its channel (T is small; C threads active)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    for (int t = 0; t < T; ++t) {
      float g = dst[t * C + c];  // written by this block above (visible after __syncthreads)
      s0 += g;
      s1 += (double)g * sx[t * C + c];
    }
    red[c] = (float)s0; red[C + c] = (float)s1;
    stat[((size_t)b * 2 + 0) * C + c] = (float)s0;
    stat[((size_t)b * 2 + 1) * C + c] = (float)s1;
  }
}

// dx (+)= (gz - mg[c]) * inv_s[c] - xz * kc[c]
__global__ void acf_bwd_final_kernel(const float* __restrict__ gz, const float* __restrict__ xz,
                                     const float* __restrict__ mg, const float* __restrict__ kc,
                                     const float* __restrict__ inv_s, float* __restrict__ dx, long long rows, int C,
                                     int accumulate) {
  const long long n = rows * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const int c = (int)(k % C);
    float v = (gz[k] - mg[c]) * inv_s[c] - xz[k] * kc[c];
    dx[k] = accumulate ? dx[k] + v : v;
  }
}

int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)tg_num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

size_t tg_reduce_ws_bytes() { return RED_BLOCKS * sizeof(double); }

int tg_sqdiff_sum_impl(cudaStream_t st, const float* a, const float* b, long long n, float* out, void* ws, size_t wsb) {
  TG_REQUIRE(a && b && out && ws, TG_ERR_ARG, "sqdiff_sum: null pointer");
  TG_REQUIRE(n > 0, TG_ERR_SHAPE, "sqdiff_sum: n=%lld", n);
  TG_REQUIRE(wsb >= tg_reduce_ws_bytes(), TG_ERR_ARG, "sqdiff_sum: workspace too small");
  int blocks = (int)((n / 4 + RED_THREADS - 1) / RED_THREADS);
  if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
  if (blocks < 1) blocks = 1;
  const int vec = tg_aligned16(a) && tg_aligned16(b);
  sqdiff_partial_kernel<<<blocks, RED_THREADS, 0, st>>>(a, b, n, vec, (double*)ws);
  int rc = tg_check_launch("sqdiff_partial");
  if (rc) return rc;
  final_sum_kernel<<<1, RED_THREADS, 0, st>>>((const double*)ws, blocks, out);
  return tg_check_launch("final_sum");
}

int tg_scaled_diff_impl(cudaStream_t st, const float* a, const float* b, const float* coef, float* out, long long n,
                        int accumulate) {
  TG_REQUIRE(a && b && coef && out, TG_ERR_ARG, "scaled_diff: null pointer");
  TG_REQUIRE(n > 0, TG_ERR_SHAPE, "scaled_diff: n=%lld", n);
  scaled_diff_kernel<<<ew_blocks(n), 256, 0, st>>>(a, b, coef, out, n, accumulate);
  return tg_check_launch("scaled_diff");
}

int tg_diff1_sum_impl(cudaStream_t st, const float* h, int B, int T, int H, float* out, void* ws, size_t wsb) {
  TG_REQUIRE(h && out && ws, TG_ERR_ARG, "diff1_sum: null pointer");
  TG_REQUIRE(B > 0 && T > 1 && H > 0, TG_ERR_SHAPE, "diff1_sum: bad shape");
  TG_REQUIRE(wsb >= tg_reduce_ws_bytes(), TG_ERR_ARG, "diff1_sum: workspace too small");
  long long n = (long long)B * T * H;
  int blocks = (int)((n + RED_THREADS - 1) / RED_THREADS);
  if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
  diff1_partial_kernel<<<blocks, RED_THREADS, 0, st>>>(h, B, T, H, (double*)ws);
  int rc = tg_check_launch("diff1_partial");
  if (rc) return rc;
  final_sum_kernel<<<1, RED_THREADS, 0, st>>>((const double*)ws, blocks, out);
  return tg_check_launch("final_sum");
}

int tg_diff1_grad_impl(cudaStream_t st, const float* h, const float* coef, float* out, int B, int T, int H,
                       int accumulate) {
  TG_REQUIRE(h && coef && out, TG_ERR_ARG, "diff1_grad: null pointer");
  TG_REQUIRE(B > 0 && T > 1 && H > 0, TG_ERR_SHAPE, "diff1_grad: bad shape");
  diff1_grad_kernel<<<ew_blocks((long long)B * T * H), 256, 0, st>>>(h, coef, out, B, T, H, accumulate);
  return tg_check_launch("diff1_grad");
}

int tg_center_scale_impl(cudaStream_t st, const float* x, const float* mean, const float* scale, float* out,
                         long long rows, int C) {
  TG_REQUIRE(x && mean && out, TG_ERR_ARG, "center_scale: null pointer");
  TG_REQUIRE(rows > 0 && C > 0, TG_ERR_SHAPE, "center_scale: bad shape");
  center_scale_kernel<<<ew_blocks(rows * C), 256, 0, st>>>(x, mean, scale, out, rows, C);
  return tg_check_launch("center_scale");
}

int tg_acf_fwd_impl(cudaStream_t st, const float* xz, int B, int T, int C, int L, float* part) {
  TG_REQUIRE(xz && part, TG_ERR_ARG, "acf_fwd: null pointer");
  TG_REQUIRE(B > 0 && T > 1 && C > 0 && L >= 1 && L < T, TG_ERR_SHAPE, "acf_fwd: bad shape B=%d T=%d C=%d L=%d", B, T, C, L);
  size_t smem = (size_t)T * C * 4;
  TG_REQUIRE(smem <= 200 * 1024, TG_ERR_UNSUPPORTED, "acf_fwd: T*C=%d too large for one CTA", T * C);
  TG_OPT_IN_SMEM(acf_fwd_kernel, "acf_fwd");
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("acf_fwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  acf_fwd_kernel<<<B, 256, smem, st>>>(xz, T, C, L, part);
  return tg_check_launch("acf_fwd");
}

int tg_acf_bwd_impl(cudaStream_t st, const float* xz, const float* S, int B, int T, int C, int L, float* gz,
                    float* stat) {
  TG_REQUIRE(xz && S && gz && stat, TG_ERR_ARG, "acf_bwd: null pointer");
  TG_REQUIRE(B > 0 && T > 1 && C > 0 && L >= 1 && L < T, TG_ERR_SHAPE, "acf_bwd: bad shape");
  size_t smem = ((size_t)T * C + (size_t)L * C + 2 * C) * 4;
  TG_REQUIRE(smem <= 200 * 1024, TG_ERR_UNSUPPORTED, "acf_bwd: T*C=%d too large for one CTA", T * C);
  TG_OPT_IN_SMEM(acf_bwd_kernel, "acf_bwd");
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("acf_bwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  acf_bwd_kernel<<<B, 256, smem, st>>>(xz, S, T, C, L, gz, stat);
  return tg_check_launch("acf_bwd");
}

int tg_acf_bwd_final_impl(cudaStream_t st, const float* gz, const float* xz, const float* mg, const float* kc,
                          const float* inv_s, float* dx, long long rows, int C, int accumulate) {
  TG_REQUIRE(gz && xz && mg && kc && inv_s && dx, TG_ERR_ARG, "acf_bwd_final: null pointer");
  TG_REQUIRE(rows > 0 && C > 0, TG_ERR_SHAPE, "acf_bwd_final: bad shape");
  acf_bwd_final_kernel<<<ew_blocks(rows * C), 256, 0, st>>>(gz, xz, mg, kc, inv_s, dx, rows, C, accumulate);
  return tg_check_launch("acf_bwd_final");
}

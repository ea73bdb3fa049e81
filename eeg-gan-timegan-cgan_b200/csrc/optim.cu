// Fused gradient-clip + Adam (north_star kernel family (4)).
//
// Reference: nn.utils.clip_grad_norm_(params, clip) followed by optim.Adam.step()
// (timeGAN/train_timegan.py:141-142,160-161,220-221,268-272; Adam(betas=(0.5,0.9), eps=1e-8), no weight
// decay, no amsgrad -- SURVEY.md Appendix A.3):
//     g <- g * min(1, c / (||g||_2 + 1e-6))            (norm over ALL tensors of the list)
//     m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2
//     p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// Two launches per optimiser step, both multi-tensor (pointer tables travel as kernel arguments):
//   sumsq_multi : fixed-order block partials of sum g^2  -> one scalar (deterministic)
//   adam_multi  : every thread derives the clip coefficient from that scalar, then updates p,m,v in place.
#include "common.cuh"
#include "kernels.h"
#include "losses.h"

namespace {

struct MtArgs {
  float* p[TG_MT_MAX];
  const float* g[TG_MT_MAX];
  float* m[TG_MT_MAX];
  float* v[TG_MT_MAX];
  long long size[TG_MT_MAX];
  int blk_start[TG_MT_MAX + 1];  // first block of tensor i; blk_start[n] = total blocks
  int n;
};

constexpr int MT_THREADS = 256;
constexpr int MT_CHUNK = 4096;  // elements per block

__device__ __forceinline__ int find_tensor(const MtArgs& a, int blk) {
  int t = 0;
  while (t + 1 < a.n && blk >= a.blk_start[t + 1]) ++t;
  return t;
}

__global__ void __launch_bounds__(MT_THREADS) sumsq_multi_kernel(const __grid_constant__ MtArgs a, double* part) {
  const int t = find_tensor(a, blockIdx.x);
  const long long base = (long long)(blockIdx.x - a.blk_start[t]) * MT_CHUNK;
  const long long end = min(a.size[t], base + MT_CHUNK);
  const float* g = a.g[t];
  double s = 0.0;
  for (long long i = base + threadIdx.x; i < end; i += MT_THREADS) { float x = g[i]; s += (double)x * x; }
  __shared__ double sm[MT_THREADS / 32];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < MT_THREADS / 32 ? sm[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) sumsq_final_kernel(const double* part, int n, float* out, int accumulate) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  __shared__ double sm[8];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sm[i];
    out[0] = accumulate ? out[0] + (float)t : (float)t;
  }
}

// Device-resident optimiser clock for CUDA-graph replay: state = {lr, step}; one thread advances the step and
// derives the bias-correction factors in double, like the host path does.
__global__ void adam_tick_kernel(float* __restrict__ state, float* __restrict__ derived, float beta1, float beta2) {
  const float step = state[1] + 1.0f;
  state[1] = step;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  derived[0] = (float)((double)state[0] / bc1);   // step_size
  derived[1] = (float)sqrt(bc2);                  // bc2_sqrt
}

__global__ void __launch_bounds__(MT_THREADS) adam_multi_kernel(const __grid_constant__ MtArgs a,
                                                                const float* __restrict__ sumsq, float max_norm,
                                                                float step_size, float bc2_sqrt, float beta1,
                                                                float beta2, float eps, float grad_scale,
                                                                const float* __restrict__ derived) {
  if (derived) { step_size = derived[0]; bc2_sqrt = derived[1]; }
  const int t = find_tensor(a, blockIdx.x);
  const long long base = (long long)(blockIdx.x - a.blk_start[t]) * MT_CHUNK;
  const long long end = min(a.size[t], base + MT_CHUNK);
  float clip = grad_scale;
  if (max_norm > 0.f) {
    const float norm = sqrtf(sumsq[0]) * grad_scale;
    const float c = max_norm / (norm + 1e-6f);
    clip = grad_scale * fminf(c, 1.0f);
  }
  float* p = a.p[t];
  const float* g = a.g[t];
  float* m = a.m[t];
  float* v = a.v[t];
  for (long long i = base + threadIdx.x; i < end; i += MT_THREADS) {
    const float gi = g[i] * clip;
    // exp_avg.lerp_(g, 1-b1) exactly as ATen evaluates it; exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    const float w1 = 1.f - beta1, mo = m[i];
    const float mi = (w1 < 0.5f) ? mo + w1 * (gi - mo) : gi - (gi - mo) * (1.f - w1);
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// Best-checkpoint snapshot without a host round trip (train_timegan.py:410-413: "if g_total < best: save"):
// every block reads the candidate loss and the best-so-far from device memory and copies its chunk of the
// weights / optimiser state into the snapshot only when the step improved; a one-thread kernel launched after
// it commits the new best and the step number.  m[] / v[] of MtArgs are unused here.
__global__ void __launch_bounds__(MT_THREADS) snapshot_if_better_kernel(const __grid_constant__ MtArgs a,
                                                                        const float* __restrict__ value,
                                                                        const float* __restrict__ best) {
  if (!(value[0] < best[0])) return;
  const int t = find_tensor(a, blockIdx.x);
  const long long base = (long long)(blockIdx.x - a.blk_start[t]) * MT_CHUNK;
  const long long end = min(a.size[t], base + MT_CHUNK);
  float* dst = a.p[t];
  const float* src = a.g[t];
  for (long long i = base + threadIdx.x; i < end; i += MT_THREADS) dst[i] = src[i];
}

__global__ void snapshot_commit_kernel(const float* __restrict__ value, float* __restrict__ best,
                                       float* __restrict__ best_step, float step) {
  if (value[0] < best[0]) { best[0] = value[0]; best_step[0] = step; }
}

int fill_blocks(MtArgs& a, const long long* sizes, int n) {
  int blk = 0;
  for (int i = 0; i < n; ++i) {
    a.size[i] = sizes[i];
    a.blk_start[i] = blk;
    blk += tg_ceil_div(sizes[i], MT_CHUNK);
  }
  a.blk_start[n] = blk;
  a.n = n;
  return blk;
}

}  // namespace

size_t tg_sumsq_ws_bytes(int n, const long long* sizes) {
  long long blk = 0;
  for (int i = 0; i < n; ++i) blk += tg_ceil_div(sizes[i], MT_CHUNK);
  return (size_t)blk * sizeof(double);
}

int tg_sumsq_multi_impl(cudaStream_t st, int n, const float* const* grads, const long long* sizes, float* out_sumsq,
                        void* ws, size_t wsb) {
  TG_REQUIRE(n > 0 && grads && sizes && out_sumsq && ws, TG_ERR_ARG, "sumsq_multi: bad arguments");
  size_t ws_off = 0;
  for (int i0 = 0; i0 < n; i0 += TG_MT_MAX) {
    const int cnt = (n - i0 < TG_MT_MAX) ? n - i0 : TG_MT_MAX;
    MtArgs a{};
    for (int i = 0; i < cnt; ++i) {
      TG_REQUIRE(grads[i0 + i] && sizes[i0 + i] > 0, TG_ERR_ARG, "sumsq_multi: tensor %d null/empty", i0 + i);
      a.g[i] = grads[i0 + i];
    }
    const int blocks = fill_blocks(a, sizes + i0, cnt);
    TG_REQUIRE(wsb >= (ws_off + blocks) * sizeof(double), TG_ERR_ARG, "sumsq_multi: workspace too small");
    double* part = (double*)ws + ws_off;
    sumsq_multi_kernel<<<blocks, MT_THREADS, 0, st>>>(a, part);
    int rc = tg_check_launch("sumsq_multi");
    if (rc) return rc;
    sumsq_final_kernel<<<1, 256, 0, st>>>(part, blocks, out_sumsq, i0 > 0);
    rc = tg_check_launch("sumsq_final");
    if (rc) return rc;
    ws_off += blocks;
  }
  return TG_OK;
}

int tg_adam_multi_impl(cudaStream_t st, int n, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const long long* sizes, const float* sumsq, float max_norm, float lr,
                       float beta1, float beta2, float eps, int step, float grad_scale, float* dev_state) {
  TG_REQUIRE(n > 0 && params && grads && exp_avg && exp_avg_sq && sizes, TG_ERR_ARG, "adam_multi: bad arguments");
  TG_REQUIRE(dev_state || step >= 1, TG_ERR_ARG, "adam_multi: step must be >= 1 (got %d)", step);
  TG_REQUIRE(max_norm <= 0.f || sumsq, TG_ERR_ARG, "adam_multi: clipping requested without sumsq");
  float step_size = 0.f, bc2_sqrt = 1.f;
  const float* derived = nullptr;
  if (dev_state) {
    // dev_state: float[4] = {lr, step, step_size, bc2_sqrt}; lr is written by the host outside any graph
    adam_tick_kernel<<<1, 1, 0, st>>>(dev_state, dev_state + 2, beta1, beta2);
    int rc = tg_check_launch("adam_tick");
    if (rc) return rc;
    derived = dev_state + 2;
  } else {
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    step_size = (float)((double)lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
  }
  for (int i0 = 0; i0 < n; i0 += TG_MT_MAX) {
    const int cnt = (n - i0 < TG_MT_MAX) ? n - i0 : TG_MT_MAX;
    MtArgs a{};
    for (int i = 0; i < cnt; ++i) {
      TG_REQUIRE(params[i0 + i] && grads[i0 + i] && exp_avg[i0 + i] && exp_avg_sq[i0 + i] && sizes[i0 + i] > 0,
                 TG_ERR_ARG, "adam_multi: tensor %d null/empty", i0 + i);
      a.p[i] = params[i0 + i]; a.g[i] = grads[i0 + i]; a.m[i] = exp_avg[i0 + i]; a.v[i] = exp_avg_sq[i0 + i];
    }
    const int blocks = fill_blocks(a, sizes + i0, cnt);
    adam_multi_kernel<<<blocks, MT_THREADS, 0, st>>>(a, sumsq, max_norm, step_size, bc2_sqrt, beta1, beta2, eps,
                                                     grad_scale, derived);
    int rc = tg_check_launch("adam_multi");
    if (rc) return rc;
  }
  return TG_OK;
}

int tg_snapshot_if_better_impl(cudaStream_t st, int n, float* const* dst, const float* const* src,
                               const long long* sizes, const float* value, float* best, float* best_step, float step) {
  TG_REQUIRE(n >= 0 && value && best && best_step && (n == 0 || (dst && src && sizes)), TG_ERR_ARG,
             "snapshot_if_better: bad arguments");
  for (int i0 = 0; i0 < n; i0 += TG_MT_MAX) {
    const int cnt = (n - i0 < TG_MT_MAX) ? n - i0 : TG_MT_MAX;
    MtArgs a{};
    for (int i = 0; i < cnt; ++i) {
      TG_REQUIRE(dst[i0 + i] && src[i0 + i] && sizes[i0 + i] > 0, TG_ERR_ARG, "snapshot_if_better: tensor %d null/empty",
                 i0 + i);
      a.p[i] = dst[i0 + i]; a.g[i] = src[i0 + i];
    }
    const int blocks = fill_blocks(a, sizes + i0, cnt);
    snapshot_if_better_kernel<<<blocks, MT_THREADS, 0, st>>>(a, value, best);
    int rc = tg_check_launch("snapshot_if_better");
    if (rc) return rc;
  }
  snapshot_commit_kernel<<<1, 1, 0, st>>>(value, best, best_step, step);
  return tg_check_launch("snapshot_commit");
}

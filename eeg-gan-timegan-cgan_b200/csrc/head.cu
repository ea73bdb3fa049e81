// Fused discriminator head: spectral-norm power iteration + Linear(H -> 1) + sigmoid + BCE + accuracy + soft throttle
// + the R1 seed, forward AND hand-derived backward (north_star kernel family (4): "BCE adversarial losses ... fused").
//
// Replaces, per optimiser step, ~150 (disc_step) / ~40 (gen_step) one-block ATen launches:
//   timegan_model.py:92-98   U.spectral_norm(nn.Linear(hidden, 1)) -> sigmoid(fc(y[:, -1]))     (legacy hook: one power
//                            iteration per forward call in train mode, eps 1e-12, u and v updated in place)
//   train_timegan.py:70      bce = nn.BCELoss()  (log terms clamped at -100, backward divides by max(p(1-p), 1e-12))
//   train_timegan.py:196     loss = 0.5 * (bce(d_real, y_real) + bce(d_fake, y_fake))
//   train_timegan.py:199-202 R1: grad of d_real.sum() w.r.t. the D input (this file: its seed dL/dy_last) and the
//                            term 0.5*gamma*mean_b ||grad_b||^2, differentiated as (gamma/B) * sdot
//   train_timegan.py:205-215 balanced accuracy and the throttle scale = max(0.2, 1 - max(0, acc - target)/band)
//   train_timegan.py:241     g_adv = bce(d_fake, ones)  (gen_step, D frozen)
//
// Everything here is a few hundred KFLOP on (B, H) <= (4096, 1024): ONE CTA per kernel, fixed summation order
// (bit-reproducible), no atomics.  Under data parallelism the four batch sums of `head_fwd` are all-reduced by the
// host between the kernels (dist.allreduce_stats), so accuracy, scale and the BCE mean are those of the global batch.
#include "common.cuh"
#include "kernels.h"
#include "losses.h"

namespace {

constexpr int HD_THREADS = 256;
constexpr int HD_WARPS = HD_THREADS / 32;

// sum over the CTA, same value returned to every thread; fixed order
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HD_WARPS; ++i) s += red[i];
  return s;
}

// One power iteration of the legacy spectral-norm hook for a (1 x H) weight (torch/nn/utils/spectral_norm.py):
//   v <- normalize(W^T u) = w*u / max(||w*u||, eps);  u <- normalize(W v) = (w.v) / max(|w.v|, eps);  sigma = u * (w.v)
// ws: w in shared memory; vs: v in shared memory (in/out); *u in/out.  Returns sigma.
__device__ float power_iteration(const float* ws, float* vs, float* u, int H, bool iterate, float* red) {
  const float eps = 1e-12f;
  if (iterate) {
    float part = 0.f;
    for (int k = threadIdx.x; k < H; k += HD_THREADS) { const float t = ws[k] * (*u); part += t * t; }
    const float nrm = sqrtf(block_sum(part, red));
    const float inv = 1.f / fmaxf(nrm, eps);
    const float uu = *u;
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += HD_THREADS) vs[k] = ws[k] * uu * inv;
    __syncthreads();
  }
  float part = 0.f;
  for (int k = threadIdx.x; k < H; k += HD_THREADS) part += ws[k] * vs[k];
  const float wv = block_sum(part, red);
  if (iterate) {
    __syncthreads();
    if (threadIdx.x == 0) *u = wv / fmaxf(fabsf(wv), eps);
    __syncthreads();
  }
  return (*u) * wv;
}

struct HeadFwd {
  const float* yl;      // (n_half*B, H) rows ld floats apart: last hidden state of every sequence
  long long ld;
  const float* w;       // (H)   fc.weight_orig
  const float* bias;    // (1)
  float* u;             // (1)   fc.weight_u   (in/out when training)
  float* v;             // (H)   fc.weight_v   (in/out when training)
  const float* labels;  // (n_half*B) targets, or NULL = all ones (gen_step)
  float* wbar;          // (n_half, H)  out: w / sigma of each call
  float* uv;            // (n_half, 1+H) out: the (u, v) each sigma was formed with (constants of the backward)
  float* sigma;         // (n_half) out
  float* p;             // (n_half*B) out: probabilities
  float* stats;         // (4) out: sum bce(half 0), sum bce(half 1), #(p0 > 0.5), #(p1 < 0.5)   (local sums)
  int B, H, n_half, training;
};

__global__ void __launch_bounds__(HD_THREADS) head_fwd_kernel(HeadFwd a) {
  extern __shared__ float sm[];
  float* ws = sm;                   // [H]
  float* vs = ws + a.H;             // [H]
  float* wb = vs + a.H;             // [n_half][H]
  float* red = wb + a.n_half * a.H; // [HD_WARPS]
  __shared__ float u_s;
  const int H = a.H;
  for (int k = threadIdx.x; k < H; k += HD_THREADS) { ws[k] = a.w[k]; vs[k] = a.v[k]; }
  if (threadIdx.x == 0) u_s = a.u[0];
  __syncthreads();
  for (int h = 0; h < a.n_half; ++h) {        // one forward call of D per half: real first, then fake (tt:192-193)
    const float sigma = power_iteration(ws, vs, &u_s, H, a.training != 0, red);
    const float inv = 1.f / sigma;
    for (int k = threadIdx.x; k < H; k += HD_THREADS) {
      wb[h * H + k] = ws[k] * inv;
      a.wbar[h * H + k] = ws[k] * inv;
      a.uv[h * (1 + H) + 1 + k] = vs[k];
    }
    if (threadIdx.x == 0) { a.sigma[h] = sigma; a.uv[h * (1 + H)] = u_s; }
    __syncthreads();
  }
  if (a.training) {
    for (int k = threadIdx.x; k < H; k += HD_THREADS) a.v[k] = vs[k];
    if (threadIdx.x == 0) a.u[0] = u_s;
  }
  // probabilities + batch sums: warp per row, lanes over H
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float b0 = a.bias ? a.bias[0] : 0.f;
  float s_bce[2] = {0.f, 0.f}, s_acc[2] = {0.f, 0.f};
  for (int h = 0; h < a.n_half; ++h) {
    for (int b = warp; b < a.B; b += HD_WARPS) {
      const float* row = a.yl + (long long)(h * a.B + b) * a.ld;
      float z = 0.f;
      for (int k = lane; k < H; k += 32) z += row[k] * wb[h * H + k];
      z = warp_sum(z) + b0;
      const float p = sigmoid_acc(z);
      if (lane == 0) {
        a.p[h * a.B + b] = p;
        const float y = a.labels ? a.labels[h * a.B + b] : 1.f;
        const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
        s_bce[h] += -(y * lp + (1.f - y) * l1p);
        s_acc[h] += (h == 0) ? (p > 0.5f ? 1.f : 0.f) : (p < 0.5f ? 1.f : 0.f);
      }
    }
  }
  __shared__ float part[HD_WARPS][4];
  if (lane == 0) { part[warp][0] = s_bce[0]; part[warp][1] = s_bce[1]; part[warp][2] = s_acc[0]; part[warp][3] = s_acc[1]; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < HD_WARPS; ++w) s += part[w][threadIdx.x];
    a.stats[threadIdx.x] = s;
  }
}

// dz of one sample for BCE through the sigmoid: ATen's binary_cross_entropy_backward divides by max(p(1-p), 1e-12),
// sigmoid's backward multiplies by p(1-p)
__device__ __forceinline__ float bce_dz(float p, float y) {
  const float pq = p * (1.f - p);
  return (p - y) * (pq / fmaxf(pq, 1e-12f));
}

struct HeadSeed {
  const float* p;       // (2B) from head_fwd
  const float* labels;  // (2B)
  const float* wbar;    // (2,H)
  const float* stats;   // (4) GLOBAL sums
  float* scal;          // (4) out: loss_bce, acc, scale, (unused)
  float* seed;          // (B,H) out: d d_real.sum() / d y_last(real) = p(1-p) wbar_r      (R1, tt:200)  -- may be NULL
  float* gyf;           // (B,H) out: d (scale * loss) / d y_last(fake)
  int B, H;
  float Bg, target, band;
};

__global__ void __launch_bounds__(HD_THREADS) head_seed_kernel(HeadSeed a) {
  const float loss = 0.5f * (a.stats[0] + a.stats[1]) / a.Bg;
  const float acc = 0.5f * (a.stats[2] / a.Bg + a.stats[3] / a.Bg);
  float scale = 1.f;
  if (a.band > 0.f) scale = fmaxf(0.2f, 1.f - fmaxf(0.f, acc - a.target) / a.band);
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.scal[0] = loss; a.scal[1] = acc; a.scal[2] = scale; a.scal[3] = 0.f; }
  const int n = a.B * a.H;
  for (int i = blockIdx.x * HD_THREADS + threadIdx.x; i < n; i += gridDim.x * HD_THREADS) {
    const int b = i / a.H, k = i - b * a.H;
    if (a.seed) { const float p = a.p[b]; a.seed[i] = p * (1.f - p) * a.wbar[k]; }
    const float pf = a.p[a.B + b];
    a.gyf[i] = scale * 0.5f / a.Bg * bce_dz(pf, a.labels[a.B + b]) * a.wbar[a.H + k];
  }
}

struct HeadBwd {
  const float* yl;      // (2B,H) rows ld apart
  long long ld;
  const float* hd;      // (B,H) rows ld_hd apart: tangent of y_last(real) (R1), or NULL
  long long ld_hd;
  const float* p;       // (2B)
  const float* labels;  // (2B)
  const float* w;       // (H)
  const float* wbar;    // (2,H)
  const float* uv;      // (2,1+H)
  const float* sigma;   // (2)
  const float* scal;    // loss_bce, acc, scale
  const float* r1;      // (1) GLOBAL mean_b ||grad_b||^2, or NULL
  float* gyr;           // (B,H) out
  float* ghd;           // (B,H) out (NULL without R1)
  float* gw;            // (H)   out: gradient of fc.weight_orig
  float* gb;            // (1)   out
  float* loss_val;      // (1)   out: (loss_bce + 0.5*gamma*r1) * scale   (the value disc_step returns, tt:225)
  int B, H;
  float Bg, gamma;
};

// one CTA; the weight gradient needs sums over the batch for every k: thread k walks the rows (coalesced across k)
__global__ void __launch_bounds__(HD_THREADS) head_bwd_kernel(HeadBwd a) {
  extern __shared__ float sm[];
  const int B = a.B, H = a.H;
  float* cz = sm;            // [B]  dL/dz of the real rows (BCE + R1 part)
  float* cf = cz + B;        // [B]  dL/dz of the fake rows
  float* ch = cf + B;        // [B]  coefficient of hd_b in g_wbar_r:  scale*gamma/Bg * p(1-p)
  float* gwr = ch + B;       // [H]  g_wbar (real call)
  float* gwf = gwr + H;      // [H]  g_wbar (fake call)
  float* red = gwf + H;      // [HD_WARPS]
  const float scale = a.scal[2];
  const float cb = scale * 0.5f / a.Bg, cr = a.hd ? scale * a.gamma / a.Bg : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-row scalars: s_b = hd_b . wbar_r
  for (int b = warp; b < B; b += HD_WARPS) {
    float s = 0.f;
    if (a.hd) {
      const float* row = a.hd + (long long)b * a.ld_hd;
      for (int k = lane; k < H; k += 32) s += row[k] * a.wbar[k];
      s = warp_sum(s);
    }
    if (lane == 0) {
      const float p = a.p[b], pq = p * (1.f - p);
      cz[b] = cb * bce_dz(p, a.labels[b]) + cr * pq * (1.f - 2.f * p) * s;
      ch[b] = cr * pq;
      cf[b] = cb * bce_dz(a.p[B + b], a.labels[B + b]);
    }
  }
  __syncthreads();
  // input gradients
  for (int i = threadIdx.x; i < B * H; i += HD_THREADS) {
    const int b = i / H, k = i - b * H;
    a.gyr[i] = cz[b] * a.wbar[k];
    if (a.ghd) a.ghd[i] = ch[b] * a.wbar[k];
  }
  // g_wbar of both calls and the bias gradient
  for (int k = threadIdx.x; k < H; k += HD_THREADS) {
    float sr = 0.f, sf = 0.f;
    for (int b = 0; b < B; ++b) {
      sr += cz[b] * a.yl[(long long)b * a.ld + k];
      if (a.hd) sr += ch[b] * a.hd[(long long)b * a.ld_hd + k];
      sf += cf[b] * a.yl[(long long)(B + b) * a.ld + k];
    }
    gwr[k] = sr; gwf[k] = sf;
  }
  float pb = 0.f;
  for (int b = threadIdx.x; b < B; b += HD_THREADS) pb += cz[b] + cf[b];
  const float gbias = block_sum(pb, red);
  // through wbar = w / sigma, sigma = u (w . v) with (u, v) constants:  g_w = g_wbar/sigma - (g_wbar . w)/sigma^2 * u v
  float dr = 0.f, df = 0.f;
  for (int k = threadIdx.x; k < H; k += HD_THREADS) { dr += gwr[k] * a.w[k]; df += gwf[k] * a.w[k]; }
  dr = block_sum(dr, red);
  df = block_sum(df, red);
  const float s0 = a.sigma[0], s1 = a.sigma[1];
  const float u0 = a.uv[0], u1 = a.uv[1 + H];
  for (int k = threadIdx.x; k < H; k += HD_THREADS)
    a.gw[k] = gwr[k] / s0 - dr / (s0 * s0) * u0 * a.uv[1 + k] + gwf[k] / s1 - df / (s1 * s1) * u1 * a.uv[1 + H + 1 + k];
  if (threadIdx.x == 0) {
    a.gb[0] = gbias;
    a.loss_val[0] = (a.scal[0] + (a.r1 ? 0.5f * a.gamma * a.r1[0] : 0.f)) * scale;
  }
}

struct AdvBwd {
  const float* p;      // (B)
  const float* wbar;   // (H)
  const float* gout;   // (1) upstream gradient of the scalar loss
  float* gy;           // (B,H)
  int B, H;
  float Bg;
};

// gen_step: g_adv = mean_b BCE(p_b, 1) over the global batch; D frozen -> only the input gradient
__global__ void __launch_bounds__(HD_THREADS) head_adv_bwd_kernel(AdvBwd a) {
  const float g = a.gout[0] / a.Bg;
  const int n = a.B * a.H;
  for (int i = blockIdx.x * HD_THREADS + threadIdx.x; i < n; i += gridDim.x * HD_THREADS) {
    const int b = i / a.H, k = i - b * a.H;
    a.gy[i] = g * bce_dz(a.p[b], 1.f) * a.wbar[k];
  }
}

}  // namespace

int tg_head_fwd_impl(cudaStream_t st, const float* yl, long long ld, int B, int H, int n_half, const float* w,
                     const float* bias, float* u, float* v, int training, const float* labels, float* wbar, float* uv,
                     float* sigma, float* p, float* stats) {
  TG_REQUIRE(yl && w && u && v && wbar && uv && sigma && p && stats, TG_ERR_ARG, "head_fwd: null pointer");
  TG_REQUIRE(B > 0 && H > 0 && H <= 4096 && (n_half == 1 || n_half == 2) && ld >= H, TG_ERR_SHAPE,
             "head_fwd: bad shape B=%d H=%d halves=%d ld=%lld", B, H, n_half, ld);
  HeadFwd a{yl, ld, w, bias, u, v, labels, wbar, uv, sigma, p, stats, B, H, n_half, training};
  const size_t smem = (size_t)(2 * H + n_half * H + HD_WARPS) * sizeof(float);
  head_fwd_kernel<<<1, HD_THREADS, smem, st>>>(a);
  return tg_check_launch("head_fwd");
}

int tg_head_seed_impl(cudaStream_t st, const float* p, const float* labels, const float* wbar, const float* stats,
                      float* scal, float* seed, float* gyf, int B, int H, float Bg, float target, float band) {
  TG_REQUIRE(p && labels && wbar && stats && scal && gyf, TG_ERR_ARG, "head_seed: null pointer");
  TG_REQUIRE(B > 0 && H > 0 && Bg > 0.f, TG_ERR_SHAPE, "head_seed: bad shape");
  HeadSeed a{p, labels, wbar, stats, scal, seed, gyf, B, H, Bg, target, band};
  int blocks = tg_ceil_div((long long)B * H, HD_THREADS * 4);
  if (blocks > 64) blocks = 64;
  head_seed_kernel<<<blocks, HD_THREADS, 0, st>>>(a);
  return tg_check_launch("head_seed");
}

int tg_head_bwd_impl(cudaStream_t st, const float* yl, long long ld, const float* hd, long long ld_hd, const float* p,
                     const float* labels, const float* w, const float* wbar, const float* uv, const float* sigma,
                     const float* scal, const float* r1, float* gyr, float* ghd, float* gw, float* gb, float* loss_val,
                     int B, int H, float Bg, float gamma) {
  TG_REQUIRE(yl && p && labels && w && wbar && uv && sigma && scal && gyr && gw && gb && loss_val, TG_ERR_ARG,
             "head_bwd: null pointer");
  TG_REQUIRE(!hd || ghd, TG_ERR_ARG, "head_bwd: tangent given without ghd");
  TG_REQUIRE(B > 0 && H > 0 && Bg > 0.f, TG_ERR_SHAPE, "head_bwd: bad shape");
  const size_t smem = (size_t)(3 * B + 2 * H + HD_WARPS) * sizeof(float);
  TG_REQUIRE(smem <= 200 * 1024, TG_ERR_UNSUPPORTED, "head_bwd: batch %d too large for the one-CTA head kernel", B);
  HeadBwd a{yl, ld, hd, ld_hd, p, labels, w, wbar, uv, sigma, scal, r1, gyr, ghd, gw, gb, loss_val, B, H, Bg, gamma};
  if (smem > 48 * 1024) { TG_OPT_IN_SMEM(head_bwd_kernel, "head_bwd"); }
  head_bwd_kernel<<<1, HD_THREADS, smem, st>>>(a);
  return tg_check_launch("head_bwd");
}

int tg_head_adv_bwd_impl(cudaStream_t st, const float* p, const float* wbar, const float* gout, float* gy, int B, int H,
                         float Bg) {
  TG_REQUIRE(p && wbar && gout && gy && B > 0 && H > 0 && Bg > 0.f, TG_ERR_ARG, "head_adv_bwd: bad arguments");
  AdvBwd a{p, wbar, gout, gy, B, H, Bg};
  int blocks = tg_ceil_div((long long)B * H, HD_THREADS * 4);
  if (blocks > 64) blocks = 64;
  head_adv_bwd_kernel<<<blocks, HD_THREADS, 0, st>>>(a);
  return tg_check_launch("head_adv_bwd");
}

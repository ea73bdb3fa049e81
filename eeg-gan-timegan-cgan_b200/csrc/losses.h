// Internal declarations for losses.cu / optim.cu / rng.cu (behind the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

size_t tg_reduce_ws_bytes();
int tg_sqdiff_sum_impl(cudaStream_t st, const float* a, const float* b, long long n, float* out, void* ws, size_t wsb);
int tg_scaled_diff_impl(cudaStream_t st, const float* a, const float* b, const float* coef, float* out, long long n,
                        int accumulate);
int tg_diff1_sum_impl(cudaStream_t st, const float* h, int B, int T, int H, float* out, void* ws, size_t wsb);
int tg_diff1_grad_impl(cudaStream_t st, const float* h, const float* coef, float* out, int B, int T, int H,
                       int accumulate);
int tg_center_scale_impl(cudaStream_t st, const float* x, const float* mean, const float* scale, float* out,
                         long long rows, int C);
int tg_acf_fwd_impl(cudaStream_t st, const float* xz, int B, int T, int C, int L, float* part);
int tg_acf_bwd_impl(cudaStream_t st, const float* xz, const float* S, int B, int T, int C, int L, float* gz,
                    float* stat);
int tg_acf_bwd_final_impl(cudaStream_t st, const float* gz, const float* xz, const float* mg, const float* kc,
                          const float* inv_s, float* dx, long long rows, int C, int accumulate);

// evaluation statistics (eval_stats.cu)
int tg_acf_score_impl(cudaStream_t st, const float* x, int N, int T, int C, int maxlag, double* out);

// optimiser
#define TG_MT_MAX 48   // tensors per multi-tensor launch
int tg_sumsq_multi_impl(cudaStream_t st, int n, const float* const* grads, const long long* sizes, float* out_sumsq,
                        void* ws, size_t wsb);
int tg_adam_multi_impl(cudaStream_t st, int n, float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, const long long* sizes, const float* sumsq, float max_norm, float lr,
                       float beta1, float beta2, float eps, int step, float grad_scale, float* dev_state);
int tg_snapshot_if_better_impl(cudaStream_t st, int n, float* const* dst, const float* const* src,
                               const long long* sizes, const float* value, float* best, float* best_step, float step);

// fused discriminator head (head.cu)
int tg_head_fwd_impl(cudaStream_t st, const float* yl, long long ld, int B, int H, int n_half, const float* w,
                     const float* bias, float* u, float* v, int training, const float* labels, float* wbar, float* uv,
                     float* sigma, float* p, float* stats);
int tg_head_seed_impl(cudaStream_t st, const float* p, const float* labels, const float* wbar, const float* stats,
                      float* scal, float* seed, float* gyf, int B, int H, float Bg, float target, float band);
int tg_head_bwd_impl(cudaStream_t st, const float* yl, long long ld, const float* hd, long long ld_hd, const float* p,
                     const float* labels, const float* w, const float* wbar, const float* uv, const float* sigma,
                     const float* scal, const float* r1, float* gyr, float* ghd, float* gw, float* gb, float* loss_val,
                     int B, int H, float Bg, float gamma);
int tg_head_adv_bwd_impl(cudaStream_t st, const float* p, const float* wbar, const float* gout, float* gy, int B, int H,
                         float Bg);

// rng
int tg_rng_uniform_impl(cudaStream_t st, float* out, long long n, unsigned long long seed, unsigned long long offset,
                        float lo, float hi, const unsigned long long* ctr);
int tg_rng_add_normal_impl(cudaStream_t st, const float* in, float* out, long long n, float std,
                           unsigned long long seed, unsigned long long offset, const unsigned long long* ctr,
                           const float* std_dev);

/* timegan_b200.h -- C ABI of libtimegan_b200.so: the B200 (sm_100a) hot path of the TimeGAN training step.
 *
 * The reference (Jeniya1378/eeg-gan-timegan-cgan) has no FFI layer: its hot path is the PyTorch calls made
 * from timeGAN/timegan_model.py and timeGAN/train_timegan.py.  Each entry point below replaces one of those
 * call sites (cited per function as file:line under /root/reference/timeGAN/); the Python host package
 * binds them with ctypes inside torch.autograd.Function wrappers.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all tensors are fp32, contiguous, batch-first (B,T,*).
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator); the library never
 *     allocates or frees device memory and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - return 0 = ok; negative = argument/shape/alignment error detected before launch (TG_ERR_*);
 *     positive = cudaError_t from the launch.  tg_last_error() returns a thread-local message.
 *   - no CPU fallback exists: without a CUDA device every compute call fails with a cudaError.
 */
#ifndef TIMEGAN_B200_H
#define TIMEGAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_ABI_VERSION 8

#define TG_OK 0
#define TG_ERR_ARG (-1)
#define TG_ERR_SHAPE (-2)
#define TG_ERR_ALIGN (-3)
#define TG_ERR_UNSUPPORTED (-4)

/* flags of the recurrent kernels (bits 8..15 optionally force sequences-per-CTA to 1, 2 or 4) */
#define TG_GRU_SAVE 1    /* forward: keep r,z,n (written over gi) and q for the backward pass */
#define TG_GRU_NO_BULK 2 /* use generic loads/stores instead of cp.async.bulk (testing) */
#define TG_GRU_DY_LAST 4 /* backward: the output gradient is (B,H) and applies to t = T-1 only */

/* projection precision */
#define TG_PROJ_FP32 0   /* CUDA-core FFMA, exact fp32 */
#define TG_PROJ_TF32 1   /* tcgen05 tensor cores, one TF32 pass over the fp32 operands (reduced precision, bound 2e-2) */
#define TG_PROJ_TF32X3 2 /* tcgen05 tensor cores, 3xTF32 split (hi*hi + hi*lo + lo*hi): fp32-parity mode, 1e-4 */

int tg_version(void);
const char* tg_last_error(void);
int tg_device_sm_count(void);
/* Process-wide tuning knobs.  "wgrad_ctas" = N > 0: the fused weight-gradient kernel (tg_wgrad_gru) uses at most N
 * CTAs (one per SM), leaving the other SMs to recurrent kernels issued on another stream; 0 = one CTA per SM.
 * Must not change between tg_wgrad_gru_workspace_bytes and the tg_wgrad_gru call that uses the workspace. */
int tg_set_option(const char* key, int value);
/* diagnostic: how many thread-block clusters of the H = 128 / 256 recurrent kernels (backward: 0 forward, 1 BPTT, 2 reverse-over-
 * tangent; groups = sequence groups per cluster) can be resident at once on this device; 0 = no such instantiation */
int tg_cluster_capacity(int H, int backward, int groups);

/* ---- launch accounting and per-family device timing (measurement only; used by bench.py) -------------------
 * tg_launch_count: kernels launched by this library since load.  With tg_prof_enable(1) every call below is
 * bracketed by a cudaEvent pair on its own stream; tg_prof_read(kind) synchronises those events and returns
 * the summed elapsed ms, the number of calls and the algorithmic bytes / FLOPs they declared since the last
 * tg_prof_reset().  Kinds 0..tg_prof_kinds()-1 are named by tg_prof_kind_name(). */
long long tg_launch_count(void);
int tg_prof_kinds(void);
const char* tg_prof_kind_name(int kind);
void tg_prof_enable(int on);
void tg_prof_reset(void);
int tg_prof_read(int kind, double* ms, long long* calls, double* bytes, double* flops);

/* ---- GRU layer, time-batched input projection:  C[M,N] (+)= A[M,K] W[N,K]^T + bias[N] -------------------
 * Replaces `params.linear_ih(input)` inside at::gru reached from timegan_model.py:33 (GRUStack.forward),
 * and the head Linears timegan_model.py:53 (Recovery.out), :66 (Generator.proj), :79 (Supervisor.proj).
 * bias may be NULL.  mode: TG_PROJ_FP32 | TG_PROJ_TF32 | TG_PROJ_TF32X3. */
int tg_proj(void* stream, const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
            int M, int N, int K, int accumulate, int mode);

/* ---- the "bf16 input projection" mode (BASELINE config c3): C16[M,N] = bf16(A) bf16(W16)^T + bias as bf16 -----------
 * A (M,K) fp32 is converted to bf16 inside the kernel (no bf16 copy of an activation is written to memory), W16 (N,K) is
 * bf16, the MMA is tcgen05 kind::f16 with fp32 accumulation, the result C16 (M,N) is bf16: 2 bytes per gate
 * pre-activation out of the projection and into tg_gru_fwd_bf16gi.  lda in floats, ldw / ldc in bf16 elements.
 * tg_bf16_gi_supported: 1 when both kernels take this layer (B sequences of T steps, K inputs, hidden size H: H = 64 or
 * 128, and not a batch for which the cluster forward kernel -- fp32 gi only -- is the faster recurrence). */
int tg_bf16_gi_supported(int B, int T, int K, int H);
int tg_proj_bf16(void* stream, const float* A, int lda, const void* W16, int ldw, const float* bias, void* C16, int ldc,
                 int M, int N, int K);

/* ---- dX[M,N] (+)= dG[M,K] W[K,N]  (autograd of the projection w.r.t. its input; loss.backward() at
 * train_timegan.py:140,159,219,267) */
int tg_dgrad(void* stream, const float* dG, int ldg, const float* W, int ldw, float* dX, int ldx, int M, int N, int K,
             int accumulate);

/* ---- weight gradients: dW[N,K] (+)= dG[M,N]^T A[M,K];  db[N] (+)= colsum(dG)  (db may be NULL) -------------
 * M = B*T.  a_shift_T > 0: A row m is taken as A[m-1] and as zero when m % a_shift_T == 0 (h_{t-1} read from
 * the layer output y).  Deterministic split-M reduction through `ws` (>= tg_wgrad_workspace_bytes). */
size_t tg_wgrad_workspace_bytes(int M, int N, int K);
int tg_wgrad(void* stream, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw, float* db, int M,
             int N, int K, int a_shift_T, int accumulate, void* ws, size_t ws_bytes, int mode /* TG_PROJ_* */);

/* ---- all weight gradients of one GRU layer in one pass over dGI, dq, x, y:
 * dW_ih (3H,I) (+)= dGI^T x, db_ih (+)= colsum(dGI), dW_hh (3H,H) (+)= [dGI[:, :2H] | dq]^T h_prev, db_hh likewise,
 * h_prev = y shifted by one step inside every sequence (zero at t = 0).  dgi (B*T,3H), dq (B*T,H), y (B*T,H)
 * contiguous; x (B*T,I) with leading dimension ldx, or NULL to skip dW_ih; db_* may be NULL. */
size_t tg_wgrad_gru_workspace_bytes(int B, int T, int I, int H);
int tg_wgrad_gru(void* stream, const float* dgi, const float* dq, const float* x, int ldx, const float* y, float* dW_ih,
                 float* dW_hh, float* db_ih, float* db_hh, int B, int T, int I, int H, int accumulate, void* ws,
                 size_t ws_bytes, int mode /* TG_PROJ_* */);

/* ---- persistent fused GRU layer forward (timegan_model.py:32-34 -> nn.GRU per-timestep loop) --------------
 * H <= 128: W_hh register-resident (gru_fwd.cu ...); 128 < H <= 1024: capacity fallback that streams W_hh from L2
 * every step (gru_bigh.cu; H % 4 == 0).
 * gi (B,T,3H) holds X W_ih^T + b_ih on entry; with TG_GRU_SAVE it holds r,z,n on exit and q (B,T,H) receives
 * h_{t-1} W_hn^T + b_hn.  y (B,T,H) receives h_t.  h0 = 0 (the reference never passes an initial state). */
int tg_gru_fwd(void* stream, float* gi, const float* w_hh, const float* b_hh, float* y, float* q, int B, int T, int H,
               int flags);

/* Same recurrence fed by tg_proj_bf16: gi16 (B,T,3H) bf16 is read-only (6H instead of 12H bytes per cell); with
 * TG_GRU_SAVE r,z,n go to rzn (B,T,3H) fp32 and q as above, so BPTT and the R1 passes read what they always read. */
int tg_gru_fwd_bf16gi(void* stream, const void* gi16, const float* w_hh, const float* b_hh, float* y, float* q,
                      float* rzn, int B, int T, int H, int flags);

/* ---- persistent BPTT of one layer (autograd of the nn.GRU loop; train_timegan.py:140,159,219,267,200) -----
 * in: dy (B,T,H) [or (B,H) with TG_GRU_DY_LAST], saved rzn,q, layer output y.  out: dgi (B,T,3H) =
 * [dar,daz,dan] (gradient of gi; also rows 0..2H of dGH) and dq (B,T,H) = dan*r (rows 2H..3H of dGH). */
int tg_gru_bwd(void* stream, const float* dy, const float* rzn, const float* q, const float* y, const float* w_hh,
               float* dgi, float* dq, int B, int T, int H, int flags,
               const float* w_hh_t /* (H,3H) = w_hh transposed; only read (and required) when H > 128 */);

/* ---- R1 penalty (train_timegan.py:198-202) without generic double backward: tangent forward ... ----------
 * gid (B,T,3H) holds xdot W_ih^T on entry and the tangent pre-activations a_r,a_z,a_n on exit. */
int tg_gru_jvp_fwd(void* stream, float* gid, const float* rzn, const float* q, const float* y, const float* w_hh,
                   float* ydot, float* qdot, int B, int T, int H, int flags);
/* ... and its reverse.  hbar/hdbar: adjoints of y / ydot ((B,H) with TG_GRU_DY_LAST).  Outputs are the
 * adjoints of the primal (gib,qb) and tangent (gidb,qdb) gate pre-activations, laid out like dgi/dq. */
int tg_gru_jvp_bwd(void* stream, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                   const float* ta, const float* qdot, const float* y, const float* ydot, const float* w_hh,
                   float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int flags,
                   const float* w_hh_t /* as for tg_gru_bwd */);

/* ---- losses (train_timegan.py:72-74 recon, :156-158 sup MSE, :79-80 first difference, :82-126 cov/ACF) ----
 * Scalars live on the device (float*), so no host synchronisation is needed between kernels. */
size_t tg_reduce_workspace_bytes(void);
int tg_sqdiff_sum(void* stream, const float* a, const float* b, long long n, float* out, void* ws, size_t ws_bytes);
int tg_scaled_diff(void* stream, const float* a, const float* b, const float* coef, float* out, long long n,
                   int accumulate); /* out (+)= coef[0]*(a-b) */
int tg_diff1_sum(void* stream, const float* h, int B, int T, int H, float* out, void* ws, size_t ws_bytes);
int tg_diff1_grad(void* stream, const float* h, const float* coef, float* out, int B, int T, int H, int accumulate);
int tg_center_scale(void* stream, const float* x, const float* mean, const float* scale, float* out, long long rows,
                    int C); /* out = (x-mean[c])*scale[c]; scale may be NULL */
/* out[N] (+)= column sums of X (M rows, leading dimension ld): per-channel means for cov/ACF, bias gradients */
size_t tg_colsum_workspace_bytes(int N);
int tg_colsum(void* stream, const float* X, int ld, int M, int N, float* out, int accumulate, void* ws,
              size_t ws_bytes);
int tg_acf_fwd(void* stream, const float* xz, int B, int T, int C, int L, float* part /* (B,L,C) */);
int tg_acf_bwd(void* stream, const float* xz, const float* S /* (L,C) */, int B, int T, int C, int L,
               float* gz /* (B,T,C) */, float* stat /* (B,2,C) */);
int tg_acf_bwd_final(void* stream, const float* gz, const float* xz, const float* mean_gz, const float* kc,
                     const float* inv_s, float* dx, long long rows, int C, int accumulate);

/* ---- evaluation statistics (timeGAN/evaluation.py:63-71,126-131) --------------------------------
 * out[n*C + c] (fp64) = autocorr_seq(x[n,:,c], maxlag): 0 if std < 1e-8, else the mean over lag = 1..min(maxlag,T-1)
 * of the Pearson correlation of x[:-lag] and x[lag:] (each slice with its own mean / variance, like np.corrcoef). */
int tg_acf_score(void* stream, const float* x /* (N,T,C) */, int N, int T, int C, int maxlag, double* out /* (N,C) */);

/* ---- clip_grad_norm_ + Adam (train_timegan.py:141-142,160-161,220-221,268-272) ---------------------------
 * Host arrays of n device pointers / element counts.  tg_sumsq writes sum g^2 over all tensors to a device
 * scalar; tg_adam applies g*grad_scale, the global-norm clip (max_norm <= 0: none) and one Adam step. */
size_t tg_sumsq_workspace_bytes(int n, const long long* sizes);
int tg_sumsq(void* stream, int n, const float* const* grads, const long long* sizes, float* out_sumsq, void* ws,
             size_t ws_bytes);
int tg_adam(void* stream, int n, float* const* params, const float* const* grads, float* const* exp_avg,
            float* const* exp_avg_sq, const long long* sizes, const float* sumsq, float max_norm, float lr, float beta1,
            float beta2, float eps, int step, float grad_scale,
            float* dev_state /* NULL, or device float[4] {lr, step, -, -}: the step counter and bias corrections
                                then live on the device (CUDA-graph replay); `lr`/`step` arguments are ignored */);

/* ---- fused discriminator head (timegan_model.py:92-98 spectral_norm(Linear(h,1)) + sigmoid; train_timegan.py:70 BCELoss,
 *      :196 loss, :199-202 R1 seed / sdot term, :205-215 accuracy + throttle, :241 g_adv) -------------------------------
 * y_last: the GRU stack's output at t = T-1, (n_half*B, H) rows `ld` floats apart; half 0 = real batch, half 1 = fake
 * batch (one forward call of D each, i.e. one power iteration each when `training`; u, v are updated in place like the
 * reference's buffers).  tg_head_fwd writes wbar (n_half,H) = w/sigma, uv (n_half,1+H), sigma (n_half), p (n_half*B)
 * and the four LOCAL batch sums stats = {sum bce_0, sum bce_1, #(p_0 > .5), #(p_1 < .5)} (labels NULL = all ones).
 * After the caller has summed `stats` over the ranks (data parallel), tg_head_seed turns them into scal = {loss_bce,
 * acc, scale} for a global batch of Bg samples per half and writes the R1 seed d(sum p_real)/d y_last (may be NULL)
 * and the fake half's input gradient; tg_head_bwd (after the tangent forward produced hd = d y_last(real)/d eps, NULL
 * without R1) writes the real half's gradients, the head's weight / bias gradients and the logged loss value. */
int tg_head_fwd(void* stream, const float* y_last, long long ld, int B, int H, int n_half, const float* w,
                const float* bias, float* u, float* v, int training, const float* labels, float* wbar, float* uv,
                float* sigma, float* p, float* stats);
int tg_head_seed(void* stream, const float* p, const float* labels, const float* wbar, const float* stats, float* scal,
                 float* seed, float* gyf, int B, int H, float Bg, float target, float band);
int tg_head_bwd(void* stream, const float* y_last, long long ld, const float* hd, long long ld_hd, const float* p,
                const float* labels, const float* w, const float* wbar, const float* uv, const float* sigma,
                const float* scal, const float* r1, float* gyr, float* ghd, float* gw, float* gb, float* loss_val, int B,
                int H, float Bg, float gamma);
/* gen_step (tt:241, D frozen): gy (B,H) = gout * d mean_b BCE(p_b, 1) / d y_last */
int tg_head_adv_bwd(void* stream, const float* p, const float* wbar, const float* gout, float* gy, int B, int H,
                    float Bg);

/* Best-checkpoint rule of train_timegan.py:410-413 ("if g_total < best_ckpt_loss: save_ckpt(best_path, ...)") without
 * a host round trip per step: when value[0] < best[0] (both device floats) the n source tensors (weights, Adam
 * moments) are copied into the snapshot tensors `dst`, then best[0] = value[0] and best_step[0] = step.  The host
 * reads best_step at its own pace and writes the snapshot to ckpt_best.pt when it has changed. */
int tg_snapshot_if_better(void* stream, int n, float* const* dst, const float* const* src, const long long* sizes,
                          const float* value, float* best, float* best_step, float step);

/* ---- noise (train_timegan.py:64-65 sample_noise, :46-47 add_instance_noise, :40-43 smooth_labels) --------
 * Philox4x32-10 keyed by (seed, offset); each call consumes ceil(n/4) counter values. */
int tg_rng_uniform(void* stream, float* out, long long n, unsigned long long seed, unsigned long long offset, float lo,
                   float hi, const unsigned long long* ctr /* NULL or device counter added to offset */);
int tg_rng_add_normal(void* stream, const float* in /* may be NULL */, float* out, long long n, float std,
                      unsigned long long seed, unsigned long long offset, const unsigned long long* ctr);
/* same, with the standard deviation read from device memory (one float): the instance-noise level of tt:352-353,404
 * decays every step while a captured CUDA graph stays the same */
int tg_rng_add_normal_dev(void* stream, const float* in /* may be NULL */, float* out, long long n,
                          const float* std_dev, unsigned long long seed, unsigned long long offset,
                          const unsigned long long* ctr);

/* ---- data-parallel all-reduce over NVLink peer memory (net-new: the reference is single-process; replaces the
 * torch.distributed all_reduce a DDP port of train_timegan.py:141,160,220,268 would issue) ------------------
 * Every rank owns one "peer region" (tg_peer_alloc: cudaMalloc'd + zeroed, the only device memory this library
 * allocates, because CUDA IPC needs a base allocation), exports it (64-byte cudaIpcMemHandle_t), and maps the
 * regions of the other ranks of the box (tg_peer_open).  tg_peer_allreduce SUMs n tensors IN PLACE across
 * `world` ranks in ONE ordinary kernel launch (graph-capturable, no host synchronisation): `regions[r]` is rank
 * r's region as seen from this process; data_off / flag_off locate this call site's staging area
 * (tg_peer_site_bytes) and flag words inside every region -- identical offsets on all ranks; `epoch` and
 * `status` are caller-owned LOCAL device words (the call site's replay counter and a peer-timeout flag).  All ranks must issue the same call sites with the same
 * tensor sizes; the sums are bit-identical on every rank. */
#define TG_PEER_MAX 8
int tg_peer_chunk_floats(void);
int tg_peer_alloc(void** ptr, size_t bytes);
int tg_peer_free(void* ptr);
int tg_peer_export(void* ptr, unsigned char* handle64);
int tg_peer_open(const unsigned char* handle64, void** peer_ptr);
int tg_peer_close(void* peer_ptr);
size_t tg_peer_site_bytes(int n, const long long* sizes, int world, size_t* flag_bytes);
int tg_peer_allreduce(void* stream, int rank, int world, void* const* regions, size_t data_off, size_t flag_off,
                      unsigned int* epoch /* device uint32[2], zero-initialised, one per call site */,
                      unsigned int* status /* device uint32 error word */, int n, float* const* tensors,
                      const long long* sizes);

#ifdef __cplusplus
}
#endif
#endif /* TIMEGAN_B200_H */

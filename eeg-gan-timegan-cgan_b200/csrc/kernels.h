// Internal C++ entry points behind the C ABI (api.cu forwards to these). Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// flags shared with include/timegan_b200.h
#define TG_GRU_SAVE 1      // forward: keep r,z,n (in place of gi) and q for BPTT
#define TG_GRU_NO_BULK 2   // force the generic (non-TMA) streaming path (testing)
#define TG_GRU_DY_LAST 4   // backward: dy is (B,H) and applies to t = T-1 only (discriminator head)

int tg_num_sms();
int tg_pick_bt(int B, int HP, int bt_override);

int tg_gru_fwd_impl(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T,
                    int H, int flags);
int tg_gru_bwd_impl(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y,
                    const float* whh, float* dgi, float* dq, int B, int T, int H, int flags, const float* whh_t);
int tg_gru_jvp_fwd_impl(cudaStream_t st, float* gid, const float* rzn, const float* q, const float* y,
                        const float* whh, float* ydot, float* qdot, int B, int T, int H, int flags);
int tg_gru_jvp_bwd_impl(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                        const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh,
                        float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int flags,
                        const float* whh_t);

// capacity fallback for H > 128 (gru_bigh.cu): W_hh streamed from L2 every step
int tg_bigh_fwd(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T, int H,
                int save);
int tg_bigh_bwd(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y, const float* whh_t,
                float* dgi, float* dq, int B, int T, int H, int dy_last);
int tg_bigh_jvp_fwd(cudaStream_t st, float* gid, const float* rzn, const float* q, const float* y, const float* whh,
                    float* ydot, float* qdot, int B, int T, int H);
int tg_bigh_jvp_bwd(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                    const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh_t,
                    float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int last_only);

// cluster kernels for H = 128 / 256 (gru_cluster.cu): W_hh split over the register files of 2 / 8 SMs, state
// all-gathered through distributed shared memory every step
bool tg_cluster_takes(int H, int B, bool backward);
int tg_use_cluster();
int tg_cluster_jvp256();
int tg_cluster_dio();
int tg_cluster_no();
int tg_gru_cl_fwd(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T, int H,
                  int save);
int tg_gru_cl_bwd(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y, const float* whh,
                  float* dgi, float* dq, int B, int T, int H, int dy_last);

bool tg_cluster_takes_jvp_bwd(int H, int B);
int tg_gru_cl_jvp_bwd(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                      const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh, float* gib,
                      float* qb, float* gidb, float* qdb, int B, int T, int H, int last_only);

// time-batched contractions (FFMA baseline path; fp32 exact)
int tg_gemm_nt_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                    int ldc, int M, int N, int K, int accumulate);
int tg_gemm_nn_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N,
                    int K, int accumulate);
int tg_wgrad_impl(cudaStream_t st, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw, float* db,
                  int M, int N, int K, int a_shift_T, int accumulate, float* ws, size_t ws_bytes);
size_t tg_wgrad_ws_bytes(int M, int N, int K);

// tensor-core projection (tcgen05 + TMA); returns TG_ERR_UNSUPPORTED for shapes it cannot take
int tg_proj_tc_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                    int ldc, int M, int N, int K, int accumulate, int passes);

// bf16 input projection (proj_bf16.cu): fp32 A converted in shared memory, bf16 W, bf16 result; and the forward
// recurrence that reads it (gru_fwd.cu, H = 64 / 128)
int tg_proj_bf16_impl(cudaStream_t st, const float* A, int lda, const void* W16, int ldw, const float* bias, void* C16,
                      int ldc, int M, int N, int K);
bool tg_gru_fwd_bf16gi_ok(int H);
int tg_gru_fwd_bf16gi_impl(cudaStream_t st, const void* gi16, const float* whh, const float* bhh, float* y, float* q,
                           float* rzn, int B, int T, int H, int flags);

// tensor-core weight gradient (tcgen05, MN-major operands); TG_ERR_UNSUPPORTED for shapes it cannot take
size_t tg_wgrad_tc_ws_bytes(int M, int N, int K);
int tg_wgrad_tc_impl(cudaStream_t st, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw,
                     float* db, int M, int N, int K, int a_shift_T, int accumulate, float* ws, size_t ws_bytes,
                     int passes);
size_t tg_wgrad_gru_ws_bytes(int M, int I, int H);
int tg_wgrad_gru_tc_impl(cudaStream_t st, const float* dgi, const float* dq, const float* x, int ldx, const float* y,
                         float* dW_ih, float* dW_hh, float* db_ih, float* db_hh, int B, int T, int I, int H,
                         int accumulate, float* ws, size_t ws_bytes, int passes);
int tg_max_optin_smem();
int tg_gemm_smem_budget();
int tg_long_chunks();
int tg_bwd_pair();
int tg_wgrad_cta_cap();
int tg_peer_timeout_ms();

// column sums out[N] (+)= sum_m X[m*ld + n]; ws >= tg_colsum_ws_bytes(N)
size_t tg_colsum_ws_bytes(int N);
int tg_colsum_impl(cudaStream_t st, const float* X, int ld, int M, int N, float* out, int accumulate, float* ws,
                   size_t ws_bytes);

// extern "C" surface of libtimegan_b200.so (see include/timegan_b200.h). Thin forwarding + error plumbing.
#include "../../include/timegan_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "losses.h"
#include <atomic>
#include <mutex>
#include <string.h>
#include <stdlib.h>
#include <vector>

static thread_local char g_err[512] = "";

void tg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting + optional per-family CUDA-event profiling (bench.py's roofline numbers) -------------
// Every kernel launch goes through tg_check_launch, which counts it.  With profiling enabled each C-ABI call
// is bracketed by a cudaEvent pair on ITS stream; tg_prof_read sums the elapsed times per kernel family
// together with the algorithmic bytes / FLOPs the calls declared.  Disabled (the default) it costs nothing.
namespace {
enum { K_GRU_FWD = 0, K_GRU_BWD, K_JVP_FWD, K_JVP_BWD, K_PROJ, K_DGRAD, K_WGRAD, K_LOSS, K_OPTIM, K_RNG, K_COUNT };
const char* const kKindNames[K_COUNT] = {"gru_fwd", "gru_bwd", "gru_jvp_fwd", "gru_jvp_bwd", "proj",
                                         "dgrad",   "wgrad",   "loss",        "optim",       "rng"};
struct ProfRec { cudaEvent_t a, b; int kind; };
std::mutex g_prof_mu;
std::atomic<long long> g_launches{0};
bool g_prof_on = false;
std::vector<ProfRec> g_recs;          // recorded pairs of the current session
std::vector<ProfRec> g_pool;          // reusable events
double g_bytes[K_COUNT], g_flops[K_COUNT];
long long g_calls[K_COUNT];
constexpr size_t kMaxRecs = 1 << 17;

struct ProfScope {
  cudaStream_t st; int idx = -1;
  ProfScope(void* stream, int kind, double bytes, double flops) : st((cudaStream_t)stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_bytes[kind] += bytes; g_flops[kind] += flops; g_calls[kind] += 1;
    if (g_recs.size() >= kMaxRecs) return;
    ProfRec r;
    if (!g_pool.empty()) { r = g_pool.back(); g_pool.pop_back(); }
    else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    r.kind = kind;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    idx = (int)g_recs.size() - 1;
  }
  ~ProfScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_recs[idx].b, st);
  }
};
}  // namespace

int tg_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    tg_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int tg_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      return 148;  // B200; do not cache a failed query
  }
  return sms;
}

int tg_max_optin_smem() {
  static std::atomic<int> v{0};
  int x = v.load();
  if (x == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&x, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || x <= 0)
      return 227 * 1024;
    v.store(x);
  }
  return x;
}

// Shared-memory budget of the persistent tensor-core GEMM CTAs.  Default: the whole opt-in maximum (one CTA owns
// the SM).  A smaller budget (TIMEGAN_B200_GEMM_SMEM_KB) leaves room for CTAs of the recurrent kernels issued on
// another stream to be co-resident, at the price of shallower TMA rings.
int tg_gemm_smem_budget() {
  static std::atomic<int> v{0};
  int x = v.load();
  if (x == 0) {
    x = tg_max_optin_smem();
    const char* e = getenv("TIMEGAN_B200_GEMM_SMEM_KB");
    if (e && atoi(e) >= 48 && atoi(e) * 1024 < x) x = atoi(e) * 1024;
    v.store(x);
  }
  return x;
}

// long chunks (16 timesteps per ring stage instead of 8) for the one-sequence-per-CTA recurrent kernels;
// TIMEGAN_B200_LONG_CHUNKS=0 switches back
int tg_long_chunks() {
  static std::atomic<int> v{-1};
  int x = v.load();
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_LONG_CHUNKS");
    x = (e && atoi(e) == 0) ? 0 : 1;
    v.store(x);
  }
  return x;
}

// cluster kernels (gru_cluster.cu) for H = 128 / 256: 1 = where they are the faster kernel (default), 0 = never,
// 2 = for every H = 128 / 256 launch (TIMEGAN_B200_CLUSTER / tg_set_option("cluster", v))
static std::atomic<int> g_use_cluster{-1};
int tg_use_cluster() {
  int x = g_use_cluster.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_CLUSTER");
    x = e ? atoi(e) : 1;
    if (x < 0 || x > 2) x = 1;
    g_use_cluster.store(x);
  }
  return x;
}

// cluster reverse-over-tangent kernel at H = 256 (one group of 8 sequences per 8-CTA cluster): on by default where it
// is instantiated; TIMEGAN_B200_CLUSTER_JVP256=0 / tg_set_option("cluster_jvp256", 0) keeps the L2-streaming kernel
static std::atomic<int> g_cluster_jvp256{-1};
int tg_cluster_jvp256() {
  int x = g_cluster_jvp256.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_CLUSTER_JVP256");
    x = (e && atoi(e) == 0) ? 0 : 1;
    g_cluster_jvp256.store(x);
  }
  return x;
}

// direct-I/O variant of the cluster BPTT kernel (registers instead of the shared-memory input ring / output staging, no
// block barrier per step).  OFF: measured slower on a B200 (H = 128, B = 256: 1302 vs 1057 us per pass) -- the per-lane
// global loads / stores (four 32-byte sectors per warp instruction) go through the same LSU pipe as the mat-vec's LDS.128
// operand stream, which is what the kernel waits on.  TIMEGAN_B200_CLUSTER_DIO=1 / tg_set_option("cluster_dio", 1).
static std::atomic<int> g_cluster_dio{-1};
int tg_cluster_dio() {
  int x = g_cluster_dio.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_CLUSTER_DIO");
    x = (e && atoi(e) != 0) ? 1 : 0;
    g_cluster_dio.store(x);
  }
  return x;
}

// output columns per thread in the H = 128 cluster BPTT kernel (2 or 4): TIMEGAN_B200_CLUSTER_NO / tg_set_option("cluster_no", v)
static std::atomic<int> g_cluster_no{-1};
int tg_cluster_no() {
  int x = g_cluster_no.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_CLUSTER_NO");
    x = (e && (atoi(e) == 2 || atoi(e) == 8)) ? atoi(e) : 4;     // 8: also two hidden units per thread in the H = 128 forward
    g_cluster_no.store(x);
  }
  return x;
}

// two-columns-per-thread BPTT kernel for one-sequence-per-CTA launches at H <= 64 (gru_bwd.cu): OFF by default -- it
// halves the shared-memory operand fetches (12 instead of 24 LDS.128 per step) but pays a second shuffle round on the
// per-step dependency chain, and measured 283 vs 275 us at the c2 layer shape (profiles/r02_probe_bwd_pair.log).
// TIMEGAN_B200_BWD_PAIR=1 / tg_set_option("bwd_pair", 1) selects it.
static std::atomic<int> g_bwd_pair{-1};
int tg_bwd_pair() {
  int x = g_bwd_pair.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_BWD_PAIR");
    x = (e && atoi(e) != 0) ? 1 : 0;
    g_bwd_pair.store(x);
  }
  return x;
}

static std::atomic<int> g_wgrad_cta_cap{0};
int tg_wgrad_cta_cap() { return g_wgrad_cta_cap.load(std::memory_order_relaxed); }

// how long a peer all-reduce CTA waits for another rank before it reports failure (TIMEGAN_B200_PEER_TIMEOUT_MS,
// tg_set_option("peer_timeout_ms")): long enough for a rank that is busy writing a checkpoint or collecting garbage
static std::atomic<int> g_peer_timeout_ms{-1};
int tg_peer_timeout_ms() {
  int x = g_peer_timeout_ms.load(std::memory_order_relaxed);
  if (x < 0) {
    const char* e = getenv("TIMEGAN_B200_PEER_TIMEOUT_MS");
    x = (e && atoi(e) > 0) ? atoi(e) : 10000;
    g_peer_timeout_ms.store(x);
  }
  return x;
}

size_t tg_sumsq_ws_bytes(int n, const long long* sizes);

extern "C" {

int tg_version(void) { return TG_ABI_VERSION; }
const char* tg_last_error(void) { return g_err; }
int tg_device_sm_count(void) { return tg_num_sms(); }
int tg_set_option(const char* key, int value) {
  if (key && strcmp(key, "wgrad_ctas") == 0) { g_wgrad_cta_cap.store(value < 0 ? 0 : value); return TG_OK; }
  if (key && strcmp(key, "bwd_pair") == 0) { g_bwd_pair.store(value ? 1 : 0); return TG_OK; }
  if (key && strcmp(key, "cluster_no") == 0) { g_cluster_no.store((value == 2 || value == 8) ? value : 4); return TG_OK; }
  if (key && strcmp(key, "cluster_dio") == 0) { g_cluster_dio.store(value ? 1 : 0); return TG_OK; }
  if (key && strcmp(key, "cluster_jvp256") == 0) { g_cluster_jvp256.store(value ? 1 : 0); return TG_OK; }
  if (key && strcmp(key, "cluster") == 0) { g_use_cluster.store(value < 0 || value > 2 ? 1 : value); return TG_OK; }
  if (key && strcmp(key, "peer_timeout_ms") == 0) { g_peer_timeout_ms.store(value < 1 ? 1 : value); return TG_OK; }
  tg_set_error("set_option: unknown key '%s'", key ? key : "(null)");
  return TG_ERR_ARG;
}

long long tg_launch_count(void) { return g_launches.load(); }
int tg_prof_kinds(void) { return K_COUNT; }
const char* tg_prof_kind_name(int kind) { return (kind >= 0 && kind < K_COUNT) ? kKindNames[kind] : ""; }
void tg_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}
void tg_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) g_pool.push_back(r);
  g_recs.clear();
  for (int k = 0; k < K_COUNT; ++k) { g_bytes[k] = g_flops[k] = 0.0; g_calls[k] = 0; }
}
int tg_prof_read(int kind, double* ms, long long* calls, double* bytes, double* flops) {
  if (kind < 0 || kind >= K_COUNT) { tg_set_error("prof_read: bad kind %d", kind); return TG_ERR_ARG; }
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double total = 0.0;
  for (auto& r : g_recs) {
    if (r.kind != kind) continue;
    cudaError_t e = cudaEventSynchronize(r.b);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) { tg_set_error("prof_read: %s", cudaGetErrorString(e)); return (int)e; }
    total += t;
  }
  if (ms) *ms = total;
  if (calls) *calls = g_calls[kind];
  if (bytes) *bytes = g_bytes[kind];
  if (flops) *flops = g_flops[kind];
  return TG_OK;
}

int tg_proj(void* stream, const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
            int M, int N, int K, int accumulate, int mode) {
  ProfScope _ps(stream, K_PROJ, 4.0 * ((double)M * K + (double)N * K + (double)M * N), 2.0 * M * N * K);
  if (mode == TG_PROJ_TF32 || mode == TG_PROJ_TF32X3) {
    int rc = tg_proj_tc_impl((cudaStream_t)stream, A, lda, W, ldw, bias, C, ldc, M, N, K, accumulate,
                             mode == TG_PROJ_TF32X3 ? 3 : 1);
    if (rc != TG_ERR_UNSUPPORTED) return rc;  // shapes the tensor-core tile cannot take run on the FFMA path
  }
  return tg_gemm_nt_impl((cudaStream_t)stream, A, lda, W, ldw, bias, C, ldc, M, N, K, accumulate);
}

int tg_bf16_gi_supported(int B, int T, int K, int H) {
  const long long M = (long long)B * T;
  // a batch that the cluster forward kernel takes (H = 128, more sequences than one-SM CTAs fit: the discriminator's 2B pass)
  // keeps the fp32-gi path: that kernel has no bf16-input variant and is the faster one there (1367 vs 1557 us at B = 512)
  if (tg_cluster_takes(H, B, false)) return 0;
  return (tg_gru_fwd_bf16gi_ok(H) && M >= 128 && K % 4 == 0 && K <= 512) ? 1 : 0;
}

int tg_proj_bf16(void* stream, const float* A, int lda, const void* W16, int ldw, const float* bias, void* C16, int ldc,
                 int M, int N, int K) {
  ProfScope _ps(stream, K_PROJ, 4.0 * (double)M * K + 2.0 * (double)N * K + 2.0 * (double)M * N, 2.0 * M * N * K);
  return tg_proj_bf16_impl((cudaStream_t)stream, A, lda, W16, ldw, bias, C16, ldc, M, N, K);
}

int tg_dgrad(void* stream, const float* dG, int ldg, const float* W, int ldw, float* dX, int ldx, int M, int N, int K,
             int accumulate) {
  ProfScope _ps(stream, K_DGRAD, 4.0 * ((double)M * K + (double)N * K + (double)M * N), 2.0 * M * N * K);
  return tg_gemm_nn_impl((cudaStream_t)stream, dG, ldg, W, ldw, dX, ldx, M, N, K, accumulate);
}

size_t tg_wgrad_workspace_bytes(int M, int N, int K) {
  const size_t a = tg_wgrad_ws_bytes(M, N, K), b = tg_wgrad_tc_ws_bytes(M, N, K);
  return a > b ? a : b;
}

int tg_wgrad(void* stream, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw, float* db, int M,
             int N, int K, int a_shift_T, int accumulate, void* ws, size_t ws_bytes, int mode) {
  ProfScope _ps(stream, K_WGRAD, 4.0 * ((double)M * N + (double)M * K + (double)N * K), 2.0 * M * N * K);
  if (mode == TG_PROJ_TF32 || mode == TG_PROJ_TF32X3) {
    int rc = tg_wgrad_tc_impl((cudaStream_t)stream, dG, ldg, A, lda, dW, lddw, db, M, N, K, a_shift_T, accumulate,
                              (float*)ws, ws_bytes, mode == TG_PROJ_TF32X3 ? 3 : 1);
    if (rc != TG_ERR_UNSUPPORTED) return rc;
  }
  return tg_wgrad_impl((cudaStream_t)stream, dG, ldg, A, lda, dW, lddw, db, M, N, K, a_shift_T, accumulate, (float*)ws,
                       ws_bytes);
}

size_t tg_wgrad_gru_workspace_bytes(int B, int T, int I, int H) {
  const int M = B * T;
  size_t need = tg_wgrad_gru_ws_bytes(M, I, H);
  const size_t a = tg_wgrad_workspace_bytes(M, 3 * H, I > 0 ? I : 1), b = tg_wgrad_workspace_bytes(M, 2 * H, H),
               c = tg_wgrad_workspace_bytes(M, H, H);
  if (a > need) need = a;
  if (b > need) need = b;
  if (c > need) need = c;
  return need;
}

int tg_wgrad_gru(void* stream, const float* dgi, const float* dq, const float* x, int ldx, const float* y, float* dW_ih,
                 float* dW_hh, float* db_ih, float* db_hh, int B, int T, int I, int H, int accumulate, void* ws,
                 size_t ws_bytes, int mode) {
  const int M = B * T;
  if (mode == TG_PROJ_TF32 || mode == TG_PROJ_TF32X3) {
    ProfScope _ps(stream, K_WGRAD, 4.0 * M * (4.0 * H + (x ? I : 0) + H), 2.0 * M * 3.0 * H * ((x ? I : 0) + H));
    int rc = tg_wgrad_gru_tc_impl((cudaStream_t)stream, dgi, dq, x, ldx, y, dW_ih, dW_hh, db_ih, db_hh, B, T, I, H,
                                  accumulate, (float*)ws, ws_bytes, mode == TG_PROJ_TF32X3 ? 3 : 1);
    if (rc != TG_ERR_UNSUPPORTED) return rc;
  }
  // shapes the fused tile cannot take: the three contractions one by one (each picks its own best kernel)
  int rc = 0;
  if (x) {
    rc = tg_wgrad(stream, dgi, 3 * H, x, ldx, dW_ih, I, db_ih, M, 3 * H, I, 0, accumulate, ws, ws_bytes, mode);
    if (rc) return rc;
  }
  rc = tg_wgrad(stream, dgi, 3 * H, y, H, dW_hh, H, db_hh, M, 2 * H, H, T, accumulate, ws, ws_bytes, mode);
  if (rc) return rc;
  return tg_wgrad(stream, dq, H, y, H, dW_hh + (size_t)2 * H * H, H, db_hh ? db_hh + 2 * H : nullptr, M, H, H, T,
                  accumulate, ws, ws_bytes, mode);
}

int tg_gru_fwd(void* stream, float* gi, const float* w_hh, const float* b_hh, float* y, float* q, int B, int T, int H,
               int flags) {
  ProfScope _ps(stream, K_GRU_FWD, (double)B * T * H * ((flags & TG_GRU_SAVE) ? 32.0 : 16.0), 6.0 * B * T * (double)H * H);
  return tg_gru_fwd_impl((cudaStream_t)stream, gi, w_hh, b_hh, y, q, B, T, H, flags);
}

int tg_gru_fwd_bf16gi(void* stream, const void* gi16, const float* w_hh, const float* b_hh, float* y, float* q,
                      float* rzn, int B, int T, int H, int flags) {
  ProfScope _ps(stream, K_GRU_FWD, (double)B * T * H * ((flags & TG_GRU_SAVE) ? 26.0 : 10.0), 6.0 * B * T * (double)H * H);
  return tg_gru_fwd_bf16gi_impl((cudaStream_t)stream, gi16, w_hh, b_hh, y, q, rzn, B, T, H, flags);
}

int tg_gru_bwd(void* stream, const float* dy, const float* rzn, const float* q, const float* y, const float* w_hh,
               float* dgi, float* dq, int B, int T, int H, int flags, const float* w_hh_t) {
  ProfScope _ps(stream, K_GRU_BWD, (double)B * T * H * ((flags & TG_GRU_DY_LAST) ? 36.0 : 40.0), 6.0 * B * T * (double)H * H);
  return tg_gru_bwd_impl((cudaStream_t)stream, dy, rzn, q, y, w_hh, dgi, dq, B, T, H, flags, w_hh_t);
}

int tg_gru_jvp_fwd(void* stream, float* gid, const float* rzn, const float* q, const float* y, const float* w_hh,
                   float* ydot, float* qdot, int B, int T, int H, int flags) {
  ProfScope _ps(stream, K_JVP_FWD, (double)B * T * H * 56.0, 6.0 * B * T * (double)H * H);
  return tg_gru_jvp_fwd_impl((cudaStream_t)stream, gid, rzn, q, y, w_hh, ydot, qdot, B, T, H, flags);
}

int tg_gru_jvp_bwd(void* stream, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                   const float* ta, const float* qdot, const float* y, const float* ydot, const float* w_hh,
                   float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int flags,
                   const float* w_hh_t) {
  ProfScope _ps(stream, K_JVP_BWD, (double)B * T * H * 104.0, 12.0 * B * T * (double)H * H);
  return tg_gru_jvp_bwd_impl((cudaStream_t)stream, hbar, hdbar, rzn, q, ta, qdot, y, ydot, w_hh, gib, qb, gidb, qdb, B,
                             T, H, flags, w_hh_t);
}

size_t tg_reduce_workspace_bytes(void) { return tg_reduce_ws_bytes(); }
int tg_sqdiff_sum(void* stream, const float* a, const float* b, long long n, float* out, void* ws, size_t ws_bytes) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_sqdiff_sum_impl((cudaStream_t)stream, a, b, n, out, ws, ws_bytes);
}
int tg_scaled_diff(void* stream, const float* a, const float* b, const float* coef, float* out, long long n,
                   int accumulate) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_scaled_diff_impl((cudaStream_t)stream, a, b, coef, out, n, accumulate);
}
int tg_diff1_sum(void* stream, const float* h, int B, int T, int H, float* out, void* ws, size_t ws_bytes) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_diff1_sum_impl((cudaStream_t)stream, h, B, T, H, out, ws, ws_bytes);
}
int tg_diff1_grad(void* stream, const float* h, const float* coef, float* out, int B, int T, int H, int accumulate) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_diff1_grad_impl((cudaStream_t)stream, h, coef, out, B, T, H, accumulate);
}
int tg_center_scale(void* stream, const float* x, const float* mean, const float* scale, float* out, long long rows,
                    int C) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_center_scale_impl((cudaStream_t)stream, x, mean, scale, out, rows, C);
}
int tg_acf_fwd(void* stream, const float* xz, int B, int T, int C, int L, float* part) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_acf_fwd_impl((cudaStream_t)stream, xz, B, T, C, L, part);
}
int tg_acf_bwd(void* stream, const float* xz, const float* S, int B, int T, int C, int L, float* gz, float* stat) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_acf_bwd_impl((cudaStream_t)stream, xz, S, B, T, C, L, gz, stat);
}
int tg_acf_bwd_final(void* stream, const float* gz, const float* xz, const float* mean_gz, const float* kc,
                     const float* inv_s, float* dx, long long rows, int C, int accumulate) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_acf_bwd_final_impl((cudaStream_t)stream, gz, xz, mean_gz, kc, inv_s, dx, rows, C, accumulate);
}

int tg_acf_score(void* stream, const float* x, int N, int T, int C, int maxlag, double* out) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_acf_score_impl((cudaStream_t)stream, x, N, T, C, maxlag, out);
}

size_t tg_colsum_workspace_bytes(int N) { return tg_colsum_ws_bytes(N); }
int tg_colsum(void* stream, const float* X, int ld, int M, int N, float* out, int accumulate, void* ws,
              size_t ws_bytes) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_colsum_impl((cudaStream_t)stream, X, ld, M, N, out, accumulate, (float*)ws, ws_bytes);
}

size_t tg_sumsq_workspace_bytes(int n, const long long* sizes) { return tg_sumsq_ws_bytes(n, sizes); }
int tg_sumsq(void* stream, int n, const float* const* grads, const long long* sizes, float* out_sumsq, void* ws,
             size_t ws_bytes) {
  ProfScope _ps(stream, K_OPTIM, 0.0, 0.0);
  return tg_sumsq_multi_impl((cudaStream_t)stream, n, grads, sizes, out_sumsq, ws, ws_bytes);
}
int tg_adam(void* stream, int n, float* const* params, const float* const* grads, float* const* exp_avg,
            float* const* exp_avg_sq, const long long* sizes, const float* sumsq, float max_norm, float lr, float beta1,
            float beta2, float eps, int step, float grad_scale, float* dev_state) {
  ProfScope _ps(stream, K_OPTIM, 0.0, 0.0);
  return tg_adam_multi_impl((cudaStream_t)stream, n, params, grads, exp_avg, exp_avg_sq, sizes, sumsq, max_norm, lr,
                            beta1, beta2, eps, step, grad_scale, dev_state);
}

int tg_head_fwd(void* stream, const float* y_last, long long ld, int B, int H, int n_half, const float* w,
                const float* bias, float* u, float* v, int training, const float* labels, float* wbar, float* uv,
                float* sigma, float* p, float* stats) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_head_fwd_impl((cudaStream_t)stream, y_last, ld, B, H, n_half, w, bias, u, v, training, labels, wbar, uv,
                          sigma, p, stats);
}
int tg_head_seed(void* stream, const float* p, const float* labels, const float* wbar, const float* stats, float* scal,
                 float* seed, float* gyf, int B, int H, float Bg, float target, float band) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_head_seed_impl((cudaStream_t)stream, p, labels, wbar, stats, scal, seed, gyf, B, H, Bg, target, band);
}
int tg_head_bwd(void* stream, const float* y_last, long long ld, const float* hd, long long ld_hd, const float* p,
                const float* labels, const float* w, const float* wbar, const float* uv, const float* sigma,
                const float* scal, const float* r1, float* gyr, float* ghd, float* gw, float* gb, float* loss_val, int B,
                int H, float Bg, float gamma) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_head_bwd_impl((cudaStream_t)stream, y_last, ld, hd, ld_hd, p, labels, w, wbar, uv, sigma, scal, r1, gyr, ghd,
                          gw, gb, loss_val, B, H, Bg, gamma);
}
int tg_head_adv_bwd(void* stream, const float* p, const float* wbar, const float* gout, float* gy, int B, int H,
                    float Bg) {
  ProfScope _ps(stream, K_LOSS, 0.0, 0.0);
  return tg_head_adv_bwd_impl((cudaStream_t)stream, p, wbar, gout, gy, B, H, Bg);
}

int tg_snapshot_if_better(void* stream, int n, float* const* dst, const float* const* src, const long long* sizes,
                          const float* value, float* best, float* best_step, float step) {
  ProfScope _ps(stream, K_OPTIM, 0.0, 0.0);
  return tg_snapshot_if_better_impl((cudaStream_t)stream, n, dst, src, sizes, value, best, best_step, step);
}

int tg_rng_uniform(void* stream, float* out, long long n, unsigned long long seed, unsigned long long offset, float lo,
                   float hi, const unsigned long long* ctr) {
  ProfScope _ps(stream, K_RNG, 4.0 * (double)n, 0.0);
  return tg_rng_uniform_impl((cudaStream_t)stream, out, n, seed, offset, lo, hi, ctr);
}
int tg_rng_add_normal(void* stream, const float* in, float* out, long long n, float std, unsigned long long seed,
                      unsigned long long offset, const unsigned long long* ctr) {
  ProfScope _ps(stream, K_RNG, 8.0 * (double)n, 0.0);
  return tg_rng_add_normal_impl((cudaStream_t)stream, in, out, n, std, seed, offset, ctr, nullptr);
}
int tg_rng_add_normal_dev(void* stream, const float* in, float* out, long long n, const float* std_dev,
                          unsigned long long seed, unsigned long long offset, const unsigned long long* ctr) {
  ProfScope _ps(stream, K_RNG, 8.0 * (double)n, 0.0);
  TG_REQUIRE(std_dev, TG_ERR_ARG, "rng_add_normal_dev: null std pointer");
  return tg_rng_add_normal_impl((cudaStream_t)stream, in, out, n, 0.f, seed, offset, ctr, std_dev);
}

}  // extern "C"

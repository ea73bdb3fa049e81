// fp32-exact time-batched contractions on CUDA cores (FFMA), used where 1e-4 parity is required and as
// the always-available baseline next to the tcgen05 projection path (proj_tcgen05.cu).
//
//   gemm_nt : C[M,N]  = A[M,K] W[N,K]^T (+bias)   -- the input projection X W_ih^T + b_ih inside at::gru
//                                                  (timegan_model.py:33) and the heads (tm:53,66,79)
//   gemm_nn : C[M,N]  = A[M,K] W[K,N]             -- dX = dGI W_ih
//   wgrad   : dW[N,K] = dG[M,N]^T A[M,K], db = colsum(dG), reduction over M = B*T split across CTAs and
//             reduced in a fixed order (deterministic, no float atomics)  -- dW_ih, dW_hh, db_ih, db_hh
//
// 128x64x16 block tile, 256 threads, 8x4 register micro-tile, register-prefetched smem double step.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, NTHREADS = 256;

// ---- operand accessors: a(m,k) and b(k,n) with bounds guards -----------------------------------
struct RowMajorMK {  // a(m,k) = P[m*ld + k]
  const float* P; int ld, M, K;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < M && k < K) ? P[(size_t)m * ld + k] : 0.f; }
};
struct RowMajorKN_T {  // b(k,n) = P[n*ld + k]   (weights stored (N,K))
  const float* P; int ld, K, N;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < N && k < K) ? P[(size_t)n * ld + k] : 0.f; }
};
struct RowMajorKN {  // b(k,n) = P[k*ld + n]
  const float* P; int ld, K, N;
  __device__ __forceinline__ float operator()(int k, int n) const { return (n < N && k < K) ? P[(size_t)k * ld + n] : 0.f; }
};
struct ColMajorMK {  // a(m,k) = P[k*ld + m]   (dG^T: m indexes gate columns, k indexes (b,t) rows)
  const float* P; int ld, M, K;
  __device__ __forceinline__ float operator()(int m, int k) const { return (m < M && k < K) ? P[(size_t)k * ld + m] : 0.f; }
};
struct ShiftedKN {  // b(k,n) = (k % T == 0) ? 0 : P[(k-1)*ld + n]   (h_{t-1} from y)
  const float* P; int ld, K, N, T;
  __device__ __forceinline__ float operator()(int k, int n) const {
    return (n < N && k < K && (k % T) != 0) ? P[(size_t)(k - 1) * ld + n] : 0.f;
  }
};

struct EpiStore {  // C[m*ldc+n] (+)= v + bias[n]
  float* C; int ldc; const float* bias; int accumulate;
  __device__ __forceinline__ void operator()(int m, int n, float v, int) const {
    float* c = C + (size_t)m * ldc + n;
    if (bias) v += bias[n];
    *c = accumulate ? (*c + v) : v;
  }
};
struct EpiPartial {  // ws[z][m][n] = v
  float* ws; int M, N;
  __device__ __forceinline__ void operator()(int m, int n, float v, int z) const { ws[((size_t)z * M + m) * N + n] = v; }
};

// A_KFAST: consecutive threads walk k (A is k-contiguous); else they walk m.
template <bool A_KFAST, bool B_NFAST, class LA, class LB, class EPI>
__global__ void __launch_bounds__(NTHREADS) sgemm_kernel(LA la, LB lb, EPI epi, int M, int N, int K, int klen) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * klen, kend = min(K, kbeg + klen);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int NA = BM * BK / NTHREADS, NB = BN * BK / NTHREADS;
  float ra[NA], rb[NB];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      int e = tid + i * NTHREADS;
      int mm = A_KFAST ? e / BK : e % BM, kk = A_KFAST ? e % BK : e / BM;
      ra[i] = (k0 + kk < kend) ? la(m0 + mm, k0 + kk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int e = tid + i * NTHREADS;
      int nn = B_NFAST ? e % BN : e / BK, kk = B_NFAST ? e / BN : e % BK;
      rb[i] = (k0 + kk < kend) ? lb(k0 + kk, n0 + nn) : 0.f;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      int e = tid + i * NTHREADS;
      int mm = A_KFAST ? e / BK : e % BM, kk = A_KFAST ? e % BK : e / BM;
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int e = tid + i * NTHREADS;
      int nn = B_NFAST ? e % BN : e / BK, kk = B_NFAST ? e / BN : e % BK;
      Bs[kk][nn] = rb[i];
    }
  };

  if (kbeg < kend) load_tile(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    store_tile();
    __syncthreads();
    if (k0 + BK < kend) load_tile(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < N) epi(m, n, acc[i][j], blockIdx.z);
    }
  }
}

// dW[m][n] (+)= sum_z ws[z][m][n]
__global__ void reduce_partials_kernel(const float* __restrict__ ws, float* __restrict__ out, int ld_out, int M, int N,
                                       int splits, int accumulate) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(size_t)z * M * N + idx];
  int m = idx / N, n = idx % N;
  float* o = out + (size_t)m * ld_out + n;
  *o = accumulate ? (*o + s) : s;
}

// column sums of dG (M rows, N columns) over row ranges: part[z][n]
__global__ void colsum_partial_kernel(const float* __restrict__ dG, int ldg, int M, int N, int rows_per, float* part) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int wy = threadIdx.x >> 5;  // 8 warps walk rows
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + wy; r < r1; r += 8) s += dG[(size_t)r * ldg + n];
  __shared__ float sm[8][33];
  sm[wy][threadIdx.x & 31] = s;
  __syncthreads();
  if (wy == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x & 31];
    part[(size_t)blockIdx.y * N + n] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ db, int N, int splits,
                                    int accumulate) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * N + n];
  db[n] = accumulate ? (db[n] + s) : s;
}

int wgrad_splits(int M, int N, int K) {
  const int tiles = tg_ceil_div(N, BM) * tg_ceil_div(K, BN);
  int s = (2 * tg_num_sms() + tiles - 1) / tiles;
  int maxs = tg_ceil_div(M, 4 * BK);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}
constexpr int COLSUM_SPLITS = 128;

}  // namespace

int tg_gemm_nt_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                    int ldc, int M, int N, int K, int accumulate) {
  TG_REQUIRE(A && W && C, TG_ERR_ARG, "gemm_nt: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, TG_ERR_SHAPE,
             "gemm_nt: bad shape M=%d N=%d K=%d lda=%d ldw=%d ldc=%d", M, N, K, lda, ldw, ldc);
  dim3 grid(tg_ceil_div(N, BN), tg_ceil_div(M, BM), 1);
  sgemm_kernel<true, false><<<grid, NTHREADS, 0, st>>>(RowMajorMK{A, lda, M, K}, RowMajorKN_T{W, ldw, K, N},
                                                       EpiStore{C, ldc, bias, accumulate}, M, N, K, K);
  return tg_check_launch("gemm_nt");
}

int tg_gemm_nn_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N,
                    int K, int accumulate) {
  TG_REQUIRE(A && W && C, TG_ERR_ARG, "gemm_nn: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= N && ldc >= N, TG_ERR_SHAPE,
             "gemm_nn: bad shape M=%d N=%d K=%d lda=%d ldw=%d ldc=%d", M, N, K, lda, ldw, ldc);
  dim3 grid(tg_ceil_div(N, BN), tg_ceil_div(M, BM), 1);
  sgemm_kernel<true, true><<<grid, NTHREADS, 0, st>>>(RowMajorMK{A, lda, M, K}, RowMajorKN{W, ldw, K, N},
                                                      EpiStore{C, ldc, nullptr, accumulate}, M, N, K, K);
  return tg_check_launch("gemm_nn");
}

size_t tg_wgrad_ws_bytes(int M, int N, int K) {
  return ((size_t)wgrad_splits(M, N, K) * N * K + (size_t)COLSUM_SPLITS * N) * sizeof(float);
}

int tg_wgrad_impl(cudaStream_t st, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw, float* db,
                  int M, int N, int K, int a_shift_T, int accumulate, float* ws, size_t ws_bytes) {
  TG_REQUIRE(dG && A && dW && ws, TG_ERR_ARG, "wgrad: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && ldg >= N && lda >= K && lddw >= K, TG_ERR_SHAPE,
             "wgrad: bad shape M=%d N=%d K=%d ldg=%d lda=%d lddw=%d", M, N, K, ldg, lda, lddw);
  TG_REQUIRE(ws_bytes >= tg_wgrad_ws_bytes(M, N, K), TG_ERR_ARG, "wgrad: workspace too small (%zu < %zu)", ws_bytes,
             tg_wgrad_ws_bytes(M, N, K));
  const int splits = wgrad_splits(M, N, K);
  int klen = tg_ceil_div(M, splits);
  klen = tg_ceil_div(klen, BK) * BK;
  dim3 grid(tg_ceil_div(K, BN), tg_ceil_div(N, BM), splits);
  if (a_shift_T > 0)
    sgemm_kernel<false, true><<<grid, NTHREADS, 0, st>>>(ColMajorMK{dG, ldg, N, M}, ShiftedKN{A, lda, M, K, a_shift_T},
                                                         EpiPartial{ws, N, K}, N, K, M, klen);
  else
    sgemm_kernel<false, true><<<grid, NTHREADS, 0, st>>>(ColMajorMK{dG, ldg, N, M}, RowMajorKN{A, lda, M, K},
                                                         EpiPartial{ws, N, K}, N, K, M, klen);
  int rc = tg_check_launch("wgrad");
  if (rc) return rc;
  reduce_partials_kernel<<<tg_ceil_div((long long)N * K, 256), 256, 0, st>>>(ws, dW, lddw, N, K, splits, accumulate);
  rc = tg_check_launch("wgrad_reduce");
  if (rc) return rc;
  if (db) {
    float* part = ws + (size_t)splits * N * K;
    const int rows_per = tg_ceil_div(M, COLSUM_SPLITS);
    colsum_partial_kernel<<<dim3(tg_ceil_div(N, 32), COLSUM_SPLITS), 256, 0, st>>>(dG, ldg, M, N, rows_per, part);
    rc = tg_check_launch("colsum");
    if (rc) return rc;
    colsum_final_kernel<<<tg_ceil_div(N, 128), 128, 0, st>>>(part, db, N, COLSUM_SPLITS, accumulate);
    rc = tg_check_launch("colsum_final");
  }
  return rc;
}

size_t tg_colsum_ws_bytes(int N) { return (size_t)COLSUM_SPLITS * N * sizeof(float); }

int tg_colsum_impl(cudaStream_t st, const float* X, int ld, int M, int N, float* out, int accumulate, float* ws,
                   size_t ws_bytes) {
  TG_REQUIRE(X && out && ws, TG_ERR_ARG, "colsum: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && ld >= N, TG_ERR_SHAPE, "colsum: bad shape M=%d N=%d ld=%d", M, N, ld);
  TG_REQUIRE(ws_bytes >= tg_colsum_ws_bytes(N), TG_ERR_ARG, "colsum: workspace too small");
  const int rows_per = tg_ceil_div(M, COLSUM_SPLITS);
  colsum_partial_kernel<<<dim3(tg_ceil_div(N, 32), COLSUM_SPLITS), 256, 0, st>>>(X, ld, M, N, rows_per, ws);
  int rc = tg_check_launch("colsum");
  if (rc) return rc;
  colsum_final_kernel<<<tg_ceil_div(N, 128), 128, 0, st>>>(ws, out, N, COLSUM_SPLITS, accumulate);
  return tg_check_launch("colsum_final");
}

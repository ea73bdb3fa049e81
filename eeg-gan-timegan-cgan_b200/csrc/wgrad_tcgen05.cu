// Weight-gradient contractions on the tensor cores:
//     dW[N,K] (+)= dG[M,N]^T A[M,K]          db[N] (+)= colsum(dG)           M = B*T is the contraction
// i.e. dW_ih = sum_t dGI_t^T x_t, dW_hh = sum_t dGH_t^T h_{t-1}, db = sum_t dG_t of SURVEY.md A.2 (autograd of the
// nn.GRU loop, train_timegan.py:140,159,219,267) and the head Linears' weight gradients.
//
// Both operands are stored with the contraction index (the (b,t) row) SLOWEST, so they enter tcgen05.mma as
// MN-major tiles: a TMA box of R rows x 32 columns written with the 128B-span / 32B-atom swizzle is exactly the
// canonical MN-major SWIZZLE_128B_BASE32B layout (the only one tcgen05 takes for 32-bit MN-major operands): 4 K-rows
// per 512-byte atom, two atoms = one kind::tf32 instruction (K = 8).
//   D[n, k] accumulates in TMEM as ceil(N/128) tiles of 128 lanes x (32 ceil(K/32)) columns.
// The M rows are split across one persistent CTA per SM; each CTA streams its row range through a TMA ring and
// finally writes its partial dW / db to a workspace that a small kernel reduces in a fixed order
// (deterministic: no floating-point atomics).
//   warp 0     TMA producer          warp 1   TMEM alloc + single-thread MMA issue
//   warps 2-5  epilogue (TMEM -> partial dW)
//   warps 6-9  tile fix-up: zero the h_{-1} rows of the shifted operand (a_shift_T), accumulate the bias-gradient
//              column sums, and in 3-pass mode split both operands into TF32-exact hi + residual lo
// passes = 1: TF32 operands (reduced-precision mode); passes = 3: 3xTF32 (fp32-parity mode).
#include "tc_common.cuh"
#include "kernels.h"

namespace {

constexpr int WG_THREADS = 320;
constexpr int WG_TAIL = 1024;  // barriers + tmem slot + slack for descriptor over-read of the last n-tile

struct WgParams {
  float* ws_dw;   // [splits][N][K]
  float* ws_db;   // [splits][N] or nullptr
  int M, N, K;
  int R;          // rows (contraction) per ring stage: 8..64
  int NCH, KCH;   // 32-column chunks of dG / A
  int MT;         // 128-row accumulator tiles = ceil(N/128)
  int nstage;
  int rows_per_cta;
  int shift_T;    // > 0: A row m is A[m-1], zero when m % shift_T == 0
  int tmem_cols;
};

template <int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmA, WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R, NS = p.nstage, NCH = p.NCH, KCH = p.KCH;
  const int chunk_bytes = R * 128;
  const int copy_bytes = (NCH + KCH) * chunk_bytes;            // one precision copy of a stage
  const int stage_bytes = (PASSES == 3 ? 2 : 1) * copy_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NS * stage_bytes);
  uint64_t* full = bars;            // [NS] TMA -> fix-up
  uint64_t* empty = bars + NS;      // [NS] MMA -> TMA
  uint64_t* ready = bars + 2 * NS;  // [NS] fix-up -> MMA
  uint64_t* acc_full = bars + 3 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* db_red = reinterpret_cast<float*>(smem);  // reused after the main loop: [16][NCH*32] floats

  const int NB = KCH * 32;  // UMMA N
  const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
  long long row_end = row_begin + p.rows_per_cta;
  if (row_end > p.M) row_end = p.M;
  const int steps = row_begin < row_end ? (int)((row_end - row_begin + R - 1) / R) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 128); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmA); }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < steps; ++it) {
        const int s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        mbar_wait_bounded(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], (uint32_t)copy_bytes);
        unsigned char* dst = smem + (size_t)s * stage_bytes;
        const int r0 = (int)(row_begin + (long long)it * R);
        for (int c = 0; c < NCH; ++c) tma_load_2d(dst + c * chunk_bytes, &tmG, &full[s], c * 32, r0);
        const int ra = r0 - (p.shift_T > 0 ? 1 : 0);   // row -1 is out of bounds -> zero fill
        for (int c = 0; c < KCH; ++c) tma_load_2d(dst + (NCH + c) * chunk_bytes, &tmA, &full[s], c * 32, ra);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, NB, 1, 1);
      uint32_t started = 0;  // bit mt set once accumulator tile mt has received its first MMA
      for (int it = 0; it < steps; ++it) {
        const int s = it % NS;
        const uint32_t ph = (uint32_t)((it / NS) & 1);
        mbar_wait_bounded(&ready[s], ph);
        tc_fence_after();
        const uint32_t g_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = g_hi + (uint32_t)(NCH * chunk_bytes);
        const uint32_t g_lo = g_hi + (uint32_t)copy_bytes, a_lo = a_hi + (uint32_t)copy_bytes;
        for (int i = 0; i < R / 8; ++i) {
          const uint32_t ko = (uint32_t)(i * 1024);
          const uint64_t db_hi = umma_desc_sw128_base32(a_hi + ko, (uint32_t)chunk_bytes, 512);
          const uint64_t db_lo = umma_desc_sw128_base32(a_lo + ko, (uint32_t)chunk_bytes, 512);
          for (int mt = 0; mt < p.MT; ++mt) {
            const uint32_t go = (uint32_t)(mt * 4 * chunk_bytes) + ko;
            const uint32_t d_addr = tmem_base + (uint32_t)(mt * NB);
            const uint64_t da_hi = umma_desc_sw128_base32(g_hi + go, (uint32_t)chunk_bytes, 512);
            umma_tf32(d_addr, da_hi, db_hi, idesc, (started >> mt) & 1u);
            started |= 1u << mt;
            if (PASSES == 3) {
              const uint64_t da_lo = umma_desc_sw128_base32(g_lo + go, (uint32_t)chunk_bytes, 512);
              umma_tf32(d_addr, da_hi, db_lo, idesc, 1);
              umma_tf32(d_addr, da_lo, db_hi, idesc, 1);
            }
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> partial dW (warps 2..5 -> lane quarters 2,3,0,1) ===============
    const int quarter = warp & 3;
    mbar_wait_bounded(acc_full, 0);
    tc_fence_after();
    float* out = p.ws_dw + (size_t)blockIdx.x * p.N * p.K;
    for (int mt = 0; mt < p.MT; ++mt) {
      const int n = mt * 128 + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * NB);
      for (int c0 = 0; c0 < NB; c0 += 16) {
        float v[16];
        if (steps > 0) {
          tmem_ld16(taddr + (uint32_t)c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (n < p.N) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.K) out[(size_t)n * p.K + c0 + i] = v[i];
        }
      }
    }
  } else {
    // ===================== fix-up / split / bias-gradient warps (6..9) =====================
    const int t = threadIdx.x - 192;              // 0..127
    const int u = t & 7, rsub = t >> 3;           // 16-byte unit within a 128-byte row, row residue mod 16
    // logical 4-column group this thread always sees: the 32-byte atom index is XORed with (row % 4)
    const int cu = ((((u >> 1) ^ (rsub & 3)) << 1) | (u & 1));
    float4 colsum[16];                            // per dG chunk (NCH <= 16)
#pragma unroll
    for (int c = 0; c < 16; ++c) colsum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < steps; ++it) {
      const int s = it % NS;
      const uint32_t ph = (uint32_t)((it / NS) & 1);
      mbar_wait_bounded(&full[s], ph);
      unsigned char* base = smem + (size_t)s * stage_bytes;
      const long long r0 = row_begin + (long long)it * R;
      // (1) rows of the shifted operand that would pair dG[m] with the previous sequence's last state
      if (p.shift_T > 0) {
        for (int r = rsub; r < R; r += 16) {
          if ((r0 + r) % p.shift_T == 0) {
            for (int c = 0; c < KCH; ++c)
              reinterpret_cast<float4*>(base + (NCH + c) * chunk_bytes + r * 128)[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      // (2) column sums of dG (+ hi/lo split of dG in 3-pass mode)
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c < NCH) {
          for (int r = rsub; r < R; r += 16) {
            float4* ptr = reinterpret_cast<float4*>(base + c * chunk_bytes + r * 128) + u;
            const float4 a = *ptr;
            colsum[c].x += a.x; colsum[c].y += a.y; colsum[c].z += a.z; colsum[c].w += a.w;
            if (PASSES == 3) {
              float4 h, l;
              tf32_split(a.x, h.x, l.x); tf32_split(a.y, h.y, l.y); tf32_split(a.z, h.z, l.z); tf32_split(a.w, h.w, l.w);
              *ptr = h;
              *(reinterpret_cast<float4*>(base + copy_bytes + c * chunk_bytes + r * 128) + u) = l;
            }
          }
        }
      }
      // (3) hi/lo split of the A operand
      if (PASSES == 3) {
        for (int c = 0; c < KCH; ++c)
          for (int r = rsub; r < R; r += 16) {
            float4* ptr = reinterpret_cast<float4*>(base + (NCH + c) * chunk_bytes + r * 128) + u;
            const float4 a = *ptr;
            float4 h, l;
            tf32_split(a.x, h.x, l.x); tf32_split(a.y, h.y, l.y); tf32_split(a.z, h.z, l.z); tf32_split(a.w, h.w, l.w);
            *ptr = h;
            *(reinterpret_cast<float4*>(base + copy_bytes + (NCH + c) * chunk_bytes + r * 128) + u) = l;
          }
      }
      fence_async_smem();
      mbar_arrive(&ready[s]);
    }
    // per-CTA bias-gradient partial: 16 row-residue classes per column -> fixed-order sum
    if (p.ws_db) {
      // the ring is idle once the MMA warp has consumed the last stage; wait for that before reusing smem
      mbar_wait_bounded(acc_full, 0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < NCH) *reinterpret_cast<float4*>(db_red + (size_t)rsub * (NCH * 32) + c * 32 + cu * 4) = colsum[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int n = t; n < p.N; n += 128) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) sum += db_red[(size_t)r * (NCH * 32) + n];
        p.ws_db[(size_t)blockIdx.x * p.N + n] = sum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// dW[n][k] (+)= sum_z ws_dw[z][n][k] ;  db[n] (+)= sum_z ws_db[z][n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws_dw, const float* __restrict__ ws_db,
                                    float* __restrict__ dW, int lddw, float* __restrict__ db, int N, int K, int splits,
                                    int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk = N * K;
  if (idx < nk) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws_dw[(size_t)z * nk + idx];
    float* o = dW + (size_t)(idx / K) * lddw + (idx % K);
    *o = accumulate ? (*o + s) : s;
  } else if (db && idx < nk + N) {
    const int n = idx - nk;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws_db[(size_t)z * N + n];
    db[n] = accumulate ? (db[n] + s) : s;
  }
}

int pow2_cols(int x) {
  int p = 32;
  while (p < x) p <<= 1;
  return p;
}

int wg_splits(int M) {
  int s = tg_num_sms();
  const int max_s = (M + 255) / 256;   // at least 256 rows per CTA
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

}  // namespace

size_t tg_wgrad_tc_ws_bytes(int M, int N, int K) { return (size_t)wg_splits(M) * ((size_t)N * K + N) * sizeof(float); }

int tg_wgrad_tc_impl(cudaStream_t st, const float* dG, int ldg, const float* A, int lda, float* dW, int lddw,
                     float* db, int M, int N, int K, int a_shift_T, int accumulate, float* ws, size_t ws_bytes,
                     int passes) {
  TG_REQUIRE(dG && A && dW && ws, TG_ERR_ARG, "wgrad_tc: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && ldg >= N && lda >= K && lddw >= K, TG_ERR_SHAPE, "wgrad_tc: bad shape");
  const int NCH = (N + 31) / 32, KCH = (K + 31) / 32, MT = (N + 127) / 128;
  const bool ok = (ldg % 4 == 0) && (lda % 4 == 0) && tg_aligned16(dG) && tg_aligned16(A) && NCH <= 16 && KCH <= 8 &&
                  MT * KCH * 32 <= 512 && M >= 64;
  if (!ok) { tg_set_error("wgrad_tc: shape/alignment not supported by the tensor-core tile"); return TG_ERR_UNSUPPORTED; }
  TG_REQUIRE(ws_bytes >= tg_wgrad_tc_ws_bytes(M, N, K), TG_ERR_ARG, "wgrad_tc: workspace too small");
  const int copies = passes == 3 ? 2 : 1;
  int R = 64, nstage = 0;
  for (; R >= 8; R >>= 1) {
    const int stage_bytes = copies * (NCH + KCH) * R * 128;
    nstage = (tg_gemm_smem_budget() - WG_TAIL - 4 * R * 128) / stage_bytes;
    if (nstage >= 3) break;
  }
  if (R < 8 || nstage < 3) { tg_set_error("wgrad_tc: tile does not fit shared memory"); return TG_ERR_UNSUPPORTED; }
  if (nstage > 6) nstage = 6;
  const int stage_bytes = copies * (NCH + KCH) * R * 128;
  size_t smem = (size_t)nstage * stage_bytes + WG_TAIL + 4 * (size_t)R * 128;
  if (smem < (size_t)16 * NCH * 32 * 4 + WG_TAIL) smem = (size_t)16 * NCH * 32 * 4 + WG_TAIL;

  alignas(64) CUtensorMap tmG, tmA;
  if (tg_make_map_2d(&tmG, dG, M, N, ldg, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (tg_make_map_2d(&tmA, A, M, K, lda, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;

  const int splits = wg_splits(M);
  int rows_per = (M + splits - 1) / splits;
  rows_per = (rows_per + R - 1) / R * R;
  float* ws_dw = ws;
  float* ws_db = db ? ws + (size_t)splits * N * K : nullptr;
  WgParams p{ws_dw, ws_db, M, N, K, R, NCH, KCH, MT, nstage, rows_per, a_shift_T, pow2_cols(MT * KCH * 32)};
  if (passes == 3) {
    TG_OPT_IN_SMEM(tc_wgrad_kernel<3>, "wgrad_tc");
    tc_wgrad_kernel<3><<<splits, WG_THREADS, smem, st>>>(tmG, tmA, p);
  } else {
    TG_OPT_IN_SMEM(tc_wgrad_kernel<1>, "wgrad_tc");
    tc_wgrad_kernel<1><<<splits, WG_THREADS, smem, st>>>(tmG, tmA, p);
  }
  int rc = tg_check_launch("wgrad_tc");
  if (rc) return rc;
  const int total = N * K + (db ? N : 0);
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(ws_dw, ws_db, dW, lddw, db, N, K, splits, accumulate);
  return tg_check_launch("wgrad_reduce");
}

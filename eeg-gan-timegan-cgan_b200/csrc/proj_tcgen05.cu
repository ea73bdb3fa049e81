// Time-batched projections on the 5th-generation tensor cores (north_star kernel (3)):
//     C[M,N] (+)= A[M,K] W[N,K]^T + bias[N]          M = B*T (10^5..10^6), N = 3H or a head width, K = I
// replacing `params.linear_ih(input)` inside at::gru (timegan_model.py:33), the head Linears (tm:53,66,79) and,
// with W passed transposed, dX = dGI W_ih.
//
// Persistent warp-specialised kernel, one CTA per SM, fp32 data read straight from HBM by TMA:
//   warp 0      TMA producer: W tile once (resident for the CTA's lifetime), then a ring of 128-row A tiles
//               (2-D tensor maps, 128-byte swizzle, K padded to 32-float blocks by TMA zero fill)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, kind::tf32, fp32 accumulators in TMEM
//               (double-buffered: 2 x NT columns), tcgen05.commit releases smem stages / publishes accumulators
//   warps 2-5   epilogue: tcgen05.ld (one accumulator row per thread, 32 columns per instruction, the next chunk's load
//               in flight while this one is stored) -> + bias (staged in shared memory once) -> 128B-swizzled 32x32
//               staging tile -> TMA store (cp.async.bulk.tensor), double-buffered per warp: no thread ever issues a
//               global store, and rows / columns beyond M / N are clipped by the tensor map.
//               (accumulate mode, unused by the training step, keeps the read-modify-write path through the LSU)
//   warps 6-9   (3-pass mode only) split each landed tile in place into TF32-exact hi and residual lo parts
// Two precisions:
//   passes = 1  plain TF32 operands (10-bit mantissa; the "bf16-class" 2e-2 projection mode)
//   passes = 3  3xTF32: A_hi W_hi + A_hi W_lo + A_lo W_hi, ~2^-20 relative per product -> fp32-parity mode
// The kernel is HBM-bound by design (intensity ~ 2NK/(4(K+N)) flop/B): the tensor pipe is there so that the
// arithmetic disappears behind the (M x (K+N) x 4 byte) stream.
#include "tc_common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128;           // rows per tile = UMMA M
constexpr int KBLK = 32;          // fp32 per 128-byte swizzle row
constexpr int A_BLK_BYTES = BM * 128;
constexpr int NUM_THREADS_1P = 192, NUM_THREADS_3P = 320;
constexpr int STG_BYTES = 2 * 4096;                  // per epilogue warp: two 32-row x 128-byte staging tiles
constexpr int TAIL_BYTES = 2048 + 4 * STG_BYTES;     // barriers (512) + bias (<= 256 floats) + pad | epilogue staging

struct TcParams {
  float* C;
  const float* bias;
  int ldc, M, N, K;
  int NT;         // accumulator width of this CTA's n-tile (multiple of 16, <= 256)
  int KB;         // K blocks of 32
  int nstage;
  int accumulate;
  int tmem_cols;  // power of two >= 2*NT
};

template <int PASSES>
__global__ void __launch_bounds__(PASSES == 3 ? NUM_THREADS_3P : NUM_THREADS_1P, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmC, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT, KB = p.KB, NS = p.nstage;
  const int w_blk_bytes = NT * 128;
  const int w_bytes = KB * w_blk_bytes;          // one copy (hi)
  const int a_bytes = A_BLK_BYTES;               // one copy (hi) of one K block: the ring streams K blocks
  const int w_total = (((PASSES == 3 ? 2 : 1) * w_bytes) + 1023) / 1024 * 1024;
  const int stage_bytes = (PASSES == 3 ? 2 : 1) * a_bytes;
  unsigned char* w_hi = smem;
  unsigned char* w_lo = smem + w_bytes;
  unsigned char* a_base = smem + w_total;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + (size_t)NS * stage_bytes);
  uint64_t* full = bars;                 // [NS]  TMA -> (splitter | MMA)
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* split = bars + 2 * NS;       // [NS]  splitter -> MMA (3-pass)
  uint64_t* w_full = bars + 3 * NS;      // TMA -> (splitter | MMA)
  uint64_t* w_ready = w_full + 1;        // splitter -> MMA (3-pass)
  uint64_t* acc_full = w_full + 2;       // [2] MMA -> epilogue
  uint64_t* acc_empty = w_full + 4;      // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 6);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);     // [NT] (zeros if none)
  unsigned char* stg_all = reinterpret_cast<unsigned char*>(bars) + 2048;                     // 4 x STG_BYTES, 1 KB aligned

  const int n0 = blockIdx.y * NT;
  const int num_tiles = (p.M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&split[s], 128); }
    mbar_init(w_full, 1);
    mbar_init(w_ready, 128);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmC); }
  for (int i = threadIdx.x; i < NT; i += blockDim.x)
    bias_s[i] = (p.bias && blockIdx.y * NT + i < p.N) ? p.bias[blockIdx.y * NT + i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)w_bytes);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_hi + kb * w_blk_bytes, &tmW, w_full, kb * KBLK, n0);
      int kc = 0;  // K blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (uint32_t)((kc / NS) & 1);
          mbar_wait_bounded(&empty[s], ph ^ 1u);
          mbar_expect_tx(&full[s], (uint32_t)a_bytes);
          tma_load_2d(a_base + (size_t)s * stage_bytes, &tmA, &full[s], kb * KBLK, tile * BM);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(BM, NT, 0, 0);
      mbar_wait_bounded(PASSES == 3 ? w_ready : w_full, 0);
      tc_fence_after();
      int it = 0, kc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        mbar_wait_bounded(&acc_empty[as], aph ^ 1u);
        const uint32_t d_addr = tmem_base + (uint32_t)(as * NT);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (uint32_t)((kc / NS) & 1);
          mbar_wait_bounded(PASSES == 3 ? &split[s] : &full[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(a_base + (size_t)s * stage_bytes);
          const uint32_t a_lo = a_hi + (uint32_t)a_bytes;
#pragma unroll
          for (int k = 0; k < KBLK / 8; ++k) {
            if (kb * KBLK + k * 8 >= p.K) break;
            const uint32_t wo = (uint32_t)(kb * w_blk_bytes + k * 32);
            const uint64_t da_hi = umma_desc_sw128(a_hi + k * 32, 16, 1024);
            const uint64_t dw_hi = umma_desc_sw128(smem_u32(w_hi) + wo, 16, 1024);
            umma_tf32(d_addr, da_hi, dw_hi, idesc, acc);
            acc = 1;
            if (PASSES == 3) {
              const uint64_t da_lo = umma_desc_sw128(a_lo + k * 32, 16, 1024);
              const uint64_t dw_lo = umma_desc_sw128(smem_u32(w_lo) + wo, 16, 1024);
              umma_tf32(d_addr, da_hi, dw_lo, idesc, 1);
              umma_tf32(d_addr, da_lo, dw_hi, idesc, 1);
            }
          }
          umma_commit(&empty[s]);     // smem stage reusable once these MMAs have read it
        }
        umma_commit(&acc_full[as]);   // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    unsigned char* stg = stg_all + (warp - 2) * STG_BYTES;
    int it = 0;
    uint32_t nstore = 0;     // TMA stores issued by this warp (staging buffer = nstore & 1)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      mbar_wait_bounded(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * NT);
      const long long row0 = (long long)tile * BM + quarter * 32;   // first row of this warp's 32-row slab
      if (!p.accumulate) {
        // ---- TMA-store path ----
        uint32_t ra[32];
        tmem_ld32_issue(taddr, ra);
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]);
          if (c0 + 32 < NT) tmem_ld32_issue(taddr + (uint32_t)(c0 + 32), ra);   // next chunk in flight during the store
          unsigned char* buf = stg + (nstore & 1u) * 4096;
          if (lane == 0) bulk_wait_read<1>();     // the store that last read this buffer (two stores ago) is done
          __syncwarp();
          const uint32_t rowa = smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);   // broadcast read
            const uint32_t dst = rowa + (uint32_t)((j ^ (lane & 7)) << 4);             // 128-byte swizzle
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v[4 * j] + bv.x),
                         "f"(v[4 * j + 1] + bv.y), "f"(v[4 * j + 2] + bv.z), "f"(v[4 * j + 3] + bv.w)
                         : "memory");
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M && n0 + c0 < p.N) {
            tma_store_2d(&tmC, buf, n0 + c0, (int)row0);
            bulk_commit();
          } else if (lane == 0) {
            bulk_commit();                        // keep the group count in step with nstore
          }
          ++nstore;
        }
      } else {
        // ---- read-modify-write path (C += ...): staging tile [32 rows][36 floats], coalesced 128-byte row segments ----
        float* stf = reinterpret_cast<float*>(stg);
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        const int c4 = lane & 7, rsub = lane >> 3;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<float(*)[16]>(&v[0]));
          if (c0 + 16 < NT) tmem_ld16(taddr + (uint32_t)(c0 + 16), *reinterpret_cast<float(*)[16]>(&v[16]));
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            *reinterpret_cast<float4*>(stf + lane * 36 + q4 * 4) = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
          __syncwarp();
          const int n = n0 + c0 + c4 * 4;
          const bool col_ok = (c0 + c4 * 4 < NT) && (n + 3 < p.N);
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok) bv = *reinterpret_cast<const float4*>(bias_s + c0 + c4 * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = rsub + 4 * i;
            const long long row = row0 + r;
            if (col_ok && row < p.M) {
              float4 o = *reinterpret_cast<const float4*>(stf + r * 36 + c4 * 4);
              float4* dst = reinterpret_cast<float4*>(p.C + row * p.ldc + n);
              const float4 old = *dst;
              o.x += bv.x + old.x; o.y += bv.y + old.y; o.z += bv.z + old.z; o.w += bv.w + old.w;
              *dst = o;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    if (lane == 0) bulk_wait_all<0>();       // every TMA store of this warp has landed before the CTA exits
  } else {
    // ===================== TF32 hi/lo splitter (3-pass mode, warps 6..9) =====================
    if (PASSES == 3) {
      const int t = threadIdx.x - 192;  // 0..127
      mbar_wait_bounded(w_full, 0);
      for (int i = t; i < w_bytes / 16; i += 128) {
        float4 a = reinterpret_cast<float4*>(w_hi)[i], h, l;
        tf32_split(a.x, h.x, l.x); tf32_split(a.y, h.y, l.y); tf32_split(a.z, h.z, l.z); tf32_split(a.w, h.w, l.w);
        reinterpret_cast<float4*>(w_hi)[i] = h;
        reinterpret_cast<float4*>(w_lo)[i] = l;
      }
      fence_async_smem();
      mbar_arrive(w_ready);
      int kc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (uint32_t)((kc / NS) & 1);
          mbar_wait_bounded(&full[s], ph);
          float4* hi = reinterpret_cast<float4*>(a_base + (size_t)s * stage_bytes);
          float4* lo = reinterpret_cast<float4*>(a_base + (size_t)s * stage_bytes + a_bytes);
          // A_BLK_BYTES / 16 / 128 = 8 units per thread: all loads first, then split + stores
          float4 a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = hi[t + j * 128];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 h, l;
            tf32_split(a[j].x, h.x, l.x); tf32_split(a[j].y, h.y, l.y); tf32_split(a[j].z, h.z, l.z); tf32_split(a[j].w, h.w, l.w);
            hi[t + j * 128] = h;
            lo[t + j * 128] = l;
          }
          fence_async_smem();
          mbar_arrive(&split[s]);
        }
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

int pow2_at_least(int x) {
  int p = 32;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

// passes: 1 = TF32, 3 = 3xTF32.  Returns TG_ERR_UNSUPPORTED for shapes/alignments the tile cannot take.
int tg_proj_tc_impl(cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias, float* C,
                    int ldc, int M, int N, int K, int accumulate, int passes) {
  TG_REQUIRE(A && W && C, TG_ERR_ARG, "proj_tc: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, TG_ERR_SHAPE,
             "proj_tc: bad shape M=%d N=%d K=%d lda=%d ldw=%d ldc=%d", M, N, K, lda, ldw, ldc);
  const bool ok = (K % 4 == 0) && (N % 4 == 0) && (lda % 4 == 0) && (ldw % 4 == 0) && (ldc % 4 == 0) && tg_aligned16(A) &&
                  tg_aligned16(W) && tg_aligned16(C) && (!bias || tg_aligned16(bias)) && K <= 512 && M >= BM;
  if (!ok) { tg_set_error("proj_tc: shape/alignment not supported by the tensor-core tile"); return TG_ERR_UNSUPPORTED; }
  const int KB = (K + KBLK - 1) / KBLK;
  const int n_pad = (N + 15) / 16 * 16;
  const int copies = (passes == 3) ? 2 : 1;
  const int stage_bytes = copies * A_BLK_BYTES;
  // widest n-tile (<= 256 columns) whose resident W copy leaves room for >= 3 ring stages (2 when K >= 256: the
  // resident W of a 32-column tile alone is 64 KB there).  With more than one n-tile NT is a multiple of 32, the width
  // of the epilogue's store boxes: a box must never reach into the neighbouring tile's columns.
  int n_tiles = (n_pad + 255) / 256, NT = 0, w_total = 0, nstage = 0;
  for (; n_tiles <= 24; ++n_tiles) {
    NT = (n_pad + n_tiles - 1) / n_tiles;
    NT = (n_tiles > 1) ? (NT + 31) / 32 * 32 : (NT + 15) / 16 * 16;
    w_total = ((copies * KB * NT * 128) + 1023) / 1024 * 1024;
    nstage = (tg_gemm_smem_budget() - 1024 - w_total - TAIL_BYTES) / stage_bytes;
    if (nstage >= 3 || (KB >= 8 && nstage >= 2)) break;
  }
  n_tiles = (n_pad + NT - 1) / NT;      // rounding NT up may have made the last tile(s) unnecessary
  if (nstage > 8) nstage = 8;
  if (nstage < 2 || (nstage < 3 && KB < 8)) { tg_set_error("proj_tc: tile does not fit shared memory (K=%d N=%d passes=%d)", K, N, passes); return TG_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)w_total + (size_t)nstage * stage_bytes + TAIL_BYTES;

  alignas(64) CUtensorMap tmA, tmW, tmC;
  if (tg_make_map_2d(&tmA, A, M, K, lda, KBLK, BM) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (tg_make_map_2d(&tmW, W, N, K, ldw, KBLK, NT) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (tg_make_map_2d(&tmC, C, M, N, ldc, 32, 32) != TG_OK) return TG_ERR_UNSUPPORTED;   // output: 32 x 32 store boxes

  // the epilogue reads the accumulator in 32-column chunks: the last chunk of buffer 1 may reach past 2*NT
  TcParams p{C, bias, ldc, M, N, K, NT, KB, nstage, accumulate, pow2_at_least(NT + (NT + 31) / 32 * 32)};
  const int num_tiles = (M + BM - 1) / BM;
  int gx = tg_num_sms() / n_tiles;
  if (gx > num_tiles) gx = num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_tiles);
  if (passes == 3) {
    TG_OPT_IN_SMEM(tc_gemm_tn_kernel<3>, "proj_tc");
    tc_gemm_tn_kernel<3><<<grid, NUM_THREADS_3P, smem, st>>>(tmA, tmW, tmC, p);
  } else {
    TG_OPT_IN_SMEM(tc_gemm_tn_kernel<1>, "proj_tc");
    tc_gemm_tn_kernel<1><<<grid, NUM_THREADS_1P, smem, st>>>(tmA, tmW, tmC, p);
  }
  return tg_check_launch("proj_tc");
}

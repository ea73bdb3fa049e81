// The "bf16 input projection" mode of BASELINE config c3 (north_star (3)): the time-batched GRU input projection
//     GI[M,N] = bf16(A[M,K]) bf16(W[N,K])^T + bias[N]        M = B*T, N = 3H, K = I
// with bf16 operands on the tensor pipe (tcgen05.mma kind::f16, fp32 accumulation in TMEM) and a **bf16 result**:
// the projection writes 2 bytes per gate pre-activation instead of 4 and the recurrent forward kernel reads 2
// (gru_fwd.cu, the GIB instantiations), so the (B,T,3H) tensor -- the largest one of a layer pass -- costs half the HBM
// traffic on both sides.  Replaces `params.linear_ih(input)` inside at::gru (timegan_model.py:33) like proj_tcgen05.cu.
//
// The activations stay fp32 in HBM (they are the fp32 outputs of the previous recurrent kernel); they are converted on
// the way through shared memory, so no bf16 copy of any activation is ever written to memory:
//   warp 0      TMA producer: W (already bf16, K-major, 64-element = 128-byte swizzled rows) once, then a ring of
//               128-row A stages; a stage holds 64 K-values as two fp32 landing boxes (32 floats = 128 B per row each)
//   warps 6-9   converter: fp32 boxes -> one bf16 operand tile [128 rows x 64 bf16] in the 128-byte-swizzle K-major
//               layout the UMMA descriptor expects (cvt.rn.bf16x2.f32, 8-byte stores)
//   warp 1      single-thread tcgen05.mma kind::f16 (M = 128, N = NT, K = 16 per instruction), double-buffered
//               accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld 64 columns -> + bias -> bf16x2 packs -> 128B-swizzled [32 rows x 64 bf16] staging
//               tile -> TMA store (bf16 tensor map; rows >= M and columns >= N are clipped by the map)
#include <cuda_bf16.h>
#include "tc_common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128;
constexpr int A_BOX_BYTES = BM * 128;                    // one fp32 landing box: 128 rows x 32 floats
constexpr int STAGE_BYTES = 3 * A_BOX_BYTES;             // two landing boxes + the bf16 operand tile
constexpr int NUM_THREADS = 320;
constexpr int STG_BYTES = 2 * 4096;                      // per epilogue warp: two [32 x 128 B] staging tiles
constexpr int TAIL_BYTES = 2048 + 4 * STG_BYTES;         // barriers (512) + bias (<= 320 floats) | staging

struct BfParams {
  const float* bias;
  int M, N, K;
  int NT;         // accumulator width of this CTA's n-tile (multiple of 16, <= 256)
  int KB2;        // K blocks of 64
  int nstage;
  int tmem_cols;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}

// instruction descriptor, kind::f16 with bf16 operands and fp32 accumulation, both operands K-major
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmC, BfParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT, KB2 = p.KB2, NS = p.nstage;
  const int w_blk_bytes = NT * 128;
  const int w_total = (KB2 * w_blk_bytes + 1023) / 1024 * 1024;
  unsigned char* w_bf = smem;
  unsigned char* a_base = smem + w_total;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_base + (size_t)NS * STAGE_BYTES);
  uint64_t* full = bars;                 // [NS]  TMA -> converter
  uint64_t* empty = bars + NS;           // [NS]  MMA -> TMA
  uint64_t* conv = bars + 2 * NS;        // [NS]  converter -> MMA
  uint64_t* w_full = bars + 3 * NS;      // TMA -> MMA
  uint64_t* acc_full = w_full + 1;       // [2] MMA -> epilogue
  uint64_t* acc_empty = w_full + 3;      // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 5);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 512);     // [NT + 64] (zero padded)
  unsigned char* stg_all = reinterpret_cast<unsigned char*>(bars) + 2048;

  const int n0 = blockIdx.y * NT;
  const int num_tiles = (p.M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 128); }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmC); }
  for (int i = threadIdx.x; i < NT + 64; i += blockDim.x)
    bias_s[i] = (p.bias && i < NT && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(KB2 * w_blk_bytes));
      for (int kb = 0; kb < KB2; ++kb) tma_load_2d(w_bf + kb * w_blk_bytes, &tmW, w_full, kb * 64, n0);
      int kc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < KB2; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (uint32_t)((kc / NS) & 1);
          mbar_wait_bounded(&empty[s], ph ^ 1u);
          const int nbox = (kb * 64 + 32 < p.K) ? 2 : 1;
          unsigned char* st = a_base + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full[s], (uint32_t)(nbox * A_BOX_BYTES));
          tma_load_2d(st, &tmA, &full[s], kb * 64, tile * BM);
          if (nbox == 2) tma_load_2d(st + A_BOX_BYTES, &tmA, &full[s], kb * 64 + 32, tile * BM);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, NT);
      mbar_wait_bounded(w_full, 0);
      tc_fence_after();
      int it = 0, kc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        mbar_wait_bounded(&acc_empty[as], aph ^ 1u);
        const uint32_t d_addr = tmem_base + (uint32_t)(as * NT);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB2; ++kb, ++kc) {
          const int s = kc % NS;
          const uint32_t ph = (uint32_t)((kc / NS) & 1);
          mbar_wait_bounded(&conv[s], ph);
          tc_fence_after();
          const uint32_t a_bf = smem_u32(a_base + (size_t)s * STAGE_BYTES + 2 * A_BOX_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                       // 16 bf16 = 32 bytes per instruction
            if (kb * 64 + k * 16 >= p.K) break;
            const uint64_t da = umma_desc_sw128(a_bf + k * 32, 16, 1024);
            const uint64_t dw = umma_desc_sw128(smem_u32(w_bf) + (uint32_t)(kb * w_blk_bytes + k * 32), 16, 1024);
            umma_f16(d_addr, da, dw, idesc, acc);
            acc = 1;
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int quarter = warp & 3;
    unsigned char* stg = stg_all + (warp - 2) * STG_BYTES;
    int it = 0;
    uint32_t nstore = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      mbar_wait_bounded(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * NT);
      const long long row0 = (long long)tile * BM + quarter * 32;
      for (int c0 = 0; c0 < NT; c0 += 64) {
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(taddr + (uint32_t)c0, ra);
        const bool second = c0 + 32 < NT;
        if (second) tmem_ld32_issue(taddr + (uint32_t)(c0 + 32), rb);
        tmem_ld_wait();
        unsigned char* buf = stg + (nstore & 1u) * 4096;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const uint32_t rowa = smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                // 16-byte chunk j = columns c0 + 8j .. c0 + 8j + 7
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * j);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + 8 * j + 4);
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t raw = (j < 4) ? ra[8 * j + i] : (second ? rb[8 * (j - 4) + i] : 0u);
            v[i] = __uint_as_float(raw);
          }
          const uint32_t o0 = pack_bf16x2(v[0] + b0.x, v[1] + b0.y), o1 = pack_bf16x2(v[2] + b0.z, v[3] + b0.w);
          const uint32_t o2 = pack_bf16x2(v[4] + b1.x, v[5] + b1.y), o3 = pack_bf16x2(v[6] + b1.z, v[7] + b1.w);
          const uint32_t dst = rowa + (uint32_t)((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < p.M && n0 + c0 < p.N) {
          tma_store_2d(&tmC, buf, n0 + c0, (int)row0);
          bulk_commit();
        } else if (lane == 0) {
          bulk_commit();
        }
        ++nstore;
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
    if (lane == 0) bulk_wait_all<0>();
  } else {
    // ===================== fp32 -> bf16 converter (warps 6..9) =====================
    const int t = threadIdx.x - 192;                 // 0..127
    // float4 unit i = t + 128 j of a landing box: row = i / 8 = t/8 + 16 j, physical 16-byte chunk = t % 8;
    // (row & 7) = (t/8) & 7 for every j, so the logical chunk c and both swizzled positions are per-thread constants
    const int r7 = (t >> 3) & 7;
    const int c = (t & 7) ^ r7;                      // logical chunk of the fp32 row: floats 4c .. 4c+3
    int kc = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB2; ++kb, ++kc) {
        const int s = kc % NS;
        const uint32_t ph = (uint32_t)((kc / NS) & 1);
        mbar_wait_bounded(&full[s], ph);
        const int nbox = (kb * 64 + 32 < p.K) ? 2 : 1;
        unsigned char* st = a_base + (size_t)s * STAGE_BYTES;
        const uint32_t dst0 = smem_u32(st + 2 * A_BOX_BYTES) + (uint32_t)((t >> 3) * 128 + (c & 1) * 8);
        for (int box = 0; box < nbox; ++box) {
          const float4* src = reinterpret_cast<const float4*>(st + box * A_BOX_BYTES);
          float4 a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = src[t + j * 128];
          const uint32_t chunk = (uint32_t)(((box * 4 + (c >> 1)) ^ r7) << 4);   // bf16 row: 16-byte chunk box*4 + c/2
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t lo = pack_bf16x2(a[j].x, a[j].y), hi = pack_bf16x2(a[j].z, a[j].w);
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst0 + (uint32_t)(j * 16 * 128) + chunk), "r"(lo), "r"(hi)
                         : "memory");
          }
        }
        fence_async_smem();
        mbar_arrive(&conv[s]);
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

// 2-D map over a row-major bf16 matrix, 128-byte swizzle (box_cols * 2 must be 128)
int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                  int box_rows) {
  tg_encode_tiled_fn enc = tg_get_encode_tiled();
  if (!enc) { tg_set_error("cuTensorMapEncodeTiled unavailable"); return TG_ERR_UNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { tg_set_error("cuTensorMapEncodeTiled (bf16) failed: CUresult %d", (int)r); return TG_ERR_ARG; }
  return TG_OK;
}

}  // namespace

// A fp32 (M,K) with leading dimension lda floats; W16 bf16 (N,K), ldw elements; C16 bf16 (M,N), ldc elements.
int tg_proj_bf16_impl(cudaStream_t st, const float* A, int lda, const void* W16, int ldw, const float* bias, void* C16,
                      int ldc, int M, int N, int K) {
  TG_REQUIRE(A && W16 && C16, TG_ERR_ARG, "proj_bf16: null pointer");
  TG_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N, TG_ERR_SHAPE,
             "proj_bf16: bad shape M=%d N=%d K=%d lda=%d ldw=%d ldc=%d", M, N, K, lda, ldw, ldc);
  const bool ok = (K % 4 == 0) && (N % 8 == 0) && (lda % 4 == 0) && (ldw % 8 == 0) && (ldc % 8 == 0) && tg_aligned16(A) &&
                  tg_aligned16(W16) && tg_aligned16(C16) && (!bias || tg_aligned16(bias)) && K <= 512 && M >= BM;
  if (!ok) { tg_set_error("proj_bf16: shape/alignment not supported by the tensor-core tile"); return TG_ERR_UNSUPPORTED; }
  const int KB2 = (K + 63) / 64;
  const int n_pad = (N + 15) / 16 * 16;
  // widest n-tile (<= 256 columns) whose resident bf16 W leaves room for >= 3 ring stages (2 when K >= 256); with more
  // than one n-tile NT is a multiple of 64, the width of the epilogue's store boxes
  int n_tiles = (n_pad + 255) / 256, NT = 0, w_total = 0, nstage = 0;
  for (; n_tiles <= 24; ++n_tiles) {
    NT = (n_pad + n_tiles - 1) / n_tiles;
    NT = (n_tiles > 1) ? (NT + 63) / 64 * 64 : (NT + 15) / 16 * 16;
    w_total = (KB2 * NT * 128 + 1023) / 1024 * 1024;
    nstage = (tg_gemm_smem_budget() - 1024 - w_total - TAIL_BYTES) / STAGE_BYTES;
    if (nstage >= 3 || (KB2 >= 4 && nstage >= 2)) break;
  }
  n_tiles = (n_pad + NT - 1) / NT;
  if (nstage > 4) nstage = 4;
  if (nstage < 2 || (nstage < 3 && KB2 < 4)) { tg_set_error("proj_bf16: tile does not fit shared memory (K=%d N=%d)", K, N); return TG_ERR_UNSUPPORTED; }
  const size_t smem = (size_t)w_total + (size_t)nstage * STAGE_BYTES + TAIL_BYTES;

  alignas(64) CUtensorMap tmA, tmW, tmC;
  if (tg_make_map_2d(&tmA, A, M, K, lda, 32, BM) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (make_map_bf16(&tmW, W16, N, K, ldw, 64, NT) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (make_map_bf16(&tmC, C16, M, N, ldc, 64, 32) != TG_OK) return TG_ERR_UNSUPPORTED;

  int tmem_cols = pow2_at_least(NT + (NT + 63) / 64 * 64);
  if (tmem_cols > 512) tmem_cols = 512;
  BfParams p{bias, M, N, K, NT, KB2, nstage, tmem_cols};
  const int num_tiles = (M + BM - 1) / BM;
  int gx = tg_num_sms() / n_tiles;
  if (gx > num_tiles) gx = num_tiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_tiles);
  TG_OPT_IN_SMEM(tc_gemm_bf16_kernel, "proj_bf16");
  tc_gemm_bf16_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmW, tmC, p);
  return tg_check_launch("proj_bf16");
}

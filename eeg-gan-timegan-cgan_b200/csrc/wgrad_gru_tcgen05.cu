// All weight gradients of one GRU layer in ONE tensor-core kernel (autograd of the nn.GRU loop):
//     dW_ih (3H x I)  (+)= sum_m dGI[m]^T x[m]            db_ih (+)= sum_m dGI[m]
//     dW_hh (3H x H)  (+)= sum_m dGH[m]^T h_{m-1}          db_hh (+)= sum_m dGH[m]
// with dGH = [dGI[:, :2H] | dq] (the BPTT kernel's two outputs) and h_{m-1} read from the layer output y with a
// one-row shift that restarts at every sequence (h_{-1} = 0).  The three contractions share operands, so the
// kernel streams dGI, dq, x and y ONCE (vs. 1.5x with three separate GEMMs) through one TMA ring:
//     A operand (MN-major):  [ dGI (3H cols) | dq (H cols) ]      -> accumulator rows (128-row TMEM tiles)
//     B operand (MN-major):  [ x (I cols)    | y shifted (H) ]    -> accumulator columns
// Needed blocks of D = A^T B:  dGI x -> dW_ih,  dGI[:, :2H] y -> dW_hh[:2H],  dq y -> dW_hh[2H:]
// (the two cross blocks dGI[:, 2H:] y and dq x are computed and dropped: the tensor pipe is not the limit).
// Same machinery as wgrad_tcgen05.cu: 32B-atom 128B swizzle for MN-major TF32, split-M over one CTA per SM,
// deterministic partial reduction, fix-up warps (h_{-1} rows, bias column sums, 3xTF32 split).  The lo parts of the
// 3xTF32 split live in THREE shared buffers (they are only needed while a stage's MMAs run), which leaves room for
// a 6-deep ring of raw stages -- the kernel is bound by bytes in flight, not by math.
// TMEM holds 512 accumulator columns: MT M-tiles x (I+H padded) columns.  When the whole [dGI | dq] operand does not fit
// (H = 128: 4 M-tiles x 256 columns), the layer is done in several launches, each owning a range of M-tiles (i.e. of
// dGI / dq columns) and streaming only those columns plus x and y; all launches write disjoint parts of one workspace.
// When [x | y] is wider than one MMA's 256 columns (H = 256: 512), the accumulator columns are split into ranges of
// <= 8 chunks as well: (M-tile range) x (column range) launches, each streaming its two operand slices.
#include "tc_common.cuh"
#include "kernels.h"

namespace {

constexpr int WL_NLO = 3;                     // lo-part buffers of the 3xTF32 split: fix-up of step i waits for the MMAs of i-3
constexpr int WL_NG = 3;                      // fix-up groups of 128 threads (chunk c belongs to group c % WL_NG)
constexpr int WL_FIX = 128 * WL_NG;
constexpr int WL_THREADS = 192 + WL_FIX;      // TMA, MMA, 4 epilogue warps, 4 * WL_NG fix-up warps
constexpr int WL_TAIL = 1024;
constexpr int WL_MAXCH = 16;   // chunks of the A operand (dGI + dq)

struct WlParams {
  float* ws;        // [splits][3H*I + 3H*H + GCH*32] : dW_ih partial, dW_hh partial, column-sum partial
  int M, I, H, T;
  int R, nstage, rows_per_cta;
  int g1ch, g2ch;   // 32-col chunks of dGI / dq (whole layer)
  int a1ch, a2ch;   // 32-col chunks of x (0 when x is absent) / y
  int MT, tmem_cols;
  int mt0;          // first M-tile (128 rows of [dGI | dq]^T) of this launch; it covers MT tiles
  int ab, acnt;     // first 32-column chunk of [x | y] of this launch and how many it covers
};

template <int PASSES>
__global__ void __launch_bounds__(WL_THREADS, 1)
tc_wgrad_layer_kernel(const __grid_constant__ CUtensorMap tmG1, const __grid_constant__ CUtensorMap tmG2,
                      const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2, WlParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.R, NS = p.nstage;
  const int GTOT = p.g1ch + p.g2ch;                       // chunks of the whole [dGI | dq] operand
  const int cbase = p.mt0 * 4;                            // first chunk of this launch
  const int GCH = min(GTOT, cbase + p.MT * 4) - cbase;    // chunks of this launch
  const int ACH = p.acnt;                                 // [x | y] chunks of this launch: global chunks ab .. ab+acnt-1
  const int chunk_bytes = R * 128;
  const int stage_bytes = (GCH + ACH) * chunk_bytes;
  unsigned char* lo_base = smem + (size_t)NS * stage_bytes;                    // [2][stage_bytes] (3-pass only)
  unsigned char* tail = lo_base + (PASSES == 3 ? WL_NLO * (size_t)stage_bytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;            // [NS] TMA -> fix-up
  uint64_t* empty = bars + NS;      // [NS] MMA -> TMA (and -> fix-up: lo buffer of two steps ago is free)
  uint64_t* ready = bars + 2 * NS;  // [NS] fix-up -> MMA
  uint64_t* acc_full = bars + 3 * NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* db_red = reinterpret_cast<float*>(smem);   // after the main loop: [16][GCH*32]

  const int NB = ACH * 32;
  const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
  long long row_end = row_begin + p.rows_per_cta;
  if (row_end > p.M) row_end = p.M;
  const int steps = row_begin < row_end ? (int)((row_end - row_begin + R - 1) / R) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], WL_FIX); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG1); tma_prefetch_desc(&tmG2); tma_prefetch_desc(&tmA2);
    if (p.a1ch) tma_prefetch_desc(&tmA1);
  }
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
        mbar_wait_bounded(&empty[s], (uint32_t)(((it / NS) & 1) ^ 1));
        mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        unsigned char* dst = smem + (size_t)s * stage_bytes;
        const int r0 = (int)(row_begin + (long long)it * R);
        int c = 0;
        for (; c < GCH; ++c) {
          const int cg = cbase + c;
          if (cg < p.g1ch) tma_load_2d(dst + c * chunk_bytes, &tmG1, &full[s], cg * 32, r0);
          else tma_load_2d(dst + c * chunk_bytes, &tmG2, &full[s], (cg - p.g1ch) * 32, r0);
        }
        for (int k = 0; k < ACH; ++k, ++c) {
          const int ca = p.ab + k;
          if (ca < p.a1ch) tma_load_2d(dst + c * chunk_bytes, &tmA1, &full[s], ca * 32, r0);
          else tma_load_2d(dst + c * chunk_bytes, &tmA2, &full[s], (ca - p.a1ch) * 32, r0 - 1);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, NB, 1, 1);
      uint32_t started = 0;
      for (int it = 0; it < steps; ++it) {
        const int s = it % NS;
        mbar_wait_bounded(&ready[s], (uint32_t)((it / NS) & 1));
        tc_fence_after();
        const uint32_t g_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = g_hi + (uint32_t)(GCH * chunk_bytes);
        const uint32_t g_lo = smem_u32(lo_base + (size_t)(it % WL_NLO) * stage_bytes);
        const uint32_t a_lo = g_lo + (uint32_t)(GCH * chunk_bytes);
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
    // ===================== epilogue: TMEM -> partial dW_ih / dW_hh =====================
    const int quarter = warp & 3;
    mbar_wait_bounded(acc_full, 0);
    tc_fence_after();
    const int I = p.I, H = p.H;
    float* w_ih = p.ws + (size_t)blockIdx.x * ((size_t)3 * H * I + (size_t)3 * H * H + GTOT * 32);
    float* w_hh = w_ih + (size_t)3 * H * I;
    const int ycol0 = p.a1ch * 32;     // first accumulator column of the y block
    for (int mt = 0; mt < p.MT; ++mt) {
      const int n = (p.mt0 + mt) * 128 + quarter * 32 + lane;     // logical row of [dGI | pad | dq]
      const int nq = n - p.g1ch * 32;                    // row inside dq (>= 0 for the dq block)
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * NB);
      for (int c0 = 0; c0 < NB; c0 += 16) {
        float v[16];
        if (steps > 0) {
          tmem_ld16(taddr + (uint32_t)c0, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = p.ab * 32 + c0 + i;                  // column of the whole [x | y] operand
          if (c < ycol0) {                                   // x block -> dW_ih
            if (n < 3 * H && c < I) w_ih[(size_t)n * I + c] = v[i];
          } else {                                           // y block -> dW_hh
            const int k = c - ycol0;
            if (k < H) {
              if (n < 2 * H) w_hh[(size_t)n * H + k] = v[i];
              else if (nq >= 0 && nq < H) w_hh[(size_t)(2 * H + nq) * H + k] = v[i];
            }
          }
        }
      }
    }
  } else {
    // ===================== fix-up / split / column-sum warps (6..) =====================
    // WL_NG groups of 128 threads share the chunks of every stage round-robin (the fix-up is the stage of the
    // pipeline both the producer and the MMA thread wait for); inside a group thread (u, rsub) owns the 16-byte
    // unit u of the rows r == rsub (mod 16)
    const int t = threadIdx.x - 192;
    const int grp = t >> 7, tt = t & 127;
    const int u = tt & 7, rsub = tt >> 3;
    const int cu = ((((u >> 1) ^ (rsub & 3)) << 1) | (u & 1));
    float4 colsum[WL_MAXCH / 2];
#pragma unroll
    for (int c = 0; c < WL_MAXCH / 2; ++c) colsum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int NCHK = GCH + ACH;
    // (first row of the stage) mod T, kept incrementally: a 64-bit modulo per row costs ~100 instructions
    int rem0 = (int)(row_begin % p.T);
    for (int it = 0; it < steps; ++it) {
      const int s = it % NS;
      mbar_wait_bounded(&full[s], (uint32_t)((it / NS) & 1));
      if (PASSES == 3 && it >= WL_NLO) {   // the lo buffer (it % WL_NLO) was last read by the MMAs of step it-WL_NLO
        const int j = it - WL_NLO;
        mbar_wait_bounded(&empty[j % NS], (uint32_t)((j / NS) & 1));
      }
      unsigned char* base = smem + (size_t)s * stage_bytes;
      unsigned char* lo = lo_base + (size_t)(it % WL_NLO) * stage_bytes;
      for (int r = rsub; r < R; r += 16) {
        int rr = rem0 + r;
        while (rr >= p.T) rr -= p.T;
        const bool boundary = rr == 0;
        const size_t off = (size_t)r * 128 + (size_t)u * 16;
        // three chunks per batch: the loads are issued together (the compiler cannot reorder them across the
        // in-place stores by itself), then column sums / h_{-1} zeroing / TF32 split, then the stores
        auto finish = [&](int c, int ci, float4 a) {
          if (c < GCH) { colsum[ci & 7].x += a.x; colsum[ci & 7].y += a.y; colsum[ci & 7].z += a.z; colsum[ci & 7].w += a.w; }
          const bool kill = boundary && c >= GCH && (c - GCH + p.ab) >= p.a1ch;     // y row of the previous sequence
          if (kill) a = make_float4(0.f, 0.f, 0.f, 0.f);
          float4* ptr = reinterpret_cast<float4*>(base + (size_t)c * chunk_bytes + off);
          if (PASSES == 3) {
            float4 h, l;
            tf32_split(a.x, h.x, l.x); tf32_split(a.y, h.y, l.y); tf32_split(a.z, h.z, l.z); tf32_split(a.w, h.w, l.w);
            *ptr = h;
            *reinterpret_cast<float4*>(lo + (size_t)c * chunk_bytes + off) = l;
          } else if (kill) {
            *ptr = a;
          }
        };
#pragma unroll
        for (int cb = 0; cb < (WL_MAXCH + 8 + WL_NG - 1) / WL_NG; cb += 3) {
          const int c0 = WL_NG * cb + grp, c1 = c0 + WL_NG, c2 = c0 + 2 * WL_NG;
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
          if (c0 < NCHK) a0 = *reinterpret_cast<const float4*>(base + (size_t)c0 * chunk_bytes + off);
          if (c1 < NCHK) a1 = *reinterpret_cast<const float4*>(base + (size_t)c1 * chunk_bytes + off);
          if (c2 < NCHK) a2 = *reinterpret_cast<const float4*>(base + (size_t)c2 * chunk_bytes + off);
          if (c0 < NCHK) finish(c0, cb, a0);
          if (c1 < NCHK) finish(c1, cb + 1, a1);
          if (c2 < NCHK) finish(c2, cb + 2, a2);
        }
      }
      fence_async_smem();
      mbar_arrive(&ready[s]);
      rem0 += R;
      while (rem0 >= p.T) rem0 -= p.T;
    }
    // per-CTA column-sum partial (bias gradients)
    mbar_wait_bounded(acc_full, 0);
    asm volatile("bar.sync 1, %0;" ::"n"(WL_FIX) : "memory");
#pragma unroll
    for (int ci = 0; ci < WL_MAXCH / 2; ++ci) {
      const int c = WL_NG * ci + grp;
      if (c < GCH) *reinterpret_cast<float4*>(db_red + (size_t)rsub * (GCH * 32) + c * 32 + cu * 4) = colsum[ci];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(WL_FIX) : "memory");
    float* cs = p.ws + (size_t)blockIdx.x * ((size_t)3 * p.H * p.I + (size_t)3 * p.H * p.H + GTOT * 32) +
                (size_t)3 * p.H * p.I + (size_t)3 * p.H * p.H + cbase * 32;
    for (int n = t; n < GCH * 32; n += WL_FIX) {
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) sum += db_red[(size_t)r * (GCH * 32) + n];
      cs[n] = sum;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// fixed-order reduction of the per-CTA partials into the four gradient tensors
__global__ void wgrad_layer_reduce_kernel(const float* __restrict__ ws, int splits, int I, int H, int g1ch, int gch,
                                          float* __restrict__ dW_ih, float* __restrict__ dW_hh,
                                          float* __restrict__ db_ih, float* __restrict__ db_hh, int has_x,
                                          int accumulate) {
  const size_t n_ih = (size_t)3 * H * I, n_hh = (size_t)3 * H * H;
  const size_t per = n_ih + n_hh + (size_t)gch * 32;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_ih) {
    if (!has_x) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * per + idx];
    dW_ih[idx] = accumulate ? dW_ih[idx] + s : s;
  } else if (idx < n_ih + n_hh) {
    const size_t j = idx - n_ih;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * per + idx];
    dW_hh[j] = accumulate ? dW_hh[j] + s : s;
  } else if (idx < n_ih + n_hh + (size_t)6 * H) {
    const int j = (int)(idx - n_ih - n_hh);           // 0..3H-1: db_ih, 3H..6H-1: db_hh
    const bool hh = j >= 3 * H;
    const int n = hh ? j - 3 * H : j;
    float* out = hh ? db_hh : db_ih;
    if (!out) return;
    // column of the concatenated [dGI | pad | dq] operand that feeds this bias entry
    const int col = (hh && n >= 2 * H) ? g1ch * 32 + (n - 2 * H) : n;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * per + n_ih + n_hh + col];
    out[n] = accumulate ? out[n] + s : s;
  }
}

int wl_splits(int M) {
  int s = tg_num_sms();
  const int cap = tg_wgrad_cta_cap();   // > 0: leave the other SMs to kernels running on other streams
  if (cap > 0 && cap < s) s = cap;
  const int max_s = (M + 255) / 256;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

int pow2c(int x) {
  int p = 32;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

size_t tg_wgrad_gru_ws_bytes(int M, int I, int H) {
  const int gch = (3 * H + 31) / 32 + (H + 31) / 32;
  return (size_t)wl_splits(M) * ((size_t)3 * H * I + (size_t)3 * H * H + (size_t)gch * 32) * sizeof(float);
}

// x may be NULL (then only dW_hh / biases are produced).  Returns TG_ERR_UNSUPPORTED when the fused tile cannot
// take the shape; the caller then uses three tg_wgrad calls.
int tg_wgrad_gru_tc_impl(cudaStream_t st, const float* dgi, const float* dq, const float* x, int ldx, const float* y,
                         float* dW_ih, float* dW_hh, float* db_ih, float* db_hh, int B, int T, int I, int H,
                         int accumulate, float* ws, size_t ws_bytes, int passes) {
  TG_REQUIRE(dgi && dq && y && dW_hh && ws, TG_ERR_ARG, "wgrad_gru: null pointer");
  TG_REQUIRE(!x || dW_ih, TG_ERR_ARG, "wgrad_gru: x given without dW_ih");
  TG_REQUIRE(B > 0 && T > 0 && H > 0 && (!x || (I > 0 && ldx >= I)), TG_ERR_SHAPE, "wgrad_gru: bad shape");
  const long long Mll = (long long)B * T;
  const int M = (int)Mll;
  const int g1ch = (3 * H + 31) / 32, g2ch = (H + 31) / 32, a1ch = x ? (I + 31) / 32 : 0, a2ch = (H + 31) / 32;
  const int GCH = g1ch + g2ch, ACH = a1ch + a2ch, MT = (GCH + 3) / 4;
  const bool ok = (H % 4 == 0) && (!x || (ldx % 4 == 0 && tg_aligned16(x))) && tg_aligned16(dgi) && tg_aligned16(dq) &&
                  tg_aligned16(y) && M >= 256 && Mll < (1ll << 31);
  if (!ok) { tg_set_error("wgrad_gru: shape/alignment not supported by the fused tensor-core tile"); return TG_ERR_UNSUPPORTED; }
  const int Iw = x ? I : 0;
  TG_REQUIRE(ws_bytes >= tg_wgrad_gru_ws_bytes(M, Iw, H), TG_ERR_ARG, "wgrad_gru: workspace too small");
  // accumulator columns per launch: one MMA takes N <= 256 = 8 chunks of [x | y]
  const int ACL = ACH < 8 ? ACH : 8;
  // M-tiles per launch: what 512 TMEM columns hold next to each other (all of them up to H = 96; two at H >= 128)
  int MTL = 512 / (ACL * 32);
  if (MTL > MT) MTL = MT;
  if (MTL * 4 > WL_MAXCH) MTL = WL_MAXCH / 4;
  const int gch_l = (MTL * 4 < GCH) ? MTL * 4 : GCH;      // widest chunk range of a launch
  int R = 32, nstage = 0;
  for (; R >= 8; R >>= 1) {
    const int stage_bytes = (gch_l + ACL) * R * 128;
    nstage = (tg_gemm_smem_budget() - WL_TAIL - 4 * R * 128) / stage_bytes - (passes == 3 ? WL_NLO : 0);
    if (nstage >= 4) break;
  }
  if (R < 8 || nstage < 3) { tg_set_error("wgrad_gru: tile does not fit shared memory"); return TG_ERR_UNSUPPORTED; }
  if (nstage > 8) nstage = 8;
  const int stage_bytes = (gch_l + ACL) * R * 128;
  size_t smem = (size_t)(nstage + (passes == 3 ? WL_NLO : 0)) * stage_bytes + WL_TAIL + 4 * (size_t)R * 128;
  if (smem < (size_t)16 * gch_l * 32 * 4 + WL_TAIL) smem = (size_t)16 * gch_l * 32 * 4 + WL_TAIL;

  alignas(64) CUtensorMap tmG1, tmG2, tmA1, tmA2;
  if (tg_make_map_2d(&tmG1, dgi, M, 3 * H, 3 * H, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (tg_make_map_2d(&tmG2, dq, M, H, H, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (tg_make_map_2d(&tmA2, y, M, H, H, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;
  if (x) {
    if (tg_make_map_2d(&tmA1, x, M, I, ldx, 32, R, true) != TG_OK) return TG_ERR_UNSUPPORTED;
  } else {
    tmA1 = tmA2;
  }
  const int splits = wl_splits(M);
  int rows_per = (M + splits - 1) / splits;
  rows_per = (rows_per + R - 1) / R * R;
  for (int mt0 = 0; mt0 < MT; mt0 += MTL)
  for (int ab = 0; ab < ACH; ab += ACL) {
    const int mtl = (MT - mt0 < MTL) ? MT - mt0 : MTL;
    const int acnt = (ACH - ab < ACL) ? ACH - ab : ACL;
    WlParams p{ws, M, Iw, H, T, R, nstage, rows_per, g1ch, g2ch, a1ch, a2ch, mtl, pow2c(mtl * acnt * 32), mt0, ab, acnt};
    if (passes == 3) {
      TG_OPT_IN_SMEM(tc_wgrad_layer_kernel<3>, "wgrad_gru");
      tc_wgrad_layer_kernel<3><<<splits, WL_THREADS, smem, st>>>(tmG1, tmG2, tmA1, tmA2, p);
    } else {
      TG_OPT_IN_SMEM(tc_wgrad_layer_kernel<1>, "wgrad_gru");
      tc_wgrad_layer_kernel<1><<<splits, WL_THREADS, smem, st>>>(tmG1, tmG2, tmA1, tmA2, p);
    }
    int rcl = tg_check_launch("wgrad_gru");
    if (rcl) return rcl;
  }
  int rc = tg_check_launch("wgrad_gru");
  if (rc) return rc;
  const size_t total = (size_t)3 * H * Iw + (size_t)3 * H * H + (size_t)6 * H;
  wgrad_layer_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, splits, Iw, H, g1ch, GCH, dW_ih, dW_hh,
                                                                           db_ih, db_hh, x ? 1 : 0, accumulate);
  return tg_check_launch("wgrad_gru_reduce");
}

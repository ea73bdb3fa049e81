// Persistent BPTT kernel for one GRU layer on sm_100a (north_star kernel (2)).
//
// Replaces autograd's backward of the per-timestep nn.GRU loop (loss.backward() at
// timeGAN/train_timegan.py:140,159,219,267 and the dX-only pass inside autograd.grad, tt:200) --
// math of SURVEY.md Appendix A.2.
//
// Mirror image of gru_fwd.cu: one CTA owns BT sequences for all T steps (t = T-1 .. 0), W_hh^T lives in
// registers (lane q of the G-lane group of hidden unit k holds column k of W_hh for the row-slice q of each
// gate), the carried dh lives in a register of the lane that owns (k, b) after the shuffle reduce-scatter,
// the per-step dGH vector is exchanged through a double-buffered shared-memory vector (one __syncthreads
// per step), saved r,z,n / q / h_{t-1} and dy stream in through the bulk-async ring and
// dGI = [dar,daz,dan] / dq = dan*r stream out in place.
// The weight-gradient contractions (K = B*T) and dX = dGI W_ih are separate GEMM kernels.
#include "chunk_pipe.cuh"
#include "kernels.h"

namespace {

struct BwdParams {
  const float* dy;   // (B,T,H), or (B,H) when dy_last
  const float* rzn;  // (B,T,3H) saved r,z,n
  const float* q;    // (B,T,H)  saved q
  const float* y;    // (B,T,H)  layer output (h_t); h_{t-1} is read with a -1 row shift
  const float* whh;  // (3H,H)
  float* dgi;        // (B,T,3H) out
  float* dq;         // (B,T,H)  out: dGH_n
  int B, T, H;
  int dy_last;
  int bulk;
};

constexpr int DG_PAD = 16;  // dGH rows are HP+16 floats apart (bank spread between the sequences of a lane group)

template <int HP, int G>
constexpr int bwd_min_blocks() { return (HP * G <= 128) ? 3 : ((HP * G <= 256) ? 2 : 1); }

// EXACT: H == HP at compile time (c2: 64, c3: 128): strides become immediates, the "real hidden unit" predicate
// disappears.
//
// Per-step dependency chain.  Everything that does not depend on the carried dh is taken OFF the chain: the
// saved r,z,n,q,h_{t-1},dy of step t-1 are fetched from the stage while step t's mat-vec runs, and folded into
// four factors  A = (1-z)(1-n^2), Bz = (h_{t-1}-n) z (1-z), C = q r (1-r), r  -- so that once the carry arrives
// the gate derivatives are  dh = dy + carry; dan = dh A; daz = dh Bz; dar = dan C; dq = dan r; cz = dh z :
// three dependent FP32 ops between the reduce-scatter and the shared-memory exchange of dGH.
template <int HP, int G, int BT, int TC, int NST, bool EXACT>
__global__ void __launch_bounds__(HP* G, (EXACT || bwd_min_blocks<HP, G>() == 1) ? bwd_min_blocks<HP, G>() : bwd_min_blocks<HP, G>() - 1) gru_bwd_kernel(BwdParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + DG_PAD;
  constexpr bool DUAL = (HP * G <= 128);         // second accumulator set only where the register budget allows
  static_assert(KS % 4 == 0, "slice must be float4 granular");
  static_assert(G == 2 || G == 4, "lane groups of 2 or 4");
  static_assert(BT < G || BT % G == 0, "BT must be < G or a multiple of G");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int k = tid / G, ql = tid % G;
  const int H = EXACT ? HP : p.H;
  const int T = p.T;
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, p.B - b0);

  float* dgs = reinterpret_cast<float*>(smem_raw);                     // [2][BT][3][HR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dgs + 2 * BT * 3 * HR);
  float* stages = reinterpret_cast<float*>(smem_raw + ((2 * BT * 3 * HR * 4 + NST * 8 + 127) / 128) * 128);

  ChunkPipe<4, BT, TC, NST> pipe;
  pipe.g[0] = const_cast<float*>(p.rzn); pipe.gst[0] = p.dgi; pipe.w[0] = 3 * H; pipe.mode[0] = TG_STRM_LOAD | TG_STRM_STORE; pipe.shift[0] = 0;
  pipe.g[1] = const_cast<float*>(p.q);   pipe.gst[1] = p.dq;  pipe.w[1] = H;     pipe.mode[1] = TG_STRM_LOAD | TG_STRM_STORE; pipe.shift[1] = 0;
  pipe.g[2] = const_cast<float*>(p.dy);  pipe.gst[2] = nullptr; pipe.w[2] = H;   pipe.mode[2] = p.dy_last ? 0 : TG_STRM_LOAD;  pipe.shift[2] = 0;
  pipe.g[3] = const_cast<float*>(p.y);   pipe.gst[3] = nullptr; pipe.w[3] = H;   pipe.mode[3] = TG_STRM_LOAD;                  pipe.shift[3] = -1;
  pipe.layout();
  pipe.stages = stages; pipe.full = bars;
  pipe.T = T; pipe.nb = nb; pipe.b0 = b0; pipe.NC = (T + TC - 1) / TC;
  pipe.reverse = true; pipe.bulk = p.bulk != 0;

  // W_hh^T slice: wt[g][m] = W_hh[g*H + jj][k], jj = (i*G+ql)*4+c
  float2 wt[3][KS / 2];   // float2 pairs for packed FFMA2
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int i = 0; i < KS / 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int jj = (i * G + ql) * 4 + c;
        const float v = (k < H && jj < H) ? p.whh[(size_t)(g * H + jj) * H + k] : 0.f;
        if (c & 1) wt[g][2 * i + (c >> 1)].y = v; else wt[g][2 * i + (c >> 1)].x = v;
      }
  for (int i = tid; i < 2 * BT * 3 * HR; i += HP * G) dgs[i] = 0.f;

  bool act[NOWN];
  int ob[NOWN];
  // carry[o]: dL/dh_t[k] flowing in from step t+1 for the sequence this lane owns (b = o*G + ql, or b = ql).
  // dy-last mode (discriminator head: the loss reads y[:, T-1] only): the (B,H) gradient IS the initial carry.
  float carry[NOWN];
  const bool use_dy = !p.dy_last;
#pragma unroll
  for (int o = 0; o < NOWN; ++o) {
    ob[o] = (BT < G) ? ql % BT : o * G + ql;     // surplus lanes of a group repeat a sibling's (identical) work
    act[o] = (EXACT && BT == 1) || ((EXACT || k < H) && (ob[o] < nb));
    carry[o] = (p.dy_last && act[o]) ? p.dy[(size_t)(b0 + ob[o]) * H + k] : 0.f;
  }
  pipe.start();
  __syncthreads();

  const uint32_t dgs_addr = smem_u32(dgs);
  const uint32_t h_step = 4u * (uint32_t)H, g_step = 12u * (uint32_t)H;
  const uint32_t lane_k = 16u * (uint32_t)ql;

  // factors of the step about to be processed + addresses of its rows
  float fA[NOWN], fB[NOWN], fC[NOWN], fr[NOWN], fz[NOWN], fn[NOWN], fdy[NOWN];
  uint32_t a_g[NOWN], a_q[NOWN], a_dy[NOWN], a_h[NOWN];

  auto fetch = [&](bool first_t) __attribute__((always_inline)) {
    // reads the rows a_* point at (step t), leaves the carry-independent factors in registers
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const float r = lds_f32(a_g[o]), z = lds_f32(a_g[o] + h_step), n = lds_f32(a_g[o] + 2u * h_step);
      const float qv = lds_f32(a_q[o]);
      float hp = lds_f32(a_h[o]);
      hp = first_t ? 0.f : hp;                 // h_{-1} = 0 (the shifted stream never loads row -1)
      float dyv = lds_f32(a_dy[o]);
      dyv = use_dy ? dyv : 0.f;
      const float omz = 1.f - z;
      fA[o] = omz * fmaf(-n, n, 1.f);
      fB[o] = (hp - n) * (z * omz);
      fC[o] = qv * (r * (1.f - r));
      fr[o] = r; fz[o] = z; fn[o] = n; fdy[o] = dyv;
    }
  };

  int par = 0;
  float cz[NOWN];
  auto step = [&]() __attribute__((always_inline)) {
    const uint32_t dg = dgs_addr + (uint32_t)(par * BT * 3 * HR) * 4u;
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const float dh = fdy[o] + carry[o];
      const float dan = dh * fA[o];
      const float daz = dh * fB[o];
      const float dar = dan * fC[o];
      const float dqv = dan * fr[o];
      cz[o] = dh * fz[o];
      if (act[o]) {
        const uint32_t d = dg + (uint32_t)((ob[o] * 3) * HR + k) * 4u;
        sts_f32(d, dar); sts_f32(d + HR * 4u, daz); sts_f32(d + 2u * HR * 4u, dqv);
        sts_f32(a_g[o], dar); sts_f32(a_g[o] + h_step, daz); sts_f32(a_g[o] + 2u * h_step, dan);
        sts_f32(a_q[o], dqv);
      }
      a_g[o] -= g_step; a_q[o] -= h_step; a_dy[o] -= h_step; a_h[o] -= h_step;
    }
  };

  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int t0 = pipe.t0_of(c);
    const int tcn = pipe.tcn_of(c);
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const int b = act[o] ? ob[o] : 0;
      const int kk = act[o] ? k : 0;
      a_g[o] = pipe.row_addr(s, 0, b, tcn - 1) + 4u * (uint32_t)kk;
      a_q[o] = pipe.row_addr(s, 1, b, tcn - 1) + 4u * (uint32_t)kk;
      a_dy[o] = pipe.row_addr(s, 2, b, tcn - 1) + 4u * (uint32_t)kk;
      a_h[o] = pipe.row_addr(s, 3, b, tcn - 1) + 4u * (uint32_t)kk;
    }
    fetch(t0 + tcn - 1 == 0);     // first step of the chunk: nothing to hide its fetch behind
    for (int tl = tcn - 1; tl >= 0; --tl) {
      // a_* point at step tl's rows; after the stores they move to step tl-1, which is prefetched unless this
      // is the chunk's last step (tl == 0)
      const bool last = (tl == 0);
      step();
      if (last && pipe.bulk) fence_async_smem();
      __syncthreads();
      // step t-1's operands are fetched HERE, in the same basic block as the mat-vec: their shared-memory latency
      // and the factor arithmetic interleave with the FFMA2 stream instead of sitting in front of the barrier
      if (!last) fetch(t0 + tl - 1 == 0);
      // ---- mat-vec + reduce-scatter over the lane group: the owner of (k, b) receives the complete sum ----
      const uint32_t dg = dgs_addr + (uint32_t)(par * BT * 3 * HR) * 4u;
      float2 accA[BT][3], accB[BT][3];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int g = 0; g < 3; ++g) accA[b][g] = accB[b][g] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KS / 4; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const float4 dv = lds_v4(dg + (uint32_t)((b * 3 + g) * HR) * 4u + (uint32_t)(i * G) * 16u + lane_k);
            accA[b][g] = __ffma2_rn(wt[g][2 * i + 0], make_float2(dv.x, dv.y), accA[b][g]);
            if constexpr (DUAL) accB[b][g] = __ffma2_rn(wt[g][2 * i + 1], make_float2(dv.z, dv.w), accB[b][g]);
            else accA[b][g] = __ffma2_rn(wt[g][2 * i + 1], make_float2(dv.z, dv.w), accA[b][g]);
          }
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float u0 = (accA[b][0].x + accB[b][0].x) + (accA[b][0].y + accB[b][0].y);
        const float u1 = (accA[b][1].x + accB[b][1].x) + (accA[b][1].y + accB[b][1].y);
        const float u2 = (accA[b][2].x + accB[b][2].x) + (accA[b][2].y + accB[b][2].y);
        acc[b] = (u0 + u1) + u2;
      }
      if constexpr (BT < G) {
        float mine = 0.f;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float v = group_sum<G>(acc[b]);
          mine = (ql % BT == b) ? v : mine;
        }
        carry[0] = cz[0] + mine;
      } else if constexpr (G == 2) {
#pragma unroll
        for (int o = 0; o < NOWN; ++o) {
          const float send = ql ? acc[2 * o] : acc[2 * o + 1];
          const float keep = ql ? acc[2 * o + 1] : acc[2 * o];
          carry[o] = cz[o] + keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
      } else {
        const int hi = ql & 2, lo = ql & 1;
#pragma unroll
        for (int o = 0; o < NOWN; ++o) {
          const float s0 = hi ? acc[4 * o + 0] : acc[4 * o + 2];
          const float k0 = hi ? acc[4 * o + 2] : acc[4 * o + 0];
          const float s1 = hi ? acc[4 * o + 1] : acc[4 * o + 3];
          const float k1 = hi ? acc[4 * o + 3] : acc[4 * o + 1];
          const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
          const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
          const float send = lo ? a0 : a1;
          const float keep = lo ? a1 : a0;
          carry[o] = cz[o] + keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
      }
      par ^= 1;
    }
    // the closing barrier of the chunk: all in-place writes of this stage happened before the last
    // step's __syncthreads above, so the stage can be handed to the store engine now
    pipe.release(c);
  }
  pipe.drain();
}

// ---------------------------------------------------------------------------------------------------------------------
// One sequence per CTA, TWO output columns per thread (the c2 regime: B <= 2 x SMs, H <= 64).
//
// The backward mat-vec contracts over the 3H-long dGH vector, three times the forward's h: with one output column per
// lane group every lane fetches 3*H/G values for 3*H/G FMAs x ... -- 24 LDS.128 per 48 FFMA2 at H = 64, and ncu showed
// the shared-memory pipe (LSU 42 %, short-scoreboard the top stall) as what separates this kernel from the forward
// (profiles/r01_ncu_gru_bwd_hotloop.txt).  Here a group of FOUR lanes owns a PAIR of adjacent columns (k, k+1): lane q
// holds rows {(i*4+q)*4..+3} of all three gates for both columns (the same 96 weight registers), so every fetched
// float4 of dGH feeds four FFMA2 instead of two -- 12 LDS.128 per step.  The four partial sums of the two columns are
// combined by an exchange (xor 2) + a butterfly add (xor 1); lanes 2c and 2c+1 of a group then both hold column c's
// total and repeat the same few gate-derivative operations (identical values, identical addresses) instead of idling.
template <int HP, int TC, int NST, bool EXACT>
__global__ void __launch_bounds__(2 * HP, EXACT ? 3 : 2) gru_bwd_pair_kernel(BwdParams p) {
  constexpr int L = 4;                  // lanes per column pair
  constexpr int J = HP / L;             // rows per lane and gate
  constexpr int HR = HP + DG_PAD;
  static_assert(J % 4 == 0, "slice must be float4 granular");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int kp = tid / L, ql = tid % L;
  const int k = 2 * kp + (ql >> 1);     // the column this lane finishes
  const int H = EXACT ? HP : p.H;
  const int T = p.T;
  const int b0 = blockIdx.x;

  float* dgs = reinterpret_cast<float*>(smem_raw);                     // [2][3][HR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dgs + 2 * 3 * HR);
  float* stages = reinterpret_cast<float*>(smem_raw + ((2 * 3 * HR * 4 + NST * 8 + 127) / 128) * 128);

  ChunkPipe<4, 1, TC, NST> pipe;
  pipe.g[0] = const_cast<float*>(p.rzn); pipe.gst[0] = p.dgi; pipe.w[0] = 3 * H; pipe.mode[0] = TG_STRM_LOAD | TG_STRM_STORE; pipe.shift[0] = 0;
  pipe.g[1] = const_cast<float*>(p.q);   pipe.gst[1] = p.dq;  pipe.w[1] = H;     pipe.mode[1] = TG_STRM_LOAD | TG_STRM_STORE; pipe.shift[1] = 0;
  pipe.g[2] = const_cast<float*>(p.dy);  pipe.gst[2] = nullptr; pipe.w[2] = H;   pipe.mode[2] = p.dy_last ? 0 : TG_STRM_LOAD;  pipe.shift[2] = 0;
  pipe.g[3] = const_cast<float*>(p.y);   pipe.gst[3] = nullptr; pipe.w[3] = H;   pipe.mode[3] = TG_STRM_LOAD;                  pipe.shift[3] = -1;
  pipe.layout();
  pipe.stages = stages; pipe.full = bars;
  pipe.T = T; pipe.nb = 1; pipe.b0 = b0; pipe.NC = (T + TC - 1) / TC;
  pipe.reverse = true; pipe.bulk = p.bulk != 0;

  // wt[o][g][m]: rows jj = (i*L+ql)*4 + {0,1 | 2,3} (m = 2i, 2i+1) of gate g, column 2*kp + o
  float2 wt[2][3][J / 2];
#pragma unroll
  for (int o = 0; o < 2; ++o)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < J / 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jj = (i * L + ql) * 4 + c, kk = 2 * kp + o;
          const float v = (kk < H && jj < H) ? p.whh[(size_t)(g * H + jj) * H + kk] : 0.f;
          if (c & 1) wt[o][g][2 * i + (c >> 1)].y = v; else wt[o][g][2 * i + (c >> 1)].x = v;
        }
  for (int i = tid; i < 2 * 3 * HR; i += 2 * HP) dgs[i] = 0.f;

  const bool act = EXACT || k < H;
  const bool use_dy = !p.dy_last;
  float carry = (p.dy_last && act) ? p.dy[(size_t)b0 * H + k] : 0.f;
  pipe.start();
  __syncthreads();

  const uint32_t dgs_addr = smem_u32(dgs);
  const uint32_t h_step = 4u * (uint32_t)H, g_step = 12u * (uint32_t)H;
  const uint32_t lane_k = 16u * (uint32_t)ql;
  float fA, fB, fC, fr, fz, fdy;
  uint32_t a_g, a_q, a_dy, a_h;

  auto fetch = [&](bool first_t) __attribute__((always_inline)) {
    const float r = lds_f32(a_g), z = lds_f32(a_g + h_step), n = lds_f32(a_g + 2u * h_step);
    const float qv = lds_f32(a_q);
    float hp = lds_f32(a_h);
    hp = first_t ? 0.f : hp;
    float dyv = lds_f32(a_dy);
    dyv = use_dy ? dyv : 0.f;
    const float omz = 1.f - z;
    fA = omz * fmaf(-n, n, 1.f);
    fB = (hp - n) * (z * omz);
    fC = qv * (r * (1.f - r));
    fr = r; fz = z; fdy = dyv;
  };

  int par = 0;
  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int t0 = pipe.t0_of(c);
    const int tcn = pipe.tcn_of(c);
    const int kk = act ? k : 0;
    a_g = pipe.row_addr(s, 0, 0, tcn - 1) + 4u * (uint32_t)kk;
    a_q = pipe.row_addr(s, 1, 0, tcn - 1) + 4u * (uint32_t)kk;
    a_dy = pipe.row_addr(s, 2, 0, tcn - 1) + 4u * (uint32_t)kk;
    a_h = pipe.row_addr(s, 3, 0, tcn - 1) + 4u * (uint32_t)kk;
    fetch(t0 + tcn - 1 == 0);
    for (int tl = tcn - 1; tl >= 0; --tl) {
      const bool last = (tl == 0);
      // ---- gate derivatives of step t: three dependent FP32 ops after the carry arrives ----
      const uint32_t dg = dgs_addr + (uint32_t)(par * 3 * HR) * 4u;
      const float dh = fdy + carry;
      const float dan = dh * fA;
      const float daz = dh * fB;
      const float dar = dan * fC;
      const float dqv = dan * fr;
      const float cz = dh * fz;
      if (act) {
        const uint32_t d = dg + (uint32_t)k * 4u;
        sts_f32(d, dar); sts_f32(d + HR * 4u, daz); sts_f32(d + 2u * HR * 4u, dqv);
        sts_f32(a_g, dar); sts_f32(a_g + h_step, daz); sts_f32(a_g + 2u * h_step, dan);
        sts_f32(a_q, dqv);
      }
      a_g -= g_step; a_q -= h_step; a_dy -= h_step; a_h -= h_step;
      if (last && pipe.bulk) fence_async_smem();
      __syncthreads();
      if (!last) fetch(t0 + tl - 1 == 0);      // step t-1's operands, fetched behind the mat-vec
      // ---- W_hh^T dGH_t for the two columns of this lane group ----
      float2 acc[2][3];
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int g = 0; g < 3; ++g) acc[o][g] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < J / 4; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float4 dv = lds_v4(dg + (uint32_t)(g * HR) * 4u + (uint32_t)(i * L) * 16u + lane_k);
          const float2 d01 = make_float2(dv.x, dv.y), d23 = make_float2(dv.z, dv.w);
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            acc[o][g] = __ffma2_rn(wt[o][g][2 * i], d01, acc[o][g]);
            acc[o][g] = __ffma2_rn(wt[o][g][2 * i + 1], d23, acc[o][g]);
          }
        }
      float u[2];
#pragma unroll
      for (int o = 0; o < 2; ++o)
        u[o] = ((acc[o][0].x + acc[o][0].y) + (acc[o][1].x + acc[o][1].y)) + (acc[o][2].x + acc[o][2].y);
      const bool hi = (ql & 2) != 0;
      const float send = hi ? u[0] : u[1], keep = hi ? u[1] : u[0];
      float v = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      carry = cz + v;
      par ^= 1;
    }
    pipe.release(c);
  }
  pipe.drain();
}

template <int HP, int TC, int NST>
int launch_bwd_pair(cudaStream_t st, const BwdParams& p) {
  const int widths[4] = {3 * p.H, p.H, p.H, p.H};
  constexpr int HR = HP + DG_PAD;
  size_t smem = ((2 * 3 * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<4, 1, TC, NST>::stage_floats_for(widths) * 4;
  const bool exact = (p.H == HP);
  auto kern = exact ? gru_bwd_pair_kernel<HP, TC, NST, true> : gru_bwd_pair_kernel<HP, TC, NST, false>;
  if (exact) { TG_OPT_IN_SMEM((gru_bwd_pair_kernel<HP, TC, NST, true>), "gru_bwd_pair"); }
  else { TG_OPT_IN_SMEM((gru_bwd_pair_kernel<HP, TC, NST, false>), "gru_bwd_pair"); }
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_bwd_pair: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  kern<<<p.B, 2 * HP, smem, st>>>(p);
  return tg_check_launch("gru_bwd_pair");
}

template <int HP, int G, int BT, int TC, int NST>
int launch_bwd(cudaStream_t st, const BwdParams& p) {
  const int widths[4] = {3 * p.H, p.H, p.H, p.H};
  constexpr int HR = HP + DG_PAD;
  size_t smem = ((2 * BT * 3 * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<4, BT, TC, NST>::stage_floats_for(widths) * 4;
  const bool exact = (p.H == HP);
  auto kern = exact ? gru_bwd_kernel<HP, G, BT, TC, NST, true> : gru_bwd_kernel<HP, G, BT, TC, NST, false>;
  if (exact) { TG_OPT_IN_SMEM((gru_bwd_kernel<HP, G, BT, TC, NST, true>), "gru_bwd"); }
  else { TG_OPT_IN_SMEM((gru_bwd_kernel<HP, G, BT, TC, NST, false>), "gru_bwd"); }
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_bwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  dim3 grid((p.B + BT - 1) / BT), block(HP * G);
  kern<<<grid, block, smem, st>>>(p);
  return tg_check_launch("gru_bwd");
}

template <int HP, int G>
int dispatch_bt(cudaStream_t st, const BwdParams& p, int bt) {
  constexpr int TC = 8, NST = 3;   // 8-step chunks (H = 128 used 4: twice the per-chunk ring cost per step)
  // one sequence per CTA (the co-resident, latency-bound regime): 16-step chunks halve the per-chunk cost of the
  // bulk-copy ring (thread 0 issues the stores/loads while the other warps wait at the next barrier)
  if constexpr (HP <= 64) {
    // one sequence per CTA: the two-columns-per-thread kernel (half the shared-memory operand traffic)
    if (bt == 1 && tg_bwd_pair()) return tg_long_chunks() ? launch_bwd_pair<HP, 2 * TC, NST>(st, p) : launch_bwd_pair<HP, TC, NST>(st, p);
  }
  if (HP <= 64 && bt == 1 && tg_long_chunks()) return launch_bwd<HP, G, 1, 2 * TC, NST>(st, p);
  switch (bt) {
    case 1: return launch_bwd<HP, G, 1, TC, NST>(st, p);
    case 2: return launch_bwd<HP, G, 2, TC, NST>(st, p);
    case 4: return launch_bwd<HP, G, 4, (HP >= 128) ? 4 : TC, NST>(st, p);   // four H=128 sequences x 8 steps x 3 stages exceed shared memory
  }
  tg_set_error("gru_bwd: bad BT %d", bt);
  return TG_ERR_ARG;
}

}  // namespace

int tg_gru_bwd_impl(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y,
                    const float* whh, float* dgi, float* dq, int B, int T, int H, int flags, const float* whh_t) {
  TG_REQUIRE(dy && rzn && q && y && whh && dgi && dq, TG_ERR_ARG, "gru_bwd: null pointer");
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "gru_bwd: bad shape B=%d T=%d H=%d", B, T, H);
  const int dyl = (flags & TG_GRU_DY_LAST) ? 1 : 0;
  if (tg_cluster_takes(H, B, true) && !(flags & TG_GRU_NO_BULK) && tg_aligned16(rzn) && tg_aligned16(q) && tg_aligned16(y) &&
      tg_aligned16(dgi) && tg_aligned16(dq) && (dyl || tg_aligned16(dy)))
    return tg_gru_cl_bwd(st, dy, rzn, q, y, whh, dgi, dq, B, T, H, dyl);
  if (H > 128) {
    TG_REQUIRE(whh_t, TG_ERR_ARG, "gru_bwd: hidden size %d > 128 needs the transposed weight (w_hh_t)", H);
    return tg_bigh_bwd(st, dy, rzn, q, y, whh_t, dgi, dq, B, T, H, (flags & TG_GRU_DY_LAST) ? 1 : 0);
  }
  BwdParams p{dy, rzn, q, y, whh, dgi, dq, B, T, H, (flags & TG_GRU_DY_LAST) ? 1 : 0, 0};
  p.bulk = (H % 4 == 0) && tg_aligned16(rzn) && tg_aligned16(q) && tg_aligned16(y) && tg_aligned16(dgi) &&
           tg_aligned16(dq) && (p.dy_last || tg_aligned16(dy)) && !(flags & TG_GRU_NO_BULK);
  const int bto = (flags >> 8) & 0xff;
  if (H <= 32) return dispatch_bt<32, 2>(st, p, tg_pick_bt(B, 32, bto));
  if (H <= 64) return dispatch_bt<64, 2>(st, p, tg_pick_bt(B, 64, bto));
  return dispatch_bt<128, 4>(st, p, tg_pick_bt(B, 128, bto));
}

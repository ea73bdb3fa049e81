// Tangent-carrying GRU forward and its reverse, for the discriminator's R1 penalty on sm_100a.
//
// Replaces the generic autograd double backward of timeGAN/train_timegan.py:198-202
// (autograd.grad(d_real.sum(), h_real_n, create_graph=True) -> r1 -> loss.backward()) using the identity
// of SURVEY.md Appendix A.4:  d r1/d theta = (2/B) d/d theta [ JVP_x s (v) ],  v = stopgrad(ds/dx).
//   * gru_jvp_fwd : with the primal pass already saved (r,z,n,q,y), carries the tangent state hdot_t
//                   through one layer (input tangent projection gid = xdot W_ih^T pre-computed).
//   * gru_jvp_bwd : reverse of (primal + tangent) forward: two adjoint carries (for h and hdot), two
//                   W_hh^T mat-vecs per step sharing the same register-resident weights.
// Formulas are those of oracle/gru_math.py (gru_layer_jvp / gru_layer_jvp_bwd), which the CPU tests pin
// against torch autograd.  Kernel structure is the same as gru_fwd.cu / gru_bwd.cu: lane groups of G lanes per
// hidden unit, packed FFMA2 mat-vecs, shuffle reduce-scatter to the lane that owns (unit, sequence), carried
// states in that lane's registers, one __syncthreads per step.
#include "chunk_pipe.cuh"
#include "kernels.h"

namespace {

// =============================== tangent forward ===============================================
struct JfParams {
  float* gid;        // (B,T,3H) in: xdot W_ih^T ; out: [a_r, a_z, a_n] (tangent pre-activations)
  const float* rzn;  // (B,T,3H)
  const float* q;    // (B,T,H)
  const float* y;    // (B,T,H)
  const float* whh;  // (3H,H)
  float* ydot;       // (B,T,H) out
  float* qdot;       // (B,T,H) out
  int B, T, H;
  int bulk;
};

constexpr int JV_PAD = 16;   // smem state rows are HP+16 floats apart (bank spread between a lane pair's sequences)

template <int HP, int G>
constexpr int jvp_min_blocks() { return (HP * G <= 128) ? 2 : 1; }

// lane-group combine shared by both kernels: own = complete sum for the sequence this lane owns
// (b = ql when BT < G, else b = o*G + ql); NV values per sequence.
template <int G, int BT, int NV>
__device__ __forceinline__ void jvp_reduce_scatter(float (&acc)[BT][NV], float (&own)[(BT >= G) ? BT / G : 1][NV], int ql) {
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  if constexpr (BT < G) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float mine = 0.f;
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float t = group_sum<G>(acc[b][v]);
        mine = (ql % BT == b) ? t : mine;
      }
      own[0][v] = mine;
    }
  } else if constexpr (G == 2) {
#pragma unroll
    for (int o = 0; o < NOWN; ++o)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float send = ql ? acc[2 * o][v] : acc[2 * o + 1][v];
        const float keep = ql ? acc[2 * o + 1][v] : acc[2 * o][v];
        own[o][v] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
  } else {
    const int hi = ql & 2, lo = ql & 1;
#pragma unroll
    for (int o = 0; o < NOWN; ++o)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float s0 = hi ? acc[4 * o + 0][v] : acc[4 * o + 2][v];
        const float k0 = hi ? acc[4 * o + 2][v] : acc[4 * o + 0][v];
        const float s1 = hi ? acc[4 * o + 1][v] : acc[4 * o + 3][v];
        const float k1 = hi ? acc[4 * o + 3][v] : acc[4 * o + 1][v];
        const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
        const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
        const float send = lo ? a0 : a1;
        const float keep = lo ? a1 : a0;
        own[o][v] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
  }
}

// The tangent recurrence is LINEAR in hdot (the gates are those of the saved primal pass), so once the lane group's
// W_hh hdot_{t-1} sums arrive only a handful of dependent FP32 ops remain:
//     a_r = gid_r + s_r;  rdot = c1 a_r;  a_n = gid_n + rdot q + r s_n;  ndot = c3 a_n;
//     a_z = gid_z + s_z;  zdot = c2 a_z;  hdot = (1-z) ndot + z hdot_{t-1} + zdot (h_{t-1} - n)
// with c1 = r(1-r), c2 = z(1-z), c3 = 1-n^2.  The saved operands are fetched (and the coefficients formed) in the
// same basic block as the mat-vec, so they hide behind its FFMA2 stream; the body is branch-free and EXACT
// (H == HP) turns the strides into immediates.
template <int HP, int G, int BT, int TC, int NST, bool EXACT>
__global__ void __launch_bounds__(HP* G, jvp_min_blocks<HP, G>()) gru_jvp_fwd_kernel(JfParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + JV_PAD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int j = tid / G, ql = tid % G;
  const int H = EXACT ? HP : p.H;
  const int T = p.T;
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, p.B - b0);

  float* hs = reinterpret_cast<float*>(smem_raw);  // [2][BT][HR]  tangent state
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + 2 * BT * HR);
  float* stages = reinterpret_cast<float*>(smem_raw + ((2 * BT * HR * 4 + NST * 8 + 127) / 128) * 128);

  ChunkPipe<6, BT, TC, NST> pipe;
  pipe.g[0] = pipe.gst[0] = p.gid;               pipe.w[0] = 3 * H; pipe.mode[0] = TG_STRM_LOAD | TG_STRM_STORE; pipe.shift[0] = 0;
  pipe.g[1] = const_cast<float*>(p.rzn); pipe.gst[1] = nullptr; pipe.w[1] = 3 * H; pipe.mode[1] = TG_STRM_LOAD; pipe.shift[1] = 0;
  pipe.g[2] = const_cast<float*>(p.q);   pipe.gst[2] = nullptr; pipe.w[2] = H;     pipe.mode[2] = TG_STRM_LOAD; pipe.shift[2] = 0;
  pipe.g[3] = const_cast<float*>(p.y);   pipe.gst[3] = nullptr; pipe.w[3] = H;     pipe.mode[3] = TG_STRM_LOAD; pipe.shift[3] = -1;
  pipe.g[4] = pipe.gst[4] = p.qdot;              pipe.w[4] = H;     pipe.mode[4] = TG_STRM_STORE; pipe.shift[4] = 0;
  pipe.g[5] = pipe.gst[5] = p.ydot;              pipe.w[5] = H;     pipe.mode[5] = TG_STRM_STORE; pipe.shift[5] = 0;
  pipe.layout();
  pipe.stages = stages; pipe.full = bars;
  pipe.T = T; pipe.nb = nb; pipe.b0 = b0; pipe.NC = (T + TC - 1) / TC;
  pipe.reverse = false; pipe.bulk = p.bulk != 0;

  float2 w[3][KS / 2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int i = 0; i < KS / 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int k = (i * G + ql) * 4 + c;
        const float v = (j < H && k < H) ? p.whh[(size_t)(g * H + j) * H + k] : 0.f;
        if (c & 1) w[g][2 * i + (c >> 1)].y = v; else w[g][2 * i + (c >> 1)].x = v;
      }
  for (int i = tid; i < 2 * BT * HR; i += HP * G) hs[i] = 0.f;
  float hdprev[NOWN];
  bool act[NOWN];
  int ob[NOWN];
#pragma unroll
  for (int o = 0; o < NOWN; ++o) {
    hdprev[o] = 0.f;
    ob[o] = (BT < G) ? ql % BT : o * G + ql;     // surplus lanes of a group repeat a sibling's (identical) work
    act[o] = (EXACT && BT == 1) || ((EXACT || j < H) && (ob[o] < nb));
  }
  pipe.start();
  __syncthreads();

  const uint32_t hs_addr = smem_u32(hs);
  const uint32_t h_step = 4u * (uint32_t)H, g_step = 12u * (uint32_t)H;
  const uint32_t lane_k = 16u * (uint32_t)ql;
  uint32_t a_g[NOWN], a_s[NOWN], a_q[NOWN], a_h[NOWN], a_qd[NOWN], a_yd[NOWN];
  float gr[NOWN], gz[NOWN], gn[NOWN], c1[NOWN], c2[NOWN], c3[NOWN], fr[NOWN], fz[NOWN], fq[NOWN], fomz[NOWN], fhmn[NOWN];

  auto fetch = [&](bool first_t) __attribute__((always_inline)) {
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      gr[o] = lds_f32(a_g[o]); gz[o] = lds_f32(a_g[o] + h_step); gn[o] = lds_f32(a_g[o] + 2u * h_step);
      const float r = lds_f32(a_s[o]), z = lds_f32(a_s[o] + h_step), n = lds_f32(a_s[o] + 2u * h_step);
      fq[o] = lds_f32(a_q[o]);
      float hp = lds_f32(a_h[o]);
      hp = first_t ? 0.f : hp;                 // h_{-1} = 0 (the shifted stream never loads row -1)
      fr[o] = r; fz[o] = z;
      c1[o] = r * (1.f - r);
      fomz[o] = 1.f - z;
      c2[o] = z * fomz[o];
      c3[o] = fmaf(-n, n, 1.f);
      fhmn[o] = hp - n;
    }
  };

  int cur = 0;
  auto step = [&](bool first_t) __attribute__((always_inline)) {
    const uint32_t hc = hs_addr + (uint32_t)(cur * BT * HR) * 4u;
    const uint32_t hn = hs_addr + (uint32_t)((cur ^ 1) * BT * HR) * 4u;
    fetch(first_t);      // same basic block as the mat-vec: the loads and the coefficient math hide behind it
    float2 accA[BT][3], accB[BT][3];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int g = 0; g < 3; ++g) accA[b][g] = accB[b][g] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < KS / 4; ++i)
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = lds_v4(hc + (uint32_t)(b * HR) * 4u + (uint32_t)(i * G) * 16u + lane_k);
        const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          accA[b][g] = __ffma2_rn(w[g][2 * i + 0], h01, accA[b][g]);
          if constexpr (HP * G <= 128) accB[b][g] = __ffma2_rn(w[g][2 * i + 1], h23, accB[b][g]);
          else accA[b][g] = __ffma2_rn(w[g][2 * i + 1], h23, accA[b][g]);
        }
      }
    float acc[BT][3], own[NOWN][3];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int g = 0; g < 3; ++g) acc[b][g] = (accA[b][g].x + accB[b][g].x) + (accA[b][g].y + accB[b][g].y);
    jvp_reduce_scatter<G, BT, 3>(acc, own, ql);
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const float qd = own[o][2];
      const float a_r = gr[o] + own[o][0];
      const float a_z = gz[o] + own[o][1];
      const float rdot = c1[o] * a_r;
      const float zdot = c2[o] * a_z;
      const float a_n = fmaf(fr[o], qd, fmaf(rdot, fq[o], gn[o]));
      const float ndot = c3[o] * a_n;
      const float hd = fmaf(zdot, fhmn[o], fmaf(fz[o], hdprev[o], fomz[o] * ndot));
      hdprev[o] = hd;
      if (act[o]) {
        sts_f32(hn + (uint32_t)(ob[o] * HR + j) * 4u, hd);
        sts_f32(a_yd[o], hd);
        sts_f32(a_qd[o], qd);
        sts_f32(a_g[o], a_r); sts_f32(a_g[o] + h_step, a_z); sts_f32(a_g[o] + 2u * h_step, a_n);
      }
      a_g[o] += g_step; a_s[o] += g_step; a_q[o] += h_step; a_h[o] += h_step; a_qd[o] += h_step; a_yd[o] += h_step;
    }
    cur ^= 1;
  };

  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int t0 = pipe.t0_of(c);
    const int tcn = pipe.tcn_of(c);
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const int b = act[o] ? ob[o] : 0;
      const int jj = act[o] ? j : 0;
      a_g[o] = pipe.row_addr(s, 0, b, 0) + 4u * (uint32_t)jj;
      a_s[o] = pipe.row_addr(s, 1, b, 0) + 4u * (uint32_t)jj;
      a_q[o] = pipe.row_addr(s, 2, b, 0) + 4u * (uint32_t)jj;
      a_h[o] = pipe.row_addr(s, 3, b, 0) + 4u * (uint32_t)jj;
      a_qd[o] = pipe.row_addr(s, 4, b, 0) + 4u * (uint32_t)jj;
      a_yd[o] = pipe.row_addr(s, 5, b, 0) + 4u * (uint32_t)jj;
    }
    for (int tl = 0; tl < tcn - 1; ++tl) {
      step(t0 + tl == 0);
      __syncthreads();
    }
    step(t0 + tcn - 1 == 0);
    if (pipe.bulk) fence_async_smem();
    __syncthreads();
    pipe.release(c);
  }
  pipe.drain();
}

// =============================== reverse over (primal + tangent) ==============================
struct JbParams {
  const float* hbar;   // (B,T,H) or (B,H) if last-only : adjoint of y
  const float* hdbar;  // (B,T,H) or (B,H) if last-only : adjoint of ydot
  const float* rzn;    // (B,T,3H)
  const float* q;      // (B,T,H)
  const float* ta;     // (B,T,3H) tangent pre-activations a_r,a_z,a_n
  const float* qdot;   // (B,T,H)
  const float* y;      // (B,T,H)
  const float* ydot;   // (B,T,H)
  const float* whh;
  float* gib;   // (B,T,3H) out: adjoint of primal gi
  float* qb;    // (B,T,H)  out: adjoint of primal q   (gh_n)
  float* gidb;  // (B,T,3H) out: adjoint of tangent gi
  float* qdb;   // (B,T,H)  out: adjoint of tangent q
  int B, T, H;
  int last_only;
  int bulk;
};

// Every output of a step is LINEAR in the two incoming adjoints hb = hbar_t + carry_h, hdb = hdbar_t + carry_hd,
// with coefficients that only depend on saved activations:  out = alpha hb + beta hdb.  The coefficient pairs of
// step t-1 are formed in the same basic block as step t's mat-vecs (operands fetched from the stage right after
// the barrier), so the chain between the
// reduce-scatter and the shared-memory exchange is  FADD -> FMUL -> FFMA.  (Derivation: oracle/gru_math.py
// gru_layer_jvp_bwd; K1 = (1-n^2)(1-z), K2 = (1-n^2)(-zdot - 2 n a_n (1-z)).)
template <int HP, int G, int BT, int TC, int NST, bool EXACT>
__global__ void __launch_bounds__(HP* G, jvp_min_blocks<HP, G>()) gru_jvp_bwd_kernel(JbParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + JV_PAD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int k = tid / G, ql = tid % G;
  const int H = EXACT ? HP : p.H;
  const int T = p.T;
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, p.B - b0);

  float* dgs = reinterpret_cast<float*>(smem_raw);  // [2][BT][6][HR] : primal dGH (3) then tangent dGH (3)
  uint64_t* bars = reinterpret_cast<uint64_t*>(dgs + 2 * BT * 6 * HR);
  float* stages = reinterpret_cast<float*>(smem_raw + ((2 * BT * 6 * HR * 4 + NST * 8 + 127) / 128) * 128);

  ChunkPipe<8, BT, TC, NST> pipe;
  const int ld = TG_STRM_LOAD, lst = TG_STRM_LOAD | TG_STRM_STORE;
  pipe.g[0] = const_cast<float*>(p.rzn);  pipe.gst[0] = p.gib;   pipe.w[0] = 3 * H; pipe.mode[0] = lst; pipe.shift[0] = 0;
  pipe.g[1] = const_cast<float*>(p.q);    pipe.gst[1] = p.qb;    pipe.w[1] = H;     pipe.mode[1] = lst; pipe.shift[1] = 0;
  pipe.g[2] = const_cast<float*>(p.ta);   pipe.gst[2] = p.gidb;  pipe.w[2] = 3 * H; pipe.mode[2] = lst; pipe.shift[2] = 0;
  pipe.g[3] = const_cast<float*>(p.qdot); pipe.gst[3] = p.qdb;   pipe.w[3] = H;     pipe.mode[3] = lst; pipe.shift[3] = 0;
  pipe.g[4] = const_cast<float*>(p.y);    pipe.gst[4] = nullptr; pipe.w[4] = H;     pipe.mode[4] = ld;  pipe.shift[4] = -1;
  pipe.g[5] = const_cast<float*>(p.ydot); pipe.gst[5] = nullptr; pipe.w[5] = H;     pipe.mode[5] = ld;  pipe.shift[5] = -1;
  pipe.g[6] = const_cast<float*>(p.hbar); pipe.gst[6] = nullptr; pipe.w[6] = H;     pipe.mode[6] = p.last_only ? 0 : ld; pipe.shift[6] = 0;
  pipe.g[7] = const_cast<float*>(p.hdbar);pipe.gst[7] = nullptr; pipe.w[7] = H;     pipe.mode[7] = p.last_only ? 0 : ld; pipe.shift[7] = 0;
  pipe.layout();
  pipe.stages = stages; pipe.full = bars;
  pipe.T = T; pipe.nb = nb; pipe.b0 = b0; pipe.NC = (T + TC - 1) / TC;
  pipe.reverse = true; pipe.bulk = p.bulk != 0;

  float2 wt[3][KS / 2];
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
  for (int i = tid; i < 2 * BT * 6 * HR; i += HP * G) dgs[i] = 0.f;
  // carried adjoints; last-only mode (R1: the head reads y[:, T-1] / ydot[:, T-1]): the (B,H) adjoints ARE the
  // initial carries and no per-step adjoint stream exists
  float ch[NOWN], chd[NOWN];
  bool act[NOWN];
  int ob[NOWN];
  const bool use_bar = !p.last_only;
#pragma unroll
  for (int o = 0; o < NOWN; ++o) {
    ob[o] = (BT < G) ? ql % BT : o * G + ql;
    act[o] = (EXACT && BT == 1) || ((EXACT || k < H) && (ob[o] < nb));
    ch[o] = (p.last_only && act[o]) ? p.hbar[(size_t)(b0 + ob[o]) * H + k] : 0.f;
    chd[o] = (p.last_only && act[o]) ? p.hdbar[(size_t)(b0 + ob[o]) * H + k] : 0.f;
  }
  pipe.start();
  __syncthreads();

  const uint32_t dgs_addr = smem_u32(dgs);
  const uint32_t h_step = 4u * (uint32_t)H, g_step = 12u * (uint32_t)H;
  const uint32_t lane_k = 16u * (uint32_t)ql;
  uint32_t a_g[NOWN], a_q[NOWN], a_t[NOWN], a_qd[NOWN], a_h[NOWN], a_hd[NOWN], a_hb[NOWN], a_hdb[NOWN];
  // coefficient pairs (alpha on hb, beta on hdb) of the step about to be processed
  float arA[NOWN], arB[NOWN], azA[NOWN], azB[NOWN], anA[NOWN], anB[NOWN], qA[NOWN], qB[NOWN];
  float ardB[NOWN], azdB[NOWN], andB[NOWN], qdB[NOWN], nhA[NOWN], nhB[NOWN], fhb[NOWN], fhdb[NOWN];

  auto fetch = [&](bool first_t) __attribute__((always_inline)) {
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const float rt = lds_f32(a_g[o]), zt = lds_f32(a_g[o] + h_step), nt = lds_f32(a_g[o] + 2u * h_step);
      const float qt = lds_f32(a_q[o]);
      const float art = lds_f32(a_t[o]), azt = lds_f32(a_t[o] + h_step), ant = lds_f32(a_t[o] + 2u * h_step);
      const float qdt = lds_f32(a_qd[o]);
      float hp = lds_f32(a_h[o]), hdp = lds_f32(a_hd[o]);
      hp = first_t ? 0.f : hp;
      hdp = first_t ? 0.f : hdp;
      float hb = lds_f32(a_hb[o]), hdb = lds_f32(a_hdb[o]);
      fhb[o] = use_bar ? hb : 0.f;
      fhdb[o] = use_bar ? hdb : 0.f;
      const float sr = rt * (1.f - rt), omz = 1.f - zt, sz = zt * omz, sn = fmaf(-nt, nt, 1.f);
      const float rdot = sr * art, zdot = sz * azt, ndot = sn * ant, hmn = hp - nt;
      const float K1 = sn * omz;
      const float K2 = sn * (-zdot - 2.f * nt * ant * omz);
      // anb = K1 hb + K2 hdb ; anb_d = K1 hdb
      anA[o] = K1; anB[o] = K2; andB[o] = K1;
      qdB[o] = rt * K1;
      const float rdbB = qt * K1;                       // rdb = rdbB hdb
      ardB[o] = sr * rdbB;
      qA[o] = rt * K1; qB[o] = fmaf(rdot, K1, rt * K2);
      // rb = qdt anb_d + qt anb + (1-2r) a_r rdb
      arA[o] = sr * (qt * K1);
      arB[o] = sr * fmaf(qdt, K1, fmaf(qt, K2, (1.f - 2.f * rt) * art * rdbB));
      azdB[o] = sz * hmn;
      azA[o] = sz * hmn;
      azB[o] = sz * ((hdp - ndot) + (1.f - 2.f * zt) * azt * hmn);
      nhA[o] = zt; nhB[o] = zdot;
      // nhd = zt hdb (nhA doubles as its coefficient)
    }
  };

  int par = 0;
  float nh[NOWN], nhd[NOWN];
  auto step = [&]() __attribute__((always_inline)) {
    const uint32_t dg = dgs_addr + (uint32_t)(par * BT * 6 * HR) * 4u;
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const float hb = fhb[o] + ch[o];
      const float hdb = fhdb[o] + chd[o];
      const float arb = fmaf(arA[o], hb, arB[o] * hdb);
      const float azb = fmaf(azA[o], hb, azB[o] * hdb);
      const float anb = fmaf(anA[o], hb, anB[o] * hdb);
      const float qb = fmaf(qA[o], hb, qB[o] * hdb);
      const float arb_d = ardB[o] * hdb, azb_d = azdB[o] * hdb, anb_d = andB[o] * hdb, qdb = qdB[o] * hdb;
      nh[o] = fmaf(nhA[o], hb, nhB[o] * hdb);
      nhd[o] = nhA[o] * hdb;
      if (act[o]) {
        const uint32_t d = dg + (uint32_t)((ob[o] * 6) * HR + k) * 4u;
        sts_f32(d, arb); sts_f32(d + HR * 4u, azb); sts_f32(d + 2u * HR * 4u, qb);
        sts_f32(d + 3u * HR * 4u, arb_d); sts_f32(d + 4u * HR * 4u, azb_d); sts_f32(d + 5u * HR * 4u, qdb);
        sts_f32(a_g[o], arb); sts_f32(a_g[o] + h_step, azb); sts_f32(a_g[o] + 2u * h_step, anb); sts_f32(a_q[o], qb);
        sts_f32(a_t[o], arb_d); sts_f32(a_t[o] + h_step, azb_d); sts_f32(a_t[o] + 2u * h_step, anb_d);
        sts_f32(a_qd[o], qdb);
      }
      a_g[o] -= g_step; a_q[o] -= h_step; a_t[o] -= g_step; a_qd[o] -= h_step;
      a_h[o] -= h_step; a_hd[o] -= h_step; a_hb[o] -= h_step; a_hdb[o] -= h_step;
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
      const uint32_t kk = 4u * (uint32_t)(act[o] ? k : 0);
      a_g[o] = pipe.row_addr(s, 0, b, tcn - 1) + kk;
      a_q[o] = pipe.row_addr(s, 1, b, tcn - 1) + kk;
      a_t[o] = pipe.row_addr(s, 2, b, tcn - 1) + kk;
      a_qd[o] = pipe.row_addr(s, 3, b, tcn - 1) + kk;
      a_h[o] = pipe.row_addr(s, 4, b, tcn - 1) + kk;
      a_hd[o] = pipe.row_addr(s, 5, b, tcn - 1) + kk;
      a_hb[o] = pipe.row_addr(s, 6, b, tcn - 1) + kk;
      a_hdb[o] = pipe.row_addr(s, 7, b, tcn - 1) + kk;
    }
    fetch(t0 + tcn - 1 == 0);
    for (int tl = tcn - 1; tl >= 0; --tl) {
      const bool last = (tl == 0);
      step();
      if (last && pipe.bulk) fence_async_smem();
      __syncthreads();
      if (!last) fetch(t0 + tl - 1 == 0);    // next step's coefficient pairs form behind the mat-vecs below
      const uint32_t dg = dgs_addr + (uint32_t)(par * BT * 6 * HR) * 4u;
      float2 a2[BT][2][3];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int v = 0; v < 2; ++v) a2[b][v][0] = a2[b][v][1] = a2[b][v][2] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KS / 4; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const uint32_t off = (uint32_t)(i * G) * 16u + lane_k;
            const float4 dv = lds_v4(dg + (uint32_t)((b * 6 + g) * HR) * 4u + off);
            const float4 ev = lds_v4(dg + (uint32_t)((b * 6 + 3 + g) * HR) * 4u + off);
            a2[b][0][g] = __ffma2_rn(wt[g][2 * i + 0], make_float2(dv.x, dv.y), a2[b][0][g]);
            a2[b][1][g] = __ffma2_rn(wt[g][2 * i + 0], make_float2(ev.x, ev.y), a2[b][1][g]);
            a2[b][0][g] = __ffma2_rn(wt[g][2 * i + 1], make_float2(dv.z, dv.w), a2[b][0][g]);
            a2[b][1][g] = __ffma2_rn(wt[g][2 * i + 1], make_float2(ev.z, ev.w), a2[b][1][g]);
          }
      float acc[BT][2], own[NOWN][2];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int v = 0; v < 2; ++v)
          acc[b][v] = (a2[b][v][0].x + a2[b][v][0].y) + (a2[b][v][1].x + a2[b][v][1].y) + (a2[b][v][2].x + a2[b][v][2].y);
      jvp_reduce_scatter<G, BT, 2>(acc, own, ql);
#pragma unroll
      for (int o = 0; o < NOWN; ++o) {
        ch[o] = nh[o] + own[o][0];
        chd[o] = nhd[o] + own[o][1];
      }
      par ^= 1;
    }
    pipe.release(c);
  }
  pipe.drain();
}

template <int HP, int G, int BT, int TC, int NST>
int launch_jf(cudaStream_t st, const JfParams& p) {
  const int widths[6] = {3 * p.H, 3 * p.H, p.H, p.H, p.H, p.H};
  constexpr int HR = HP + JV_PAD;
  size_t smem = ((2 * BT * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<6, BT, TC, NST>::stage_floats_for(widths) * 4;
  const bool exact = (p.H == HP);
  auto kern = exact ? gru_jvp_fwd_kernel<HP, G, BT, TC, NST, true> : gru_jvp_fwd_kernel<HP, G, BT, TC, NST, false>;
  if (exact) { TG_OPT_IN_SMEM((gru_jvp_fwd_kernel<HP, G, BT, TC, NST, true>), "gru_jvp_fwd"); }
  else { TG_OPT_IN_SMEM((gru_jvp_fwd_kernel<HP, G, BT, TC, NST, false>), "gru_jvp_fwd"); }
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_jvp_fwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  kern<<<dim3((p.B + BT - 1) / BT), dim3(HP * G), smem, st>>>(p);
  return tg_check_launch("gru_jvp_fwd");
}

template <int HP, int G, int BT, int TC, int NST>
int launch_jb(cudaStream_t st, const JbParams& p) {
  const int widths[8] = {3 * p.H, p.H, 3 * p.H, p.H, p.H, p.H, p.H, p.H};
  constexpr int HR = HP + JV_PAD;
  size_t smem = ((2 * BT * 6 * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<8, BT, TC, NST>::stage_floats_for(widths) * 4;
  const bool exact = (p.H == HP);
  auto kern = exact ? gru_jvp_bwd_kernel<HP, G, BT, TC, NST, true> : gru_jvp_bwd_kernel<HP, G, BT, TC, NST, false>;
  if (exact) { TG_OPT_IN_SMEM((gru_jvp_bwd_kernel<HP, G, BT, TC, NST, true>), "gru_jvp_bwd"); }
  else { TG_OPT_IN_SMEM((gru_jvp_bwd_kernel<HP, G, BT, TC, NST, false>), "gru_jvp_bwd"); }
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_jvp_bwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  kern<<<dim3((p.B + BT - 1) / BT), dim3(HP * G), smem, st>>>(p);
  return tg_check_launch("gru_jvp_bwd");
}

// The R1 pass only ever runs on the discriminator stack: BT is limited to {1,2} to bound shared memory
// (8 streamed arrays per stage).
template <int HP, int G>
int dispatch_jf(cudaStream_t st, const JfParams& p, int bt) {
  constexpr int TC = (HP >= 128) ? 4 : 8, NST = 3;
  if (bt >= 2) return launch_jf<HP, G, 2, TC, NST>(st, p);
  return launch_jf<HP, G, 1, TC, NST>(st, p);
}
template <int HP, int G>
int dispatch_jb(cudaStream_t st, const JbParams& p, int bt) {
  constexpr int TC = (HP >= 128) ? 4 : 8, NST = 3;
  if (bt >= 2) return launch_jb<HP, G, 2, TC, NST>(st, p);
  return launch_jb<HP, G, 1, TC, NST>(st, p);
}

}  // namespace

int tg_gru_jvp_fwd_impl(cudaStream_t st, float* gid, const float* rzn, const float* q, const float* y,
                        const float* whh, float* ydot, float* qdot, int B, int T, int H, int flags) {
  TG_REQUIRE(gid && rzn && q && y && whh && ydot && qdot, TG_ERR_ARG, "gru_jvp_fwd: null pointer");
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "gru_jvp_fwd: bad shape B=%d T=%d H=%d", B, T, H);
  if (H > 128) return tg_bigh_jvp_fwd(st, gid, rzn, q, y, whh, ydot, qdot, B, T, H);
  JfParams p{gid, rzn, q, y, whh, ydot, qdot, B, T, H, 0};
  p.bulk = (H % 4 == 0) && tg_aligned16(gid) && tg_aligned16(rzn) && tg_aligned16(q) && tg_aligned16(y) &&
           tg_aligned16(ydot) && tg_aligned16(qdot) && !(flags & TG_GRU_NO_BULK);
  const int bto = (flags >> 8) & 0xff;
  if (H <= 32) return dispatch_jf<32, 2>(st, p, tg_pick_bt(B, 32, bto));
  if (H <= 64) return dispatch_jf<64, 2>(st, p, tg_pick_bt(B, 64, bto));
  return dispatch_jf<128, 4>(st, p, tg_pick_bt(B, 128, bto));
}

int tg_gru_jvp_bwd_impl(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                        const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh,
                        float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int flags,
                        const float* whh_t) {
  TG_REQUIRE(hbar && hdbar && rzn && q && ta && qdot && y && ydot && whh && gib && qb && gidb && qdb, TG_ERR_ARG,
             "gru_jvp_bwd: null pointer");
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "gru_jvp_bwd: bad shape B=%d T=%d H=%d", B, T, H);
  {
    const int lo = (flags & TG_GRU_DY_LAST) ? 1 : 0;
    if (tg_cluster_takes_jvp_bwd(H, B) && !(flags & TG_GRU_NO_BULK) && tg_aligned16(rzn) && tg_aligned16(q) &&
        tg_aligned16(ta) && tg_aligned16(qdot) && tg_aligned16(y) && tg_aligned16(ydot) && tg_aligned16(gib) &&
        tg_aligned16(qb) && tg_aligned16(gidb) && tg_aligned16(qdb) && (lo || (tg_aligned16(hbar) && tg_aligned16(hdbar))))
      return tg_gru_cl_jvp_bwd(st, hbar, hdbar, rzn, q, ta, qdot, y, ydot, whh, gib, qb, gidb, qdb, B, T, H, lo);
  }
  if (H > 128) {
    TG_REQUIRE(whh_t, TG_ERR_ARG, "gru_jvp_bwd: hidden size %d > 128 needs the transposed weight (w_hh_t)", H);
    return tg_bigh_jvp_bwd(st, hbar, hdbar, rzn, q, ta, qdot, y, ydot, whh_t, gib, qb, gidb, qdb, B, T, H,
                           (flags & TG_GRU_DY_LAST) ? 1 : 0);
  }
  JbParams p{hbar, hdbar, rzn, q, ta, qdot, y, ydot, whh, gib, qb, gidb, qdb, B, T, H,
             (flags & TG_GRU_DY_LAST) ? 1 : 0, 0};
  p.bulk = (H % 4 == 0) && tg_aligned16(rzn) && tg_aligned16(q) && tg_aligned16(ta) && tg_aligned16(qdot) &&
           tg_aligned16(y) && tg_aligned16(ydot) && tg_aligned16(gib) && tg_aligned16(qb) && tg_aligned16(gidb) &&
           tg_aligned16(qdb) && (p.last_only || (tg_aligned16(hbar) && tg_aligned16(hdbar))) &&
           !(flags & TG_GRU_NO_BULK);
  const int bto = (flags >> 8) & 0xff;
  if (H <= 32) return dispatch_jb<32, 2>(st, p, tg_pick_bt(B, 32, bto));
  if (H <= 64) return dispatch_jb<64, 2>(st, p, tg_pick_bt(B, 64, bto));
  return dispatch_jb<128, 4>(st, p, tg_pick_bt(B, 128, bto));
}

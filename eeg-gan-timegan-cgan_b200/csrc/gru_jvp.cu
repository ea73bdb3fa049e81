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
        mine = (ql == b) ? t : mine;
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

template <int HP, int G, int BT, int TC, int NST>
__global__ void __launch_bounds__(HP* G, jvp_min_blocks<HP, G>()) gru_jvp_fwd_kernel(JfParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + JV_PAD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int j = tid / G, ql = tid % G;
  const int H = p.H, T = p.T;
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
#pragma unroll
  for (int o = 0; o < NOWN; ++o) hdprev[o] = 0.f;
  pipe.start();
  __syncthreads();

  int cur = 0;
  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int t0 = pipe.t0_of(c);
    const int tcn = pipe.tcn_of(c);
    for (int tl = 0; tl < tcn; ++tl) {
      const float* hc = hs + cur * BT * HR;
      float* hn = hs + (cur ^ 1) * BT * HR;
      float2 acc2[BT][3];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc2[b][0] = acc2[b][1] = acc2[b][2] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KS / 4; ++i)
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 hv = reinterpret_cast<const float4*>(hc + b * HR)[i * G + ql];
          const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            acc2[b][g] = __ffma2_rn(w[g][2 * i + 0], h01, acc2[b][g]);
            acc2[b][g] = __ffma2_rn(w[g][2 * i + 1], h23, acc2[b][g]);
          }
        }
      float acc[BT][3], own[NOWN][3];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int g = 0; g < 3; ++g) acc[b][g] = acc2[b][g].x + acc2[b][g].y;
      jvp_reduce_scatter<G, BT, 3>(acc, own, ql);
#pragma unroll
      for (int o = 0; o < NOWN; ++o) {
        const int b = (BT < G) ? ql : o * G + ql;
        if (j < H && b < nb) {
          float* gp = pipe.row(s, 0, b, tl);
          const float* sp = pipe.row(s, 1, b, tl);
          const float r = sp[j], z = sp[H + j], n = sp[2 * H + j];
          const float qv = pipe.row(s, 2, b, tl)[j];
          const float hp = (t0 + tl > 0) ? pipe.row(s, 3, b, tl)[j] : 0.f;
          const float a_r = gp[j] + own[o][0];
          const float a_z = gp[H + j] + own[o][1];
          const float qd = own[o][2];
          const float rdot = r * (1.f - r) * a_r;
          const float zdot = z * (1.f - z) * a_z;
          const float a_n = gp[2 * H + j] + rdot * qv + r * qd;
          const float ndot = (1.f - n * n) * a_n;
          const float hd = (1.f - z) * ndot + z * hdprev[o] + zdot * (hp - n);
          hdprev[o] = hd;
          hn[b * HR + j] = hd;
          gp[j] = a_r; gp[H + j] = a_z; gp[2 * H + j] = a_n;
          pipe.row(s, 4, b, tl)[j] = qd;
          pipe.row(s, 5, b, tl)[j] = hd;
        }
      }
      if (tl == tcn - 1 && pipe.bulk) fence_async_smem();
      __syncthreads();
      cur ^= 1;
    }
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

template <int HP, int G, int BT, int TC, int NST>
__global__ void __launch_bounds__(HP* G, jvp_min_blocks<HP, G>()) gru_jvp_bwd_kernel(JbParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + JV_PAD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int k = tid / G, ql = tid % G;
  const int H = p.H, T = p.T;
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
  float ch[NOWN], chd[NOWN];
#pragma unroll
  for (int o = 0; o < NOWN; ++o) ch[o] = chd[o] = 0.f;
  pipe.start();
  __syncthreads();

  int par = 0;
  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int t0 = pipe.t0_of(c);
    const int tcn = pipe.tcn_of(c);
    for (int tl = tcn - 1; tl >= 0; --tl) {
      const int t = t0 + tl;
      float* dg = dgs + par * BT * 6 * HR;
      float nh[NOWN], nhd[NOWN];
#pragma unroll
      for (int o = 0; o < NOWN; ++o) {
        const int b = (BT < G) ? ql : o * G + ql;
        nh[o] = nhd[o] = 0.f;
        if (k < H && b < nb) {
          float* gp = pipe.row(s, 0, b, tl);
          float* qp = pipe.row(s, 1, b, tl);
          float* tp = pipe.row(s, 2, b, tl);
          float* qdp = pipe.row(s, 3, b, tl);
          const float rt = gp[k], zt = gp[H + k], nt = gp[2 * H + k], qt = qp[k];
          const float art = tp[k], azt = tp[H + k], ant = tp[2 * H + k], qdt = qdp[k];
          const float hp = (t > 0) ? pipe.row(s, 4, b, tl)[k] : 0.f;
          const float hdp = (t > 0) ? pipe.row(s, 5, b, tl)[k] : 0.f;
          float hb, hdb;
          if (p.last_only) {
            hb = (t == T - 1) ? p.hbar[(size_t)(b0 + b) * H + k] : 0.f;
            hdb = (t == T - 1) ? p.hdbar[(size_t)(b0 + b) * H + k] : 0.f;
          } else {
            hb = pipe.row(s, 6, b, tl)[k];
            hdb = pipe.row(s, 7, b, tl)[k];
          }
          hb += ch[o];
          hdb += chd[o];
          const float sr = rt * (1.f - rt), sz = zt * (1.f - zt), sn = 1.f - nt * nt;
          const float rdot = sr * art, zdot = sz * azt, ndot = sn * ant;
          // hdot_t = (1-z) ndot + z hdot_{t-1} + zdot (h_{t-1} - n)
          const float ndb = (1.f - zt) * hdb;
          float zb = hdb * (hdp - ndot);
          const float zdb = hdb * (hp - nt);
          float nb_ = -zdot * hdb;
          nh[o] = zdot * hdb;
          nhd[o] = zt * hdb;
          // h_t = n + z (h_{t-1} - n)
          nb_ += (1.f - zt) * hb;
          zb += hb * (hp - nt);
          nh[o] += zt * hb;
          // ndot = (1-n^2) a_n
          const float anb_d = sn * ndb;
          nb_ -= 2.f * nt * ant * ndb;
          // a_n = gid_n + rdot q + r qdot
          const float rdb = qt * anb_d;
          float qb = rdot * anb_d;
          float rb = qdt * anb_d;
          const float qdb = rt * anb_d;
          // n = tanh(gi_n + r q)
          const float anb = sn * nb_;
          rb += qt * anb;
          qb += rt * anb;
          // zdot = sz a_z ; rdot = sr a_r
          const float azb_d = sz * zdb;
          zb += (1.f - 2.f * zt) * azt * zdb;
          const float arb_d = sr * rdb;
          rb += (1.f - 2.f * rt) * art * rdb;
          const float azb = sz * zb;
          const float arb = sr * rb;
          gp[k] = arb; gp[H + k] = azb; gp[2 * H + k] = anb; qp[k] = qb;
          tp[k] = arb_d; tp[H + k] = azb_d; tp[2 * H + k] = anb_d; qdp[k] = qdb;
          float* d0 = dg + (b * 6) * HR;
          d0[0 * HR + k] = arb;   d0[1 * HR + k] = azb;   d0[2 * HR + k] = qb;
          d0[3 * HR + k] = arb_d; d0[4 * HR + k] = azb_d; d0[5 * HR + k] = qdb;
        }
      }
      if (tl == 0 && pipe.bulk) fence_async_smem();
      __syncthreads();
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
            const float4 dv = reinterpret_cast<const float4*>(dg + (b * 6 + g) * HR)[i * G + ql];
            const float4 ev = reinterpret_cast<const float4*>(dg + (b * 6 + 3 + g) * HR)[i * G + ql];
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
  auto kern = gru_jvp_fwd_kernel<HP, G, BT, TC, NST>;
  TG_OPT_IN_SMEM(kern, "gru_jvp_fwd");
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
  auto kern = gru_jvp_bwd_kernel<HP, G, BT, TC, NST>;
  TG_OPT_IN_SMEM(kern, "gru_jvp_bwd");
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

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

template <int HP, int G, int BT, int TC, int NST>
__global__ void __launch_bounds__(HP* G, bwd_min_blocks<HP, G>()) gru_bwd_kernel(BwdParams p) {
  constexpr int KS = HP / G;
  constexpr int NOWN = (BT >= G) ? BT / G : 1;
  constexpr int HR = HP + DG_PAD;
  static_assert(KS % 4 == 0, "slice must be float4 granular");
  static_assert(G == 2 || G == 4, "lane groups of 2 or 4");
  static_assert(BT < G || BT % G == 0, "BT must be < G or a multiple of G");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int k = tid / G, ql = tid % G;
  const int H = p.H, T = p.T;
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
  // carry[o]: dL/dh_t[k] flowing in from step t+1 for the sequence this lane owns (b = o*G + ql, or b = ql)
  float carry[NOWN];
#pragma unroll
  for (int o = 0; o < NOWN; ++o) carry[o] = 0.f;
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
      float* dg = dgs + par * BT * 3 * HR;
      float cz[NOWN];
      // ---- pointwise gate derivatives for hidden unit k of the owned sequences ----
#pragma unroll
      for (int o = 0; o < NOWN; ++o) {
        const int b = (BT < G) ? ql : o * G + ql;
        cz[o] = 0.f;
        if (k < H && b < nb) {
          float* gp = pipe.row(s, 0, b, tl);
          float* qp = pipe.row(s, 1, b, tl);
          const float r = gp[k], z = gp[H + k], n = gp[2 * H + k], qv = qp[k];
          const float hp = (t > 0) ? pipe.row(s, 3, b, tl)[k] : 0.f;
          float dyv;
          if (p.dy_last) dyv = (t == T - 1) ? p.dy[(size_t)(b0 + b) * H + k] : 0.f;
          else dyv = pipe.row(s, 2, b, tl)[k];
          const float dh = dyv + carry[o];
          const float dn = dh * (1.f - z);
          const float dz = dh * (hp - n);
          const float dan = dn * (1.f - n * n);
          const float daz = dz * z * (1.f - z);
          const float dar = dan * qv * r * (1.f - r);
          const float dqv = dan * r;
          cz[o] = dh * z;
          gp[k] = dar; gp[H + k] = daz; gp[2 * H + k] = dan; qp[k] = dqv;
          dg[(b * 3 + 0) * HR + k] = dar;
          dg[(b * 3 + 1) * HR + k] = daz;
          dg[(b * 3 + 2) * HR + k] = dqv;
        }
      }
      if (tl == 0 && pipe.bulk) fence_async_smem();
      __syncthreads();
      // ---- carry_k = dh*z + sum_rows dGH[row] * W_hh[row][k] ----
      // three independent packed accumulators per sequence (one per gate) keep the FFMA2 chains short
      float2 acc2[BT][3];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc2[b][0] = acc2[b][1] = acc2[b][2] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KS / 4; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const float4 dv = reinterpret_cast<const float4*>(dg + (b * 3 + g) * HR)[i * G + ql];
            acc2[b][g] = __ffma2_rn(wt[g][2 * i + 0], make_float2(dv.x, dv.y), acc2[b][g]);
            acc2[b][g] = __ffma2_rn(wt[g][2 * i + 1], make_float2(dv.z, dv.w), acc2[b][g]);
          }
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b)
        acc[b] = (acc2[b][0].x + acc2[b][0].y) + (acc2[b][1].x + acc2[b][1].y) + (acc2[b][2].x + acc2[b][2].y);
      // reduce-scatter over the lane group: the owner of (k, b) receives the complete sum
      if constexpr (BT < G) {
        float mine = 0.f;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float v = group_sum<G>(acc[b]);
          mine = (ql == b) ? v : mine;
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

template <int HP, int G, int BT, int TC, int NST>
int launch_bwd(cudaStream_t st, const BwdParams& p) {
  const int widths[4] = {3 * p.H, p.H, p.H, p.H};
  constexpr int HR = HP + DG_PAD;
  size_t smem = ((2 * BT * 3 * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<4, BT, TC, NST>::stage_floats_for(widths) * 4;
  auto kern = gru_bwd_kernel<HP, G, BT, TC, NST>;
  TG_OPT_IN_SMEM(kern, "gru_bwd");
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_bwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  dim3 grid((p.B + BT - 1) / BT), block(HP * G);
  kern<<<grid, block, smem, st>>>(p);
  return tg_check_launch("gru_bwd");
}

template <int HP, int G>
int dispatch_bt(cudaStream_t st, const BwdParams& p, int bt) {
  constexpr int TC = (HP >= 128) ? 4 : 8, NST = 3;
  switch (bt) {
    case 1: return launch_bwd<HP, G, 1, TC, NST>(st, p);
    case 2: return launch_bwd<HP, G, 2, TC, NST>(st, p);
    case 4: return launch_bwd<HP, G, 4, TC, NST>(st, p);
  }
  tg_set_error("gru_bwd: bad BT %d", bt);
  return TG_ERR_ARG;
}

}  // namespace

int tg_gru_bwd_impl(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y,
                    const float* whh, float* dgi, float* dq, int B, int T, int H, int flags, const float* whh_t) {
  TG_REQUIRE(dy && rzn && q && y && whh && dgi && dq, TG_ERR_ARG, "gru_bwd: null pointer");
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "gru_bwd: bad shape B=%d T=%d H=%d", B, T, H);
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

// Persistent fused GRU layer forward for sm_100a (north_star kernel (1)).
//
// Replaces the per-timestep loop inside torch.nn.GRU reached from timeGAN/timegan_model.py:32-34
// (GRUStack.forward) -- math of SURVEY.md Appendix A.1, gate order r,z,n, h0 = 0.
//
// One CTA owns BT whole sequences for all T steps.
//   * W_hh lives in registers: the G lanes of a lane group share hidden unit j, lane q holds the k-slice
//     {(i*G+q)*4 .. +3} of the three gate rows r_j, z_j, n_j  (3*HP/G registers).
//   * h_{t-1} lives in a double-buffered shared-memory vector read as broadcast LDS.128; the copy a lane needs
//     for the state update stays in a register.
//   * the per-lane partial dot products are combined with a warp-shuffle REDUCE-SCATTER: after log2(G)
//     exchange rounds lane q holds the complete gate pre-activations of the sequences b == q (mod G), so every
//     lane evaluates sigmoid/tanh for a different (j, b) -- no idle lanes in the transcendental part.
//   * gi = X W_ih^T + b_ih (B,T,3H) streams in through a bulk-async (TMA engine) ring of TC-step chunks; y and,
//     when the backward pass will follow, r,z,n (over gi, in place) and q = h W_hn^T + b_hn stream out the
//     same way.  Exactly one __syncthreads per timestep.
#include "chunk_pipe.cuh"
#include "kernels.h"

namespace {

struct FwdParams {
  float* gi;         // (B,T,3H) in: x W_ih^T + b_ih ; out (if save): r,z,n
  const float* whh;  // (3H,H)
  const float* bhh;  // (3H)
  float* y;          // (B,T,H)
  float* q;          // (B,T,H) out (if save): h_{t-1} W_hn^T + b_hn
  int B, T, H;
  int save;
  int bulk;
  // bf16-gi mode (GIB instantiations): gi16 (B,T,3H) bf16 is read-only, r,z,n go to their own fp32 tensor
  const unsigned short* gi16;
  float* rzn;
};

constexpr int HS_PAD = 16;  // h rows are HP+16 floats apart: the two sequences a lane pair writes hit different banks

template <int HP, int G, int BT>
constexpr int fwd_min_blocks() { return (HP * G <= 128) ? 3 : ((HP * G <= 256) ? 2 : 1); }

// EXACT: H == HP is a compile-time fact (c2: 64, c3: 128) -- every stride in the hot loop becomes an immediate and
// the "is this hidden unit real" predicate disappears.
// GIB: the input projection arrives as bf16 (proj_bf16.cu) -- 6H instead of 12H bytes per cell on the way in; the saved
// r,z,n stay fp32 (BPTT reads them unchanged) and therefore get their own stream instead of overwriting gi.
template <int HP, int G, int BT, int TC, int NST, bool EXACT, bool GIB = false>
__global__ void __launch_bounds__(HP* G, (EXACT || fwd_min_blocks<HP, G, BT>() == 1) ? fwd_min_blocks<HP, G, BT>() : fwd_min_blocks<HP, G, BT>() - 1) gru_fwd_kernel(FwdParams p) {
  constexpr int KS = HP / G;                     // k values per lane
  constexpr int NOWN = (BT >= G) ? BT / G : 1;   // sequences a lane finishes per step
  constexpr int HR = HP + HS_PAD;
  constexpr bool DUAL = (HP * G <= 128);         // second accumulator set only where the register budget allows
  static_assert(KS % 4 == 0, "k-slice must be float4 granular");
  static_assert(G == 2 || G == 4, "lane groups of 2 or 4");
  static_assert(BT < G || BT % G == 0, "BT must be < G or a multiple of G");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int j = tid / G, ql = tid % G;
  const int H = EXACT ? HP : p.H;
  const int T = p.T;
  const int b0 = blockIdx.x * BT;
  const int nb = min(BT, p.B - b0);

  // ---- shared memory carve-up ----
  float* hs = reinterpret_cast<float*>(smem_raw);                 // [2][BT][HR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + 2 * BT * HR);  // [NST]
  float* stages = reinterpret_cast<float*>(smem_raw + ((2 * BT * HR * 4 + NST * 8 + 127) / 128) * 128);

  ChunkPipe<GIB ? 4 : 3, BT, TC, NST> pipe;
  if constexpr (GIB) {
    // stream 0: 3H bf16 = 3H/2 floats per row, load only; stream 3: r,z,n fp32, store only
    pipe.g[0] = pipe.gst[0] = reinterpret_cast<float*>(const_cast<unsigned short*>(p.gi16));
    pipe.w[0] = 3 * H / 2; pipe.mode[0] = TG_STRM_LOAD; pipe.shift[0] = 0;
    pipe.g[3] = pipe.gst[3] = p.rzn; pipe.w[3] = 3 * H; pipe.mode[3] = p.save ? TG_STRM_STORE : 0; pipe.shift[3] = 0;
  } else {
    pipe.g[0] = pipe.gst[0] = p.gi; pipe.w[0] = 3 * H; pipe.mode[0] = TG_STRM_LOAD | (p.save ? TG_STRM_STORE : 0); pipe.shift[0] = 0;
  }
  pipe.g[1] = pipe.gst[1] = p.q;  pipe.w[1] = H;     pipe.mode[1] = p.save ? TG_STRM_STORE : 0;                  pipe.shift[1] = 0;
  pipe.g[2] = pipe.gst[2] = p.y;  pipe.w[2] = H;     pipe.mode[2] = TG_STRM_STORE;                               pipe.shift[2] = 0;
  pipe.layout();
  pipe.stages = stages; pipe.full = bars;
  pipe.T = T; pipe.nb = nb; pipe.b0 = b0; pipe.NC = (T + TC - 1) / TC;
  pipe.reverse = false; pipe.bulk = p.bulk != 0;

  // ---- W_hh slice into registers (interleaved float4 k-assignment => conflict-free LDS.128 of h) ----
  // stored as float2 pairs: the mat-vec runs on packed FFMA2 (two k values per instruction, sm_100)
  float2 w[3][KS / 2];
  float bh[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    bh[g] = (j < H) ? p.bhh[g * H + j] : 0.f;
#pragma unroll
    for (int i = 0; i < KS / 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int k = (i * G + ql) * 4 + c;
        const float v = (j < H && k < H) ? p.whh[(size_t)(g * H + j) * H + k] : 0.f;
        if (c & 1) w[g][2 * i + (c >> 1)].y = v; else w[g][2 * i + (c >> 1)].x = v;
      }
  }
  for (int i = tid; i < 2 * BT * HR; i += HP * G) hs[i] = 0.f;
  float hprev[NOWN];
#pragma unroll
  for (int o = 0; o < NOWN; ++o) hprev[o] = 0.f;
  pipe.start();
  __syncthreads();

  // per-thread constants of the hot loop: which sequences this lane finishes, shared-window addresses
  const uint32_t hs_addr = smem_u32(hs);
  bool act[NOWN];
  int ob[NOWN];
#pragma unroll
  for (int o = 0; o < NOWN; ++o) {
    // BT < G: the butterfly leaves every total in every lane, so lane q finishes sequence q % BT -- the surplus
    // lanes repeat a sibling's work and stores (same values, same addresses) instead of idling behind a branch
    ob[o] = (BT < G) ? ql % BT : o * G + ql;
    // one sequence per CTA and H == HP: every lane is live -- a compile-time fact, so the stores need no branch
    act[o] = (EXACT && BT == 1) || ((EXACT || j < H) && (ob[o] < nb));
  }
  const uint32_t gi_step = (GIB ? 6u : 12u) * (uint32_t)H, h_step = 4u * (uint32_t)H;
  const uint32_t gate_step = GIB ? 2u * (uint32_t)H : h_step;     // bytes between the r, z, n blocks of an input row
  const uint32_t lane_k = 16u * (uint32_t)ql;          // this lane's first float4 of h
  const bool save = p.save != 0;
  // the lane that owns sequence b seeds b's accumulators with the hidden biases of the r and z gates: the
  // reduce-scatter then delivers (W_h h + b_h) and two additions leave the per-step dependency chain
  float seed_r[BT], seed_z[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    const bool mine = (BT < G) ? (ql == b) : ((b % G) == ql);
    seed_r[b] = mine ? bh[0] : 0.f;
    seed_z[b] = mine ? bh[1] : 0.f;
  }

  int cur = 0;
  uint32_t a_gi[NOWN], a_q[NOWN], a_y[NOWN];
  uint32_t a_rz[GIB ? NOWN : 1];       // GIB: where r,z,n are staged (their own stream); otherwise they overwrite gi

  // One timestep.  Everything is computed unconditionally (inactive lanes work on row 0 of the stage and never
  // store), so the body is branch-free: the warps do not pay for BSSY/BSYNC reconvergence every step.
  auto step = [&]() __attribute__((always_inline)) {
    const uint32_t hc = hs_addr + (uint32_t)(cur * BT * HR) * 4u;
    const uint32_t hn = hs_addr + (uint32_t)((cur ^ 1) * BT * HR) * 4u;
    // the input-projection values are fetched first so that their latency hides behind the mat-vec
    float gr[NOWN], gz[NOWN], gn[NOWN];
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      if constexpr (GIB) {
        gr[o] = lds_bf16(a_gi[o]); gz[o] = lds_bf16(a_gi[o] + gate_step); gn[o] = lds_bf16(a_gi[o] + 2u * gate_step);
      } else {
        gr[o] = lds_f32(a_gi[o]); gz[o] = lds_f32(a_gi[o] + h_step); gn[o] = lds_f32(a_gi[o] + 2u * h_step);
      }
    }
    // two accumulator pairs per (sequence, gate): six independent FFMA2 chains per sequence keep the FMA pipe
    // issuing every other cycle instead of waiting out the FFMA2 latency
    float2 accA[BT][3], accB[BT][3];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      accA[b][0] = make_float2(seed_r[b], 0.f);
      accA[b][1] = make_float2(seed_z[b], 0.f);
      accA[b][2] = make_float2(0.f, 0.f);
      accB[b][0] = accB[b][1] = accB[b][2] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < KS / 4; ++i) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 hv = lds_v4(hc + (uint32_t)(b * HR) * 4u + (uint32_t)(i * G) * 16u + lane_k);
        const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          accA[b][g] = __ffma2_rn(w[g][2 * i + 0], h01, accA[b][g]);
          if constexpr (DUAL) accB[b][g] = __ffma2_rn(w[g][2 * i + 1], h23, accB[b][g]);
          else accA[b][g] = __ffma2_rn(w[g][2 * i + 1], h23, accA[b][g]);
        }
      }
    }
    float acc[BT][3];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int g = 0; g < 3; ++g) acc[b][g] = (accA[b][g].x + accB[b][g].x) + (accA[b][g].y + accB[b][g].y);
    // ---- combine the G partial sums: own[o][g] = complete sum for sequence b = o*G + ql ----
    float own[NOWN][3];
    if constexpr (BT < G) {
      // fewer sequences than lanes: butterfly, lane b finishes sequence b
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int g = 0; g < 3; ++g) acc[b][g] = group_sum<G>(acc[b][g]);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        own[0][g] = acc[0][g];
#pragma unroll
        for (int b = 1; b < BT; ++b) own[0][g] = (ql % BT == b) ? acc[b][g] : own[0][g];
      }
    } else if constexpr (G == 2) {
#pragma unroll
      for (int o = 0; o < NOWN; ++o)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float send = ql ? acc[2 * o][g] : acc[2 * o + 1][g];
          const float keep = ql ? acc[2 * o + 1][g] : acc[2 * o][g];
          own[o][g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    } else {  // G == 4
      const int hi = ql & 2, lo = ql & 1;
#pragma unroll
      for (int o = 0; o < NOWN; ++o)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float s0 = hi ? acc[4 * o + 0][g] : acc[4 * o + 2][g];
          const float k0 = hi ? acc[4 * o + 2][g] : acc[4 * o + 0][g];
          const float s1 = hi ? acc[4 * o + 1][g] : acc[4 * o + 3][g];
          const float k1 = hi ? acc[4 * o + 3][g] : acc[4 * o + 1][g];
          const float a0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
          const float a1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
          const float send = lo ? a0 : a1;
          const float keep = lo ? a1 : a0;
          own[o][g] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
    }
    // ---- gates + state update for the sequences this lane owns ----
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      // BT < G: the butterfly leaves the same total (incl. the seeded bias) in every lane of the group
      const float r = sigmoid_mufu(gr[o] + own[o][0]);
      const float z = sigmoid_mufu(gz[o] + own[o][1]);
      const float qv = own[o][2] + bh[2];
      const uint32_t a_s = GIB ? a_rz[GIB ? o : 0] : a_gi[o];
      if (save && act[o]) { sts_f32(a_s, r); sts_f32(a_s + h_step, z); sts_f32(a_q[o], qv); }
      const float n = tanh_mufu(fmaf(r, qv, gn[o]));
      const float h = fmaf(z, hprev[o] - n, n);
      hprev[o] = h;
      if (act[o]) {
        sts_f32(hn + (uint32_t)(ob[o] * HR + j) * 4u, h);
        sts_f32(a_y[o], h);
        if (save) sts_f32(a_s + 2u * h_step, n);
      }
      a_gi[o] += gi_step; a_q[o] += h_step; a_y[o] += h_step;
      if constexpr (GIB) a_rz[o] += 3u * h_step;
    }
    cur ^= 1;
  };

  for (int c = 0; c < pipe.NC; ++c) {
    pipe.acquire(c);
    const int s = c % NST;
    const int tcn = pipe.tcn_of(c);
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
      const int b = act[o] ? ob[o] : 0;
      const int jj = act[o] ? j : 0;
      a_gi[o] = pipe.row_addr(s, 0, b, 0) + (GIB ? 2u : 4u) * (uint32_t)jj;
      if constexpr (GIB) a_rz[o] = pipe.row_addr(s, 3, b, 0) + 4u * (uint32_t)jj;
      a_q[o] = pipe.row_addr(s, 1, b, 0) + 4u * (uint32_t)jj;
      a_y[o] = pipe.row_addr(s, 2, b, 0) + 4u * (uint32_t)jj;
    }
    // the chunk's last step is peeled: only it needs the async-proxy fence before the stage is stored
    for (int tl = 0; tl < tcn - 1; ++tl) {
      step();
      __syncthreads();
    }
    step();
    if (pipe.bulk) fence_async_smem();
    __syncthreads();
    pipe.release(c);
  }
  pipe.drain();
}

// bf16-gi launcher (H == HP only: c2's 64 and c3's 128)
template <int HP, int G, int BT, int TC, int NST>
int launch_fwd_gib(cudaStream_t st, const FwdParams& p) {
  constexpr int HR = HP + HS_PAD;
  const int widths[4] = {3 * p.H / 2, p.H, p.H, 3 * p.H};
  size_t smem = ((2 * BT * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<4, BT, TC, NST>::stage_floats_for(widths) * 4;
  TG_OPT_IN_SMEM((gru_fwd_kernel<HP, G, BT, TC, NST, true, true>), "gru_fwd(bf16 gi)");
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_fwd(bf16 gi): needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  dim3 grid((p.B + BT - 1) / BT), block(HP * G);
  gru_fwd_kernel<HP, G, BT, TC, NST, true, true><<<grid, block, smem, st>>>(p);
  return tg_check_launch("gru_fwd(bf16 gi)");
}

template <int HP, int G, int BT, int TC, int NST>
int launch_fwd(cudaStream_t st, const FwdParams& p) {
  constexpr int HR = HP + HS_PAD;
  const int widths[3] = {3 * p.H, p.H, p.H};
  size_t smem = ((2 * BT * HR * 4 + NST * 8 + 127) / 128) * 128 +
                (size_t)NST * ChunkPipe<3, BT, TC, NST>::stage_floats_for(widths) * 4;
  const bool exact = (p.H == HP);
  auto kern = exact ? gru_fwd_kernel<HP, G, BT, TC, NST, true> : gru_fwd_kernel<HP, G, BT, TC, NST, false>;
  if (exact) { TG_OPT_IN_SMEM((gru_fwd_kernel<HP, G, BT, TC, NST, true>), "gru_fwd"); }
  else { TG_OPT_IN_SMEM((gru_fwd_kernel<HP, G, BT, TC, NST, false>), "gru_fwd"); }
  if (smem > (size_t)tg_max_optin_smem()) { tg_set_error("gru_fwd: needs %zu B of shared memory", smem); return TG_ERR_UNSUPPORTED; }
  dim3 grid((p.B + BT - 1) / BT), block(HP * G);
  kern<<<grid, block, smem, st>>>(p);
  return tg_check_launch("gru_fwd");
}

template <int HP, int G>
int dispatch_bt(cudaStream_t st, const FwdParams& p, int bt) {
  constexpr int TC = 8, NST = 3;   // 8-step chunks (H = 128 used 4: twice the per-chunk ring cost per step)
  // one sequence per CTA (the co-resident, latency-bound regime): 16-step chunks halve the per-chunk cost of the
  // bulk-copy ring (thread 0 issues the stores/loads while the other warps wait at the next barrier)
  if (HP <= 64 && bt == 1 && tg_long_chunks()) return launch_fwd<HP, G, 1, 2 * TC, NST>(st, p);
  switch (bt) {
    case 1: return launch_fwd<HP, G, 1, TC, NST>(st, p);
    case 2: return launch_fwd<HP, G, 2, TC, NST>(st, p);
    case 4: return launch_fwd<HP, G, 4, (HP >= 128) ? 4 : TC, NST>(st, p);   // four H=128 sequences x 8 steps x 3 stages exceed shared memory
  }
  tg_set_error("gru_fwd: bad BT %d", bt);
  return TG_ERR_ARG;
}

template <int HP, int G>
int dispatch_bt_gib(cudaStream_t st, const FwdParams& p, int bt) {
  constexpr int NST = 3;
  // the separate r,z,n stream makes a stage 6.5H floats per sequence-step (in place: 5H)
  switch (bt) {
    case 1: return launch_fwd_gib<HP, G, 1, (HP <= 64) ? 16 : 8, NST>(st, p);
    case 2: return launch_fwd_gib<HP, G, 2, 8, NST>(st, p);
    case 4: return launch_fwd_gib<HP, G, 4, (HP <= 64) ? 8 : 4, NST>(st, p);
  }
  tg_set_error("gru_fwd(bf16 gi): bad BT %d", bt);
  return TG_ERR_ARG;
}

}  // namespace

// Sequences per CTA.  The recurrent kernels are bound by the per-step dependency chain (about 550 clk even with
// no mat-vec work, measured with tools/probe_gru.py), so the fastest layout keeps every sequence in its own CTA
// as long as all CTAs are co-resident; beyond that, two (then four) sequences share a CTA's weight registers.
int tg_pick_bt(int B, int HP, int bt_override) {
  if (bt_override == 1 || bt_override == 2 || bt_override == 4) return bt_override;
  const int cap = tg_num_sms() * ((HP <= 64) ? 2 : 1);   // CTAs that run concurrently
  if (B <= cap) return 1;
  if (B <= 2 * cap) return 2;
  return 4;
}

int tg_gru_fwd_impl(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T,
                    int H, int flags) {
  TG_REQUIRE(gi && whh && bhh && y, TG_ERR_ARG, "gru_fwd: null pointer");
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "gru_fwd: bad shape B=%d T=%d H=%d", B, T, H);
  TG_REQUIRE(H <= 1024, TG_ERR_UNSUPPORTED, "gru_fwd: hidden size %d > 1024", H);
  const int save = (flags & TG_GRU_SAVE) ? 1 : 0;
  TG_REQUIRE(!save || q, TG_ERR_ARG, "gru_fwd: save requested without q buffer");
  if (tg_cluster_takes(H, B, false) && !(flags & TG_GRU_NO_BULK) && tg_aligned16(gi) && tg_aligned16(y) && tg_aligned16(whh) &&
      (!save || tg_aligned16(q)))
    return tg_gru_cl_fwd(st, gi, whh, bhh, y, q, B, T, H, save);
  if (H > 128) return tg_bigh_fwd(st, gi, whh, bhh, y, q, B, T, H, save);
  FwdParams p{gi, whh, bhh, y, q, B, T, H, save, 0, nullptr, nullptr};
  p.bulk = (H % 4 == 0) && tg_aligned16(gi) && tg_aligned16(y) && (!save || tg_aligned16(q)) &&
           !(flags & TG_GRU_NO_BULK);
  const int bto = (flags >> 8) & 0xff;
  if (H <= 32) return dispatch_bt<32, 2>(st, p, tg_pick_bt(B, 32, bto));
  if (H <= 64) return dispatch_bt<64, 2>(st, p, tg_pick_bt(B, 64, bto));
  return dispatch_bt<128, 4>(st, p, tg_pick_bt(B, 128, bto));
}

bool tg_gru_fwd_bf16gi_ok(int H) { return H == 64 || H == 128; }

// forward pass fed by the bf16 projection (proj_bf16.cu): gi16 (B,T,3H) bf16 in; y, and with TG_GRU_SAVE rzn (B,T,3H)
// fp32 and q out
int tg_gru_fwd_bf16gi_impl(cudaStream_t st, const void* gi16, const float* whh, const float* bhh, float* y, float* q,
                           float* rzn, int B, int T, int H, int flags) {
  TG_REQUIRE(gi16 && whh && bhh && y, TG_ERR_ARG, "gru_fwd(bf16 gi): null pointer");
  TG_REQUIRE(B > 0 && T > 0, TG_ERR_SHAPE, "gru_fwd(bf16 gi): bad shape B=%d T=%d", B, T);
  if (!tg_gru_fwd_bf16gi_ok(H)) { tg_set_error("gru_fwd(bf16 gi): hidden size %d not instantiated (64, 128)", H); return TG_ERR_UNSUPPORTED; }
  const int save = (flags & TG_GRU_SAVE) ? 1 : 0;
  TG_REQUIRE(!save || (q && rzn), TG_ERR_ARG, "gru_fwd(bf16 gi): save requested without q / rzn buffers");
  TG_REQUIRE(tg_aligned16(gi16) && tg_aligned16(y) && (!save || (tg_aligned16(q) && tg_aligned16(rzn))), TG_ERR_ALIGN,
             "gru_fwd(bf16 gi): buffers must be 16-byte aligned");
  FwdParams p{nullptr, whh, bhh, y, q, B, T, H, save, 1, reinterpret_cast<const unsigned short*>(gi16), rzn};
  const int bto = (flags >> 8) & 0xff;
  if (H == 64) return dispatch_bt_gib<64, 2>(st, p, tg_pick_bt(B, 64, bto));
  return dispatch_bt_gib<128, 4>(st, p, tg_pick_bt(B, 128, bto));
}

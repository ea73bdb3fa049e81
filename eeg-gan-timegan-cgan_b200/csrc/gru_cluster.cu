// Resident GRU recurrences for hidden sizes whose W_hh does not fit ONE SM's register file: thread-block CLUSTERS.
//
//   H = 128 (BASELINE config c3):  W_hh = 196 KB  ->  cluster of 2 CTAs, each holds the 3 x 64 gate rows of 64 hidden units
//   H = 256 (config c4's top end): W_hh = 786 KB  ->  cluster of 8 CTAs, 32 hidden units each
// Every CTA has 256 threads with 96 weight registers each (98 KB of W_hh per SM, resident for all T steps), so the
// kernels run one CTA per SM with a 255-register budget: no spills (the one-SM H = 128 kernels of gru_fwd.cu / gru_bwd.cu
// are capped at 128 registers by their 512 threads and spill in the hot loop; H = 256 used to stream W_hh from L2 every
// step, gru_bigh.cu).  Replaces the same reference code: the nn.GRU time loop of timegan_model.py:32-34 and autograd's
// backward of it (train_timegan.py:140,159,219,267) -- SURVEY.md A.1 / A.2.
//
// Per timestep the CTAs of a cluster ALL-GATHER the new state through distributed shared memory: the lane that owns
// (hidden unit j, sequence b) writes h_t[b][j] (BPTT: the three dGH values) into every CTA's double-buffered state
// vector with `st.async ... mbarrier::complete_tx::bytes`, i.e. the remote store itself signals the receiving CTA's
// mbarrier -- no cluster-wide barrier, no flag polling, no fence; a CTA waits on its OWN mbarrier (hardware-suspended
// try_wait) until the H*G values of the step have landed.  The values of four consecutive hidden units are first
// gathered into one lane by warp shuffles so that a store carries 16 bytes: 64 x CS stores per CTA and step.
// (Measured at H = 128, B = 256, forward: one 4-byte st.async per value 2265 clk/step; plain st.shared::cluster stores +
// one release-arrive per warp 3031 clk/step -- the cluster-scope release waits for the remote stores' acknowledgements.)
// Two barriers (step parity) per sequence group suffice: a CTA can be at most one step ahead of its slowest peer,
// because it needs that peer's values to finish a step.
//
// Forward lane mapping: G = 256/HU lanes per hidden unit, lane q holds the k-slice {(i*G+q)*4..+3}; a group of G
// sequences is processed together and the G partial sums are combined by a shuffle reduce-scatter, after which lane q
// owns sequence q (every lane evaluates the gates of a different (j, b)).
// BPTT lane mapping: TWO outputs per thread and 2G lanes per output pair, so that every LDS.128 of the (3H long) dGH
// vector feeds four FFMA2 instead of two -- the shared-memory pipe, not the FMA pipe, is what the 3x longer input vector
// of the backward mat-vec saturates otherwise; the 2G partial sums of 2 outputs x G sequences are reduce-scattered over
// the 2G lanes (each lane ends up owning one (k, b)).
//
// HBM traffic goes through shared memory in both directions: the per-step inputs (gi, or r,z,n,q,h_{t-1},dy) are
// prefetched PF steps ahead with cp.async (16-byte pieces of coalesced 128/256-byte row segments) into a ring, the
// outputs are staged and written as coalesced float4 rows right after the step's barrier.  Every thread owns a FIXED
// set of at most two load and two store items whose addresses are computed once, so a step's I/O is a handful of
// instructions (recomputing (sequence, array, column) from the thread index every step cost a third of the step in the
// first version; a ninth, dedicated I/O warp was tried too: 288 threads are allocated like 384, which caps the compute
// threads at 168 registers and spills the weight slices).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int CL_THREADS = 256;
constexpr int CL_PF = 4;          // prefetch distance (timesteps) of the cp.async input ring
// Experiment kept as a compile-time switch (off): issuing a step's HBM traffic from inside the first group's mat-vec so
// that it rides in the FMA pipe's free issue slots.  Measured SLOWER on a B200 (H = 128, B = 256: forward 765 -> 881 us,
// BPTT 1058 -> 1123 us per pass): the extra shared-memory reads and global stores land in the middle of the operand
// stream the FFMA2s are waiting on.
#ifndef TG_CL_IO_IN_MATVEC
#define TG_CL_IO_IN_MATVEC 0
#endif
constexpr bool CL_IO_IN_MATVEC = TG_CL_IO_IN_MATVEC != 0;
constexpr int CL_HPAD = 16;       // state rows are H+16 floats apart (bank spread between the sequences of a lane group)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t laddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
  return r;
}
// 16-byte store into a cluster CTA's shared memory (own CTA included) that completes 16 bytes of that CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(raddr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)),
               "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_susp(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
  }
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// reduce-scatter of N per-lane partials over N consecutive lanes: afterwards v[0] of lane q (q = lane % N) is the total
// of index q
template <int N>
__device__ __forceinline__ void reduce_scatter(float (&v)[N], int q) {
#pragma unroll
  for (int s = N / 2; s >= 1; s >>= 1) {
    const bool up = (q & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// N lanes, M <= N partials per lane (M, N powers of two): afterwards v[0] of lane q is the total of index q % M (every
// total is held by N/M lanes): log2(M) halving exchange rounds, then log2(N/M) butterfly adds
template <int N, int M>
__device__ __forceinline__ void reduce_partial(float (&v)[M], int q) {
#pragma unroll
  for (int s = M / 2; s >= 1; s >>= 1) {
    const bool up = (q & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
#pragma unroll
  for (int s = M; s < N; s <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], s);
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct ClFwdParams {
  float* gi;         // (B,T,3H) in: x W_ih^T + b_ih ; out (if save): r,z,n
  const float* whh;  // (3H,H)
  const float* bhh;  // (3H)
  float* y;          // (B,T,H)
  float* q;          // (B,T,H) out (if save)
  int B, T, save;
};

// SG = sequences per group (G or G/2), NGRP groups per cluster, each with its own pair of barriers
template <int H, int CS, int NGRP, int SG>
struct ClFwdSmem {
  static constexpr int HU = H / CS, G = CL_THREADS / HU, BT = SG * NGRP, HR = H + CL_HPAD;
  static constexpr int HBUF = 2 * BT * HR;             // floats
  // Per-sequence blocks of the input ring / output staging are padded by 32/G floats: the lanes of a warp that work on
  // the same column of DIFFERENT sequences (G of them) would otherwise hit one bank -- (arrays x HU) is a multiple of 32
  // floats -- a G-way conflict on every per-step load of the saved activations and every staging store (ncu counted
  // 26.8 M shared-memory bank conflicts per BPTT launch at H = 128, against ~1 M in the one-SM kernels).
  static constexpr int SPAD = 32 / G, RSEQ = 3 * HU + SPAD, SSEQ = 5 * HU + SPAD;
  static constexpr int RING = CL_PF * BT * RSEQ;       // floats
  static constexpr int STG = 2 * BT * SSEQ;            // floats
  static constexpr size_t bytes = (size_t)(HBUF + RING + STG) * 4 + 128;
};

// NH = hidden units per thread in the mat-vec (1 or 2): with NH = 2 the NH*G lanes of a unit pair each hold H/(NH*G) columns
// of both units' gate rows, so one LDS.128 of h feeds 12 FFMA2 instead of 6 (half the operand fetches); the partial sums of
// 2 units x G sequences are reduce-scattered over the 2G lanes, which leaves every lane with the totals of the SAME (unit,
// sequence) it owns with NH = 1 -- gates, all-gather and staging are unchanged.
template <int H, int CS, int NGRP, int SG, int NH = 1>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(CL_THREADS, 1) gru_cl_fwd_kernel(ClFwdParams p) {
  using S = ClFwdSmem<H, CS, NGRP, SG>;
  constexpr int HU = S::HU, G = S::G, BT = S::BT, HR = S::HR, L2 = NH * G, KS = H / L2;
  static_assert(NH * 3 * KS == 96 && (NH == 1 || (NH == 2 && SG == G)), "96 weight registers per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);            // [NGRP][2]
  float* hbuf = reinterpret_cast<float*>(smem_raw + 128);            // [2][BT][HR]
  float* ring = hbuf + S::HBUF;                                      // [PF][BT][3][HU]
  float* stg = ring + S::RING;                                       // [2][BT][5][HU]
  const int tid = threadIdx.x, jl = tid / G, ql = tid % G;
  const int qk = tid % L2, uo = qk / G;            // k-slice lane inside the unit group; which unit of the group this lane owns
  const uint32_t rank = cluster_ctarank();
  const int j = (int)rank * HU + jl;
  const int T = p.T;
  const int b0 = (blockIdx.x / CS) * BT;
  const bool save = p.save != 0;

  for (int i = tid; i < S::HBUF + S::RING; i += CL_THREADS) hbuf[i] = 0.f;     // h_{-1} = 0; ring rows of absent sequences
  static_assert(SG == G || SG * 2 == G, "a group is G or G/2 sequences");
  constexpr uint32_t TXB = (uint32_t)(H * SG * 4);                  // bytes one group receives per step
  if (tid == 0) {
    for (int i = 0; i < 2 * NGRP; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    for (int i = 0; i < 2 * NGRP; ++i) mbar_expect_tx(&bars[i], TXB);     // arm phase 0 of every barrier
  }
  __syncthreads();

  // ---- this thread's share of the step's HBM traffic: fixed items, addresses computed once ----
  constexpr int Q = HU / 4;                               // float4 per row segment
  constexpr int NLI = BT * 3 * Q, NSI = BT * 5 * Q;       // load / store items per step
  constexpr int NLD = (NLI + CL_THREADS - 1) / CL_THREADS, NSD = (NSI + CL_THREADS - 1) / CL_THREADS;
  constexpr int RSEQ = S::RSEQ, SSEQ = S::SSEQ;
  constexpr uint32_t SLOT_BYTES = (uint32_t)(BT * RSEQ) * 4u, STG_BYTES = (uint32_t)(BT * SSEQ) * 4u;
  const float* lp[NLD]; uint32_t ls[NLD]; bool lv[NLD];
#pragma unroll
  for (int m = 0; m < NLD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, g = (n / Q) % 3, b = n / (3 * Q);
    lv[m] = (n < NLI) && (b0 + b < p.B);
    lp[m] = p.gi + (size_t)(b0 + (lv[m] ? b : 0)) * T * (3 * H) + g * H + (int)rank * HU + j4 * 4;
    ls[m] = smem_u32(ring) + (uint32_t)(b * RSEQ + g * HU + j4 * 4) * 4u;
  }
  float* sp[NSD]; uint32_t ss[NSD]; bool sv[NSD]; int sst[NSD];
#pragma unroll
  for (int m = 0; m < NSD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, wch = (n / Q) % 5, b = n / (5 * Q);
    sv[m] = (n < NSI) && (b0 + b < p.B) && (save || wch == 4);
    const size_t seq = (size_t)(b0 + (sv[m] ? b : 0)) * T;
    float* base = (wch < 3) ? p.gi + seq * (3 * H) + wch * H : (wch == 3 ? p.q + seq * H : p.y + seq * H);
    sp[m] = base + (int)rank * HU + j4 * 4;
    sst[m] = (wch < 3) ? 3 * H : H;
    ss[m] = smem_u32(stg) + (uint32_t)(b * SSEQ + wch * HU + j4 * 4) * 4u;
  }
  auto prefetch = [&](int t) {
    if (t < T) {
      const uint32_t so = (uint32_t)(t % CL_PF) * SLOT_BYTES;
#pragma unroll
      for (int m = 0; m < NLD; ++m)
        if (lv[m]) cp_async16(ls[m] + so, lp[m] + (size_t)t * (3 * H));
    }
    cp_async_commit();
  };
  auto store = [&](int t) {
    const uint32_t so = (uint32_t)(t & 1) * STG_BYTES;
#pragma unroll
    for (int m = 0; m < NSD; ++m)
      if (sv[m]) *reinterpret_cast<float4*>(sp[m] + (size_t)t * sst[m]) = lds_v4(ss[m] + so);
  };
  for (int t = 0; t < CL_PF - 1; ++t) prefetch(t);

  // =============================== compute warps ===============================
  // ---- W_hh slice into registers (float2 pairs for FFMA2) ----
  float2 w[NH][3][KS / 2];
  float bh[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    bh[g] = p.bhh[g * H + j];
#pragma unroll
    for (int u = 0; u < NH; ++u)
#pragma unroll
      for (int i = 0; i < KS / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(p.whh + (size_t)(g * H + j - uo + u) * H + (i * L2 + qk) * 4);
        w[u][g][2 * i] = make_float2(v.x, v.y);
        w[u][g][2 * i + 1] = make_float2(v.z, v.w);
      }
  }
  cluster_sync_all();      // every CTA's barriers are initialised and armed before anybody sends

  // Senders: the lanes whose hidden unit is the first of a run of four (jl % 4 == 0) collect the run's values of THEIR
  // sequence from the three sibling lanes (same ql, jl+1..3) and store 16 bytes into every CTA's state vector.
  // The four lanes of a run all end up with the run's four values; lane i of the run serves the destination CTAs
  // c == i (mod 4), so the stores to different CTAs leave in ONE warp instruction instead of CS consecutive ones.
  const int lane = tid & 31;
  const int src0 = lane & ~(3 * G);                 // lane of (first unit of the run, same sequence)
  constexpr int ND = (CS + 3) / 4;                  // destinations per lane
  uint32_t r_h[ND], r_bar[ND];
  bool r_ok[ND];
#pragma unroll
  const int ob = ql % SG;                           // the sequence (within a group) this lane finishes
  const bool primary = ql < SG;                     // SG < G: lanes ql and ql + SG hold the same totals; one of them stores
  for (int d = 0; d < ND; ++d) {
    const int c = (jl & 3) + 4 * d;
    r_ok[d] = primary && c < CS;
    r_h[d] = mapa_shared(smem_u32(hbuf + ob * HR + (j & ~3)), (uint32_t)(c < CS ? c : 0));
    r_bar[d] = mapa_shared(smem_u32(bars), (uint32_t)(c < CS ? c : 0));
  }
  float hprev[NGRP];
#pragma unroll
  for (int g = 0; g < NGRP; ++g) hprev[g] = 0.f;
  const uint32_t hbuf_a = smem_u32(hbuf);

  for (int t = 0; t < T; ++t) {
    cp_async_wait<CL_PF - 2>();          // this thread's share of step t's gi rows has landed
    __syncthreads();                     // ... and everybody else's; also closes step t-1's staging writes
    // The step's HBM traffic (outputs of step t-1 as coalesced float4 rows, cp.async of step t+PF-1's inputs) is issued
    // from INSIDE the first group's mat-vec: FFMA2 occupies the scheduler's issue port every other cycle only, so the
    // loads / stores / address arithmetic ride in the free slots instead of standing between two barriers.
    if (!CL_IO_IN_MATVEC) {
      if (t > 0) store(t - 1);
      prefetch(t + CL_PF - 1);
    }
    const int par = t & 1, ppar = par ^ 1;
    float* sgw = stg + par * (BT * SSEQ);
    const float* ringt = ring + (t % CL_PF) * (BT * RSEQ);
    // everything after a group's mat-vec: combine the lane partials, gates, state update, all-gather, staging
    auto post = [&](const int grp, float2 (&acc)[NH][SG][3]) __attribute__((always_inline)) {
      const float* gr_ = ringt + (grp * SG + ob) * RSEQ + jl;
      const float gr = gr_[0], gz = gr_[HU], gn = gr_[2 * HU];
      float own[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float v[NH * SG];
#pragma unroll
        for (int u = 0; u < NH; ++u)
#pragma unroll
          for (int b = 0; b < SG; ++b) v[u * SG + b] = acc[u][b][g].x + acc[u][b][g].y;
        if constexpr (NH == 1) reduce_partial<G, SG>(v, ql);
        else reduce_scatter<L2>(v, qk);
        own[g] = v[0];
      }
      const float r = sigmoid_mufu(gr + own[0]);
      const float z = sigmoid_mufu(gz + own[1]);
      const float qv = own[2] + bh[2];
      const float n = tanh_mufu(fmaf(r, qv, gn));
      const float h = fmaf(z, hprev[grp] - n, n);
      hprev[grp] = h;
      // all-gather: h_t[b][j] into every CTA's state vector (parity t&1), signalling that CTA's barrier
      if (t + 1 < T) {
        const uint32_t off = (uint32_t)((par * BT + grp * SG) * HR) * 4u;
        const float h0 = __shfl_sync(0xffffffffu, h, src0), h1 = __shfl_sync(0xffffffffu, h, src0 + G),
                    h2 = __shfl_sync(0xffffffffu, h, src0 + 2 * G), h3 = __shfl_sync(0xffffffffu, h, src0 + 3 * G);
#pragma unroll
        for (int d = 0; d < ND; ++d)
          if (r_ok[d]) st_async_v4(r_h[d] + off, h0, h1, h2, h3, r_bar[d] + (uint32_t)(grp * 2 + par) * 8u);
      }
      if (primary) {
        float* so = sgw + (grp * SG + ob) * SSEQ + jl;
        so[0] = r; so[HU] = z; so[2 * HU] = n; so[3 * HU] = qv; so[4 * HU] = h;
      }
    };
    // With several groups, group g's post-processing (a chain of dependent shuffles and MUFU ops) sits in the same basic
    // block as group g+1's mat-vec so that the scheduler may interleave them.  Half-size groups (SG = G/2, twice as many
    // of them) were tried to hide the chain at B <= 296 too: 2267 vs 2006 clk/step at H = 128 -- the duplicated gate work
    // and the second barrier wait cost more than the overlap gains -- so the launcher uses SG = G.
    float2 accs[2][NH][SG][3];
#pragma unroll
    for (int grp = 0; grp < NGRP; ++grp) {
      // ---- wait for h_{t-1} of this group (all CTAs' slices) ----
      if (t > 0) {
        mbar_wait_susp(&bars[grp * 2 + ppar], (uint32_t)(((t - 1) >> 1) & 1));
        if (tid == 0 && t + 1 < T) mbar_expect_tx(&bars[grp * 2 + ppar], TXB);   // re-arm for step t+1's values
      }
      float2 (&acc)[NH][SG][3] = accs[grp & 1];
#pragma unroll
      for (int u = 0; u < NH; ++u)
#pragma unroll
        for (int b = 0; b < SG; ++b) {
          const bool mine = (b == ql) && (u == uo);       // the lane that owns (unit, sequence) seeds its hidden biases
          acc[u][b][0] = make_float2(mine ? bh[0] : 0.f, 0.f);
          acc[u][b][1] = make_float2(mine ? bh[1] : 0.f, 0.f);
          acc[u][b][2] = make_float2(0.f, 0.f);
        }
      const uint32_t hc = hbuf_a + (uint32_t)((ppar * BT + grp * SG) * HR) * 4u + 16u * (uint32_t)qk;
      // operand fetches run one k-block ahead of the FFMA2s that consume them (two warps per scheduler cannot hide the
      // shared-memory latency by themselves: short-scoreboard was the top stall of the mat-vec)
      float4 hv[2][SG];
#pragma unroll
      for (int b = 0; b < SG; ++b) hv[0][b] = lds_v4(hc + (uint32_t)(b * HR) * 4u);
#pragma unroll
      for (int i = 0; i < KS / 4; ++i) {
        if (i + 1 < KS / 4) {
#pragma unroll
          for (int b = 0; b < SG; ++b)
            hv[(i + 1) & 1][b] = lds_v4(hc + (uint32_t)(b * HR) * 4u + (uint32_t)((i + 1) * L2) * 16u);
        }
        if (CL_IO_IN_MATVEC && grp == 0 && i == 1) {
          if (t > 0) store(t - 1);
          prefetch(t + CL_PF - 1);
        }
#pragma unroll
        for (int b = 0; b < SG; ++b) {
          const float2 h01 = make_float2(hv[i & 1][b].x, hv[i & 1][b].y), h23 = make_float2(hv[i & 1][b].z, hv[i & 1][b].w);
#pragma unroll
          for (int u = 0; u < NH; ++u)
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              acc[u][b][g] = __ffma2_rn(w[u][g][2 * i], h01, acc[u][b][g]);
              acc[u][b][g] = __ffma2_rn(w[u][g][2 * i + 1], h23, acc[u][b][g]);
            }
        }
      }
      if (grp > 0) post(grp - 1, accs[(grp - 1) & 1]);
    }
    post(NGRP - 1, accs[(NGRP - 1) & 1]);
  }
  __syncthreads();
  store(T - 1);
  cp_async_wait<0>();
  cluster_sync_all();      // nobody leaves while a peer could still be writing into its shared memory
}

// =====================================================================================================================
// BPTT
// =====================================================================================================================
struct ClBwdParams {
  const float* dy;   // (B,T,H), or (B,H) when dy_last
  const float* rzn;  // (B,T,3H)
  const float* q;    // (B,T,H)
  const float* y;    // (B,T,H): h_{t-1} is row t-1
  const float* whh;  // (3H,H)
  float* dgi;        // (B,T,3H) out: dar, daz, dan
  float* dq;         // (B,T,H)  out: dan * r
  int B, T, dy_last;
};

// PF = depth of the input ring (prefetch distance PF-1 steps); 2 where three groups per cluster leave no room for 4
// DIO ("direct I/O"): no input ring, no output staging, no block barrier in the loop -- every lane loads the six saved
// values of ITS (sequence, column) for step t-1 straight from global memory into registers while step t runs (the lanes
// of a warp cover 32-byte sectors completely: 32/G consecutive columns x G sequences), and stores its four outputs
// directly.  The only synchronisation left per step is the mbarrier of the all-gathered dGH vector.  The WAR hazard on
// the double-buffered dGH vectors is closed by data dependence exactly as between CTAs: a warp can send step s+1's
// values only after it has received step s's values from EVERY warp of the cluster, and a warp sends those after its
// step-s mat-vec, i.e. after its last read of the buffer that step s+1 overwrites.
template <int H, int CS, int NGRP, int PF = CL_PF, bool DIO = false>
struct ClBwdSmem {
  static constexpr int HU = H / CS, G = CL_THREADS / HU, BT = G * NGRP, HR = H + CL_HPAD;
  static constexpr int DBUF = 2 * BT * 3 * HR;         // floats: dGH vectors, double-buffered
  static constexpr int SPAD = 32 / G, RSEQ = 6 * HU + SPAD, SSEQ = 4 * HU + SPAD;   // padded per-sequence blocks (see ClFwdSmem)
  static constexpr int RING = DIO ? 0 : PF * BT * RSEQ;          // r,z,n,q,h_{t-1},dy
  static constexpr int STG = DIO ? 0 : 2 * BT * SSEQ;            // dar,daz,dan,dq
  static constexpr size_t bytes = (size_t)(DBUF + RING + STG) * 4 + 128;
};

// NO = output columns per thread (2 or 4): NO*G lanes share a column group, each holding H/(NO*G) rows of W_hh^T per gate
// and column, so that one LDS.128 of the dGH vector feeds 2*NO FFMA2 -- NO = 4 halves the operand fetches of the mat-vec
// again (24 instead of 48 LDS.128 per thread and step at H = 128) for one more shuffle round in the reduction.
template <int H, int CS, int NGRP, int PF = CL_PF, bool DIO = false, int NO = 2>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(CL_THREADS, 1) gru_cl_bwd_kernel(ClBwdParams p) {
  using S = ClBwdSmem<H, CS, NGRP, PF, DIO>;
  constexpr int HU = S::HU, G = S::G, BT = S::BT, HR = S::HR;
  constexpr int L2 = NO * G;           // lanes per output group
  constexpr int J = H / L2;            // j values per lane and gate
  static_assert(NO * 3 * J == 96 && (NO == 2 || NO == 4) && L2 <= 32, "96 weight registers per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);            // [NGRP][2]
  float* dbuf = reinterpret_cast<float*>(smem_raw + 128);            // [2][BT][3][HR]
  float* ring = dbuf + S::DBUF;                                      // [PF][BT][6][HU]
  float* stg = ring + S::RING;                                       // [2][BT][4][HU]
  const int tid = threadIdx.x, kp = tid / L2, ql = tid % L2;
  const uint32_t rank = cluster_ctarank();
  const int T = p.T;
  const int b0 = (blockIdx.x / CS) * BT;
  const int ob = ql % G, oo = ql / G;              // the (sequence-in-group, output-of-the-pair) this lane finishes
  const int kl = NO * kp + oo;                     // its output column inside the CTA's slice
  const int k = (int)rank * HU + kl;

  for (int i = tid; i < S::DBUF + S::RING; i += CL_THREADS) dbuf[i] = 0.f;     // (DIO: RING = 0)
  constexpr uint32_t TXB = (uint32_t)(3 * H * G * 4);
  if (tid == 0) {
    for (int i = 0; i < 2 * NGRP; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    for (int i = 0; i < 2 * NGRP; ++i) mbar_expect_tx(&bars[i], TXB);
  }
  __syncthreads();

  // ---- this thread's share of the step's HBM traffic: fixed items, addresses computed once ----
  // ring slot layout per (slot, b): [r | z | n | q | h_{t-1} | dy] x HU ; staging per (parity, b): [dar|daz|dan|dq] x HU
  constexpr int Q = HU / 4;
  constexpr int NLI = BT * 6 * Q, NSI = BT * 4 * Q;
  constexpr int NLD = (NLI + CL_THREADS - 1) / CL_THREADS, NSD = (NSI + CL_THREADS - 1) / CL_THREADS;
  constexpr int RSEQ = S::RSEQ, SSEQ = S::SSEQ;
  constexpr uint32_t SLOT_BYTES = (uint32_t)(BT * RSEQ) * 4u, STG_BYTES = (uint32_t)(BT * SSEQ) * 4u;
  const float* lp[NLD]; uint32_t ls[NLD]; bool lv[NLD]; int lst[NLD]; bool lshift[NLD];
#pragma unroll
  for (int m = 0; m < NLD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, wch = (n / Q) % 6, b = n / (6 * Q);
    lv[m] = (n < NLI) && (b0 + b < p.B) && !(wch == 5 && p.dy_last);
    const size_t seq = (size_t)(b0 + (lv[m] ? b : 0)) * T;
    const int col = (int)rank * HU + j4 * 4;
    lshift[m] = wch == 4;                                  // h_{t-1}: row t-1 of y, nothing at t = 0
    const float* base = (wch < 3) ? p.rzn + seq * (3 * H) + wch * H
                        : (wch == 3 ? p.q + seq * H : (wch == 4 ? p.y + seq * H - H : p.q + seq * H));
    if (wch == 5 && !p.dy_last) base = p.dy + seq * H;
    lp[m] = base + col;
    lst[m] = (wch < 3) ? 3 * H : H;
    ls[m] = smem_u32(ring) + (uint32_t)(b * RSEQ + wch * HU + j4 * 4) * 4u;
  }
  float* sp[NSD]; uint32_t ss[NSD]; bool sv[NSD]; int sst[NSD];
#pragma unroll
  for (int m = 0; m < NSD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, wch = (n / Q) % 4, b = n / (4 * Q);
    sv[m] = (n < NSI) && (b0 + b < p.B);
    const size_t seq = (size_t)(b0 + (sv[m] ? b : 0)) * T;
    sp[m] = ((wch < 3) ? p.dgi + seq * (3 * H) + wch * H : p.dq + seq * H) + (int)rank * HU + j4 * 4;
    sst[m] = (wch < 3) ? 3 * H : H;
    ss[m] = smem_u32(stg) + (uint32_t)(b * SSEQ + wch * HU + j4 * 4) * 4u;
  }
  auto prefetch = [&](int t) {         // t counts down; t < 0: nothing to load
    if (t >= 0) {
      const uint32_t so = (uint32_t)(t % PF) * SLOT_BYTES;
#pragma unroll
      for (int m = 0; m < NLD; ++m)
        if (lv[m] && !(lshift[m] && t == 0)) cp_async16(ls[m] + so, lp[m] + (size_t)t * lst[m]);
    }
    cp_async_commit();
  };
  auto store = [&](int t, int s) {
    const uint32_t so = (uint32_t)(s & 1) * STG_BYTES;
#pragma unroll
    for (int m = 0; m < NSD; ++m)
      if (sv[m]) *reinterpret_cast<float4*>(sp[m] + (size_t)t * sst[m]) = lds_v4(ss[m] + so);
  };
  if (!DIO)
    for (int s = 0; s < PF - 1; ++s) prefetch(T - 1 - s);

  // ---- DIO: this lane's saved activations of step t (cur) and t-1 (nxt), per sequence group ----
  struct Saved { float r, z, n, q, hp, dy; };
  Saved cur[NGRP], nxt[NGRP];
  bool seq_ok[NGRP];
  size_t row0[NGRP];                      // (b0 + grp*G + ob) * T
#pragma unroll
  for (int g = 0; g < NGRP; ++g) {
    seq_ok[g] = (b0 + g * G + ob) < p.B;
    row0[g] = (size_t)(b0 + (seq_ok[g] ? g * G + ob : 0)) * T;
  }
  auto load_saved = [&](int g, int t) {
    Saved v = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (seq_ok[g] && t >= 0) {
      const size_t row = row0[g] + (size_t)t;
      const float* a = p.rzn + row * (3 * H) + k;
      v.r = __ldg(a); v.z = __ldg(a + H); v.n = __ldg(a + 2 * H);
      v.q = __ldg(p.q + row * H + k);
      if (t > 0) v.hp = __ldg(p.y + (row - 1) * H + k);
      if (!p.dy_last) v.dy = __ldg(p.dy + row * H + k);
    }
    return v;
  };
  if (DIO) {
#pragma unroll
    for (int g = 0; g < NGRP; ++g) cur[g] = load_saved(g, T - 1);
  }

  // =============================== compute warps ===============================
  // ---- W_hh^T slices: wt[o][g][m] = (W[gH+jj][k0+o], W[gH+jj+1][k0+o]),  jj = (i*L2+ql)*4 + 2*(m&1), i = m>>1 ----
  float2 wt[NO][3][J / 2];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < J / 4; ++i) {
        const int jj = (i * L2 + ql) * 4;
        const int kk = (int)rank * HU + NO * kp + o;
        const float a0 = p.whh[(size_t)(g * H + jj + 0) * H + kk], a1 = p.whh[(size_t)(g * H + jj + 1) * H + kk];
        const float a2 = p.whh[(size_t)(g * H + jj + 2) * H + kk], a3 = p.whh[(size_t)(g * H + jj + 3) * H + kk];
        wt[o][g][2 * i] = make_float2(a0, a1);
        wt[o][g][2 * i + 1] = make_float2(a2, a3);
      }
  cluster_sync_all();

  // Senders: output columns come in runs of four consecutive kl = 2*kp + oo; the run of lane (kp, oo, ob) lives in the
  // lanes ((kr + i) >> 1) * L2 + ((kr + i) & 1) * G + ob of the same warp (kr = first column of the run in the warp)
  const int lane = tid & 31;
  const int kr = (kl & ~3) - NO * (kp - kp % (32 / L2));           // first column of the run, relative to the warp's
  int srcl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) srcl[i] = ((kr + i) / NO) * L2 + ((kr + i) % NO) * G + ob;
  (void)lane;
  // lane i of the run (kl & 3) serves the destination CTAs c == i (mod 4): one warp instruction covers four CTAs
  constexpr int ND = (CS + 3) / 4;
  uint32_t r_d[ND], r_bar[ND];
  bool r_ok[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const int c = (kl & 3) + 4 * d;
    r_ok[d] = c < CS;
    r_d[d] = mapa_shared(smem_u32(dbuf + (ob * 3) * HR + (k & ~3)), (uint32_t)(r_ok[d] ? c : 0));
    r_bar[d] = mapa_shared(smem_u32(bars), (uint32_t)(r_ok[d] ? c : 0));
  }
  float carry[NGRP];
#pragma unroll
  for (int g = 0; g < NGRP; ++g)
    carry[g] = (p.dy_last && (b0 + g * G + ob) < p.B) ? p.dy[(size_t)(b0 + g * G + ob) * H + k] : 0.f;
  const uint32_t dbuf_a = smem_u32(dbuf);

  for (int s = 0; s < T; ++s) {
    const int t = T - 1 - s;
    if (!DIO) {
      cp_async_wait<PF - 2>();
      __syncthreads();
      if (!CL_IO_IN_MATVEC || s == 0) {    // (the first step has no mat-vec to hide the I/O behind)
        if (s > 0) store(t + 1, s - 1);    // outputs of step t+1
        prefetch(t - (PF - 1));
      }
    } else {
#pragma unroll
      for (int g = 0; g < NGRP; ++g) nxt[g] = load_saved(g, t - 1);     // in flight for the whole step
    }
    const int par = s & 1, ppar = par ^ 1;
    float* sgw = stg + par * (BT * SSEQ);
#pragma unroll
    for (int grp = 0; grp < NGRP; ++grp) {
      // saved activations of (b, t, k): everything that does not depend on the carried dh first
      float r, z, n, qv, dyv, hp;
      if (DIO) {
        r = cur[grp].r; z = cur[grp].z; n = cur[grp].n; qv = cur[grp].q; dyv = cur[grp].dy; hp = cur[grp].hp;
      } else {
        const float* rg = ring + ((t % PF) * BT + grp * G + ob) * RSEQ + kl;
        r = rg[0]; z = rg[HU]; n = rg[2 * HU]; qv = rg[3 * HU]; dyv = rg[5 * HU];
        hp = (t == 0) ? 0.f : rg[4 * HU];          // h_{-1} = 0 (row -1 is never loaded; the slot is stale)
      }
      const float omz = 1.f - z;
      const float fA = omz * fmaf(-n, n, 1.f);
      const float fB = (hp - n) * (z * omz);
      const float fC = qv * (r * (1.f - r));
      // ---- W_hh^T dGH_{t+1}: wait for the all-gathered vector of the previous step ----
      if (s > 0) {
        mbar_wait_susp(&bars[grp * 2 + ppar], (uint32_t)(((s - 1) >> 1) & 1));
        if (tid == 0 && s + 1 < T) mbar_expect_tx(&bars[grp * 2 + ppar], TXB);
        float2 acc[G][NO];
#pragma unroll
        for (int b = 0; b < G; ++b)
#pragma unroll
          for (int o = 0; o < NO; ++o) acc[b][o] = make_float2(0.f, 0.f);
        const uint32_t dc = dbuf_a + (uint32_t)((ppar * BT + grp * G) * 3 * HR) * 4u + 16u * (uint32_t)ql;
        // operand fetches run one (row block, gate) ahead of the FFMA2s that consume them
        float4 dvv[2][G];
#pragma unroll
        for (int b = 0; b < G; ++b) dvv[0][b] = lds_v4(dc + (uint32_t)((b * 3) * HR) * 4u);
#pragma unroll
        for (int it = 0; it < 3 * (J / 4); ++it) {
          const int i = it / 3, g = it % 3;
          if (it + 1 < 3 * (J / 4)) {
            const int i1 = (it + 1) / 3, g1 = (it + 1) % 3;
#pragma unroll
            for (int b = 0; b < G; ++b)
              dvv[(it + 1) & 1][b] = lds_v4(dc + (uint32_t)((b * 3 + g1) * HR) * 4u + (uint32_t)(i1 * L2) * 16u);
          }
          if (CL_IO_IN_MATVEC && grp == 0 && it == 1) {      // this step's HBM traffic rides in the mat-vec's free issue slots
            store(t + 1, s - 1);
            prefetch(t - (PF - 1));
          }
#pragma unroll
          for (int b = 0; b < G; ++b) {
            const float4 dv = dvv[it & 1][b];
            const float2 d01 = make_float2(dv.x, dv.y), d23 = make_float2(dv.z, dv.w);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
              acc[b][o] = __ffma2_rn(wt[o][g][2 * i], d01, acc[b][o]);
              acc[b][o] = __ffma2_rn(wt[o][g][2 * i + 1], d23, acc[b][o]);
            }
          }
        }
        float v[L2];
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
          for (int b = 0; b < G; ++b) v[o * G + b] = acc[b][o].x + acc[b][o].y;
        reduce_scatter<L2>(v, ql);
        carry[grp] += v[0];
      }
      const float dh = dyv + carry[grp];
      const float dan = dh * fA;
      const float daz = dh * fB;
      const float dar = dan * fC;
      const float dqv = dan * r;
      carry[grp] = dh * z;            // the direct path dL/dh_{t-1} += dh * z; the mat-vec part is added next step
      if (s + 1 < T) {
        const uint32_t off = (uint32_t)((par * BT + grp * G) * 3 * HR) * 4u;
        float vr[4], vz[4], vq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vr[i] = __shfl_sync(0xffffffffu, dar, srcl[i]);
          vz[i] = __shfl_sync(0xffffffffu, daz, srcl[i]);
          vq[i] = __shfl_sync(0xffffffffu, dqv, srcl[i]);
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
          if (r_ok[d]) {
            const uint32_t bar = r_bar[d] + (uint32_t)(grp * 2 + par) * 8u;
            st_async_v4(r_d[d] + off, vr[0], vr[1], vr[2], vr[3], bar);
            st_async_v4(r_d[d] + off + (uint32_t)HR * 4u, vz[0], vz[1], vz[2], vz[3], bar);
            st_async_v4(r_d[d] + off + 2u * (uint32_t)HR * 4u, vq[0], vq[1], vq[2], vq[3], bar);
          }
      }
      if (DIO) {
        if (seq_ok[grp]) {
          const size_t row = row0[grp] + (size_t)t;
          float* o = p.dgi + row * (3 * H) + k;
          o[0] = dar; o[H] = daz; o[2 * H] = dan;
          p.dq[row * H + k] = dqv;
        }
      } else {
        float* so = sgw + (grp * G + ob) * SSEQ + kl;
        so[0] = dar; so[HU] = daz; so[2 * HU] = dan; so[3 * HU] = dqv;
      }
    }
    if (DIO) {
#pragma unroll
      for (int g = 0; g < NGRP; ++g) cur[g] = nxt[g];
    }
  }
  if (!DIO) {
    __syncthreads();
    store(0, T - 1);
    cp_async_wait<0>();
  }
  cluster_sync_all();
}

// =====================================================================================================================
// reverse over (primal + tangent) forward -- the R1 double backward (train_timegan.py:198-202)
// =====================================================================================================================
// Same skeleton as the BPTT kernel, with TWO carried adjoints (hb of y, hdb of ydot) that share the register-resident
// W_hh^T slices: six vectors are all-gathered per step (primal dGH: arb, azb, qb; tangent dGH: arb_d, azb_d, qdb) and
// two mat-vecs are accumulated from them.  Every output of a step is linear in (hb, hdb) with coefficients that only
// depend on saved activations (derivation: oracle/gru_math.py gru_layer_jvp_bwd; the one-SM kernel gru_jvp_bwd_kernel of
// gru_jvp.cu uses the same coefficient pairs), so they are formed before the mbarrier wait.
struct ClJbParams {
  const float* hbar;   // (B,T,H) or (B,H) if last_only
  const float* hdbar;
  const float* rzn;    // (B,T,3H)
  const float* q;      // (B,T,H)
  const float* ta;     // (B,T,3H) tangent pre-activations
  const float* qdot;   // (B,T,H)
  const float* y;      // (B,T,H)
  const float* ydot;   // (B,T,H)
  const float* whh;
  float* gib;          // (B,T,3H)
  float* qb;           // (B,T,H)
  float* gidb;         // (B,T,3H)
  float* qdb;          // (B,T,H)
  int B, T, last_only;
};

template <int H, int CS, int NGRP>
struct ClJbSmem {
  static constexpr int HU = H / CS, G = CL_THREADS / HU, BT = G * NGRP, HR = H + CL_HPAD;
  static constexpr int DBUF = 2 * BT * 6 * HR;
  static constexpr int SPAD = 32 / G, RSEQ = 12 * HU + SPAD, SSEQ = 8 * HU + SPAD;  // padded per-sequence blocks (see ClFwdSmem)
  static constexpr int RING = CL_PF * BT * RSEQ;       // r,z,n,q, ar,az,an,qd, h_{t-1}, hd_{t-1}, hbar, hdbar
  static constexpr int STG = 2 * BT * SSEQ;            // arb,azb,anb,qb, arb_d,azb_d,anb_d,qdb
  static constexpr size_t bytes = (size_t)(DBUF + RING + STG) * 4 + 128;
};

template <int H, int CS, int NGRP, int NO = 2>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(CL_THREADS, 1) gru_cl_jvp_bwd_kernel(ClJbParams p) {
  using S = ClJbSmem<H, CS, NGRP>;
  constexpr int HU = S::HU, G = S::G, BT = S::BT, HR = S::HR;
  constexpr int L2 = NO * G, J = H / L2;       // NO output columns per thread (see gru_cl_bwd_kernel)
  static_assert(NO * 3 * J == 96 && (NO == 2 || NO == 4) && L2 <= 32, "96 weight registers per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  float* dbuf = reinterpret_cast<float*>(smem_raw + 128);            // [2][BT][6][HR]
  float* ring = dbuf + S::DBUF;                                      // [PF][BT][12][HU]
  float* stg = ring + S::RING;                                       // [2][BT][8][HU]
  const int tid = threadIdx.x, kp = tid / L2, ql = tid % L2;
  const uint32_t rank = cluster_ctarank();
  const int T = p.T;
  const int b0 = (blockIdx.x / CS) * BT;
  const int ob = ql % G, oo = ql / G;
  const int kl = NO * kp + oo;
  const int k = (int)rank * HU + kl;

  for (int i = tid; i < S::DBUF + S::RING; i += CL_THREADS) dbuf[i] = 0.f;
  constexpr uint32_t TXB = (uint32_t)(6 * H * G * 4);
  if (tid == 0) {
    for (int i = 0; i < 2 * NGRP; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    for (int i = 0; i < 2 * NGRP; ++i) mbar_expect_tx(&bars[i], TXB);
  }
  __syncthreads();

  // ---- this thread's share of the step's HBM traffic ----
  constexpr int Q = HU / 4;
  constexpr int NLI = BT * 12 * Q, NSI = BT * 8 * Q;
  constexpr int NLD = (NLI + CL_THREADS - 1) / CL_THREADS, NSD = (NSI + CL_THREADS - 1) / CL_THREADS;
  constexpr int RSEQ = S::RSEQ, SSEQ = S::SSEQ;
  constexpr uint32_t SLOT_BYTES = (uint32_t)(BT * RSEQ) * 4u, STG_BYTES = (uint32_t)(BT * SSEQ) * 4u;
  const float* lp[NLD]; uint32_t ls[NLD]; bool lv[NLD]; int lst[NLD]; bool lshift[NLD];
#pragma unroll
  for (int m = 0; m < NLD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, wch = (n / Q) % 12, b = n / (12 * Q);
    lv[m] = (n < NLI) && (b0 + b < p.B) && !(wch >= 10 && p.last_only);
    const size_t seq = (size_t)(b0 + (lv[m] ? b : 0)) * T;
    const int col = (int)rank * HU + j4 * 4;
    lshift[m] = (wch == 8 || wch == 9);                    // h_{t-1} / hd_{t-1}: row t-1, nothing at t = 0
    const float* base = p.q + seq * H;
    if (wch < 3) base = p.rzn + seq * (3 * H) + wch * H;
    else if (wch == 3) base = p.q + seq * H;
    else if (wch < 7) base = p.ta + seq * (3 * H) + (wch - 4) * H;
    else if (wch == 7) base = p.qdot + seq * H;
    else if (wch == 8) base = p.y + seq * H - H;
    else if (wch == 9) base = p.ydot + seq * H - H;
    else if (!p.last_only) base = (wch == 10 ? p.hbar : p.hdbar) + seq * H;
    lp[m] = base + col;
    lst[m] = (wch < 3 || (wch >= 4 && wch < 7)) ? 3 * H : H;
    ls[m] = smem_u32(ring) + (uint32_t)(b * RSEQ + wch * HU + j4 * 4) * 4u;
  }
  float* sp[NSD]; uint32_t ss[NSD]; bool sv[NSD]; int sst[NSD];
#pragma unroll
  for (int m = 0; m < NSD; ++m) {
    const int n = tid + CL_THREADS * m, j4 = n % Q, wch = (n / Q) % 8, b = n / (8 * Q);
    sv[m] = (n < NSI) && (b0 + b < p.B);
    const size_t seq = (size_t)(b0 + (sv[m] ? b : 0)) * T;
    float* base = (wch < 3) ? p.gib + seq * (3 * H) + wch * H
                  : (wch == 3 ? p.qb + seq * H : (wch < 7 ? p.gidb + seq * (3 * H) + (wch - 4) * H : p.qdb + seq * H));
    sp[m] = base + (int)rank * HU + j4 * 4;
    sst[m] = (wch < 3 || (wch >= 4 && wch < 7)) ? 3 * H : H;
    ss[m] = smem_u32(stg) + (uint32_t)(b * SSEQ + wch * HU + j4 * 4) * 4u;
  }
  auto prefetch = [&](int t) {
    if (t >= 0) {
      const uint32_t so = (uint32_t)(t % CL_PF) * SLOT_BYTES;
#pragma unroll
      for (int m = 0; m < NLD; ++m)
        if (lv[m] && !(lshift[m] && t == 0)) cp_async16(ls[m] + so, lp[m] + (size_t)t * lst[m]);
    }
    cp_async_commit();
  };
  auto store = [&](int t, int s) {
    const uint32_t so = (uint32_t)(s & 1) * STG_BYTES;
#pragma unroll
    for (int m = 0; m < NSD; ++m)
      if (sv[m]) *reinterpret_cast<float4*>(sp[m] + (size_t)t * sst[m]) = lds_v4(ss[m] + so);
  };
  for (int s = 0; s < CL_PF - 1; ++s) prefetch(T - 1 - s);

  float2 wt[NO][3][J / 2];
#pragma unroll
  for (int o = 0; o < NO; ++o)
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < J / 4; ++i) {
        const int jj = (i * L2 + ql) * 4;
        const int kk = (int)rank * HU + NO * kp + o;
        const float a0 = p.whh[(size_t)(g * H + jj + 0) * H + kk], a1 = p.whh[(size_t)(g * H + jj + 1) * H + kk];
        const float a2 = p.whh[(size_t)(g * H + jj + 2) * H + kk], a3 = p.whh[(size_t)(g * H + jj + 3) * H + kk];
        wt[o][g][2 * i] = make_float2(a0, a1);
        wt[o][g][2 * i + 1] = make_float2(a2, a3);
      }
  cluster_sync_all();

  const int kr = (kl & ~3) - NO * (kp - kp % (32 / L2));
  int srcl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) srcl[i] = ((kr + i) / NO) * L2 + ((kr + i) % NO) * G + ob;
  constexpr int ND = (CS + 3) / 4;
  uint32_t r_d[ND], r_bar[ND];
  bool r_ok[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const int c = (kl & 3) + 4 * d;
    r_ok[d] = c < CS;
    r_d[d] = mapa_shared(smem_u32(dbuf + (ob * 6) * HR + (k & ~3)), (uint32_t)(r_ok[d] ? c : 0));
    r_bar[d] = mapa_shared(smem_u32(bars), (uint32_t)(r_ok[d] ? c : 0));
  }
  float ch[NGRP], chd[NGRP];
#pragma unroll
  for (int g = 0; g < NGRP; ++g) {
    const bool on = p.last_only && (b0 + g * G + ob) < p.B;
    ch[g] = on ? p.hbar[(size_t)(b0 + g * G + ob) * H + k] : 0.f;
    chd[g] = on ? p.hdbar[(size_t)(b0 + g * G + ob) * H + k] : 0.f;
  }
  const uint32_t dbuf_a = smem_u32(dbuf);

  for (int s = 0; s < T; ++s) {
    const int t = T - 1 - s;
    cp_async_wait<CL_PF - 2>();
    __syncthreads();
    if (s > 0) store(t + 1, s - 1);
    prefetch(t - (CL_PF - 1));
    const int par = s & 1, ppar = par ^ 1;
    float* sgw = stg + par * (BT * SSEQ);
#pragma unroll
    for (int grp = 0; grp < NGRP; ++grp) {
      // ---- coefficient pairs of (b, t, k): out = alpha hb + beta hdb ----
      const float* rg = ring + ((t % CL_PF) * BT + grp * G + ob) * RSEQ + kl;
      const float rt = rg[0], zt = rg[HU], nt = rg[2 * HU], qt = rg[3 * HU];
      const float art = rg[4 * HU], azt = rg[5 * HU], ant = rg[6 * HU], qdt = rg[7 * HU];
      const float hp = (t == 0) ? 0.f : rg[8 * HU], hdp = (t == 0) ? 0.f : rg[9 * HU];
      const float fhb = p.last_only ? 0.f : rg[10 * HU], fhdb = p.last_only ? 0.f : rg[11 * HU];
      const float sr = rt * (1.f - rt), omz = 1.f - zt, sz = zt * omz, sn = fmaf(-nt, nt, 1.f);
      const float rdot = sr * art, zdot = sz * azt, ndot = sn * ant, hmn = hp - nt;
      const float K1 = sn * omz;
      const float K2 = sn * (-zdot - 2.f * nt * ant * omz);
      const float rdbB = qt * K1;
      const float ardB = sr * rdbB;
      const float qA = rt * K1, qB = fmaf(rdot, K1, rt * K2);
      const float arA = sr * (qt * K1);
      const float arB = sr * fmaf(qdt, K1, fmaf(qt, K2, (1.f - 2.f * rt) * art * rdbB));
      const float azA = sz * hmn;
      const float azB = sz * ((hdp - ndot) + (1.f - 2.f * zt) * azt * hmn);
      // ---- W_hh^T (primal dGH, tangent dGH) of the previous step ----
      if (s > 0) {
        mbar_wait_susp(&bars[grp * 2 + ppar], (uint32_t)(((s - 1) >> 1) & 1));
        if (tid == 0 && s + 1 < T) mbar_expect_tx(&bars[grp * 2 + ppar], TXB);
        float2 acc[G][2][NO];         // [sequence][adjoint][output column]
#pragma unroll
        for (int b = 0; b < G; ++b)
#pragma unroll
          for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int o = 0; o < NO; ++o) acc[b][v][o] = make_float2(0.f, 0.f);
        const uint32_t dc = dbuf_a + (uint32_t)((ppar * BT + grp * G) * 6 * HR) * 4u + 16u * (uint32_t)ql;
#pragma unroll
        for (int i = 0; i < J / 4; ++i)
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int b = 0; b < G; ++b)
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const float4 dv = lds_v4(dc + (uint32_t)((b * 6 + v * 3 + g) * HR) * 4u + (uint32_t)(i * L2) * 16u);
                const float2 d01 = make_float2(dv.x, dv.y), d23 = make_float2(dv.z, dv.w);
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                  acc[b][v][o] = __ffma2_rn(wt[o][g][2 * i], d01, acc[b][v][o]);
                  acc[b][v][o] = __ffma2_rn(wt[o][g][2 * i + 1], d23, acc[b][v][o]);
                }
              }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          float u[L2];
#pragma unroll
          for (int o = 0; o < NO; ++o)
#pragma unroll
            for (int b = 0; b < G; ++b) u[o * G + b] = acc[b][v][o].x + acc[b][v][o].y;
          reduce_scatter<L2>(u, ql);
          if (v == 0) ch[grp] += u[0]; else chd[grp] += u[0];
        }
      }
      const float hb = fhb + ch[grp];
      const float hdb = fhdb + chd[grp];
      const float arb = fmaf(arA, hb, arB * hdb);
      const float azb = fmaf(azA, hb, azB * hdb);
      const float anb = fmaf(K1, hb, K2 * hdb);
      const float qb = fmaf(qA, hb, qB * hdb);
      const float arb_d = ardB * hdb, azb_d = azA * hdb, anb_d = K1 * hdb, qdb = qA * hdb;
      ch[grp] = fmaf(zt, hb, zdot * hdb);      // direct paths into h_{t-1} / hd_{t-1}; the mat-vec parts follow next step
      chd[grp] = zt * hdb;
      if (s + 1 < T) {
        const uint32_t off = (uint32_t)((par * BT + grp * G) * 6 * HR) * 4u;
        const float vals[6] = {arb, azb, qb, arb_d, azb_d, qdb};
#pragma unroll
        for (int e = 0; e < 6; ++e) {
          const float v0 = __shfl_sync(0xffffffffu, vals[e], srcl[0]), v1 = __shfl_sync(0xffffffffu, vals[e], srcl[1]),
                      v2 = __shfl_sync(0xffffffffu, vals[e], srcl[2]), v3 = __shfl_sync(0xffffffffu, vals[e], srcl[3]);
#pragma unroll
          for (int d = 0; d < ND; ++d)
            if (r_ok[d])
              st_async_v4(r_d[d] + off + (uint32_t)(e * HR) * 4u, v0, v1, v2, v3, r_bar[d] + (uint32_t)(grp * 2 + par) * 8u);
        }
      }
      float* so = sgw + (grp * G + ob) * SSEQ + kl;
      so[0] = arb; so[HU] = azb; so[2 * HU] = anb; so[3 * HU] = qb;
      so[4 * HU] = arb_d; so[5 * HU] = azb_d; so[6 * HU] = anb_d; so[7 * HU] = qdb;
    }
  }
  __syncthreads();
  store(0, T - 1);
  cp_async_wait<0>();
  cluster_sync_all();
}

template <int H, int CS, int NGRP, int NO = 2>
int launch_cl_jb(cudaStream_t st, const ClJbParams& p) {
  using S = ClJbSmem<H, CS, NGRP>;
  auto kern = gru_cl_jvp_bwd_kernel<H, CS, NGRP, NO>;
  TG_OPT_IN_SMEM(kern, "gru_cl_jvp_bwd");
  const int clusters = (p.B + S::BT - 1) / S::BT;
  kern<<<clusters * CS, CL_THREADS, S::bytes, st>>>(p);
  return tg_check_launch("gru_cl_jvp_bwd");
}

template <int H, int CS, int NGRP, int SG, int NH = 1>
int launch_cl_fwd(cudaStream_t st, const ClFwdParams& p) {
  using S = ClFwdSmem<H, CS, NGRP, SG>;
  auto kern = gru_cl_fwd_kernel<H, CS, NGRP, SG, NH>;
  TG_OPT_IN_SMEM(kern, "gru_cl_fwd");
  const int clusters = (p.B + S::BT - 1) / S::BT;
  kern<<<clusters * CS, CL_THREADS, S::bytes, st>>>(p);
  return tg_check_launch("gru_cl_fwd");
}

template <int H, int CS, int NGRP, int PF = CL_PF, bool DIO = false, int NO = 2>
int launch_cl_bwd(cudaStream_t st, const ClBwdParams& p) {
  using S = ClBwdSmem<H, CS, NGRP, PF, DIO>;
  auto kern = gru_cl_bwd_kernel<H, CS, NGRP, PF, DIO, NO>;
  TG_OPT_IN_SMEM(kern, "gru_cl_bwd");
  const int clusters = (p.B + S::BT - 1) / S::BT;
  kern<<<clusters * CS, CL_THREADS, S::bytes, st>>>(p);
  return tg_check_launch("gru_cl_bwd");
}

// How many clusters of a kernel can be resident at once (cudaOccupancyMaxActiveClusters): clusters are placed inside
// one GPC, so an 8-CTA cluster does not simply get 148 / 8 slots.  A launch with more clusters than that runs in waves,
// which for a persistent 768-step kernel doubles its time -- the group count per cluster is chosen so that ONE wave
// covers the batch whenever the shared-memory budget allows.
template <typename K>
int max_active_clusters(K kern, int cs, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cs * 64));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return n;
}

struct ClCaps { int f128[2], b128[2], j128[2], f256[4], b256[3], j256[1]; bool ready; };
ClCaps& cl_caps() {
  static ClCaps c = {};
  if (!c.ready) {
    // opt in to the shared memory first (the occupancy query honours the function attribute)
    auto optin = [](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_max_optin_smem()); };
    optin(gru_cl_fwd_kernel<128, 2, 1, 4>); optin(gru_cl_fwd_kernel<128, 2, 2, 4>);
    optin(gru_cl_bwd_kernel<128, 2, 1>); optin(gru_cl_bwd_kernel<128, 2, 2>);
    optin(gru_cl_jvp_bwd_kernel<128, 2, 1>); optin(gru_cl_jvp_bwd_kernel<128, 2, 2>);
    optin(gru_cl_fwd_kernel<256, 8, 1, 8>); optin(gru_cl_fwd_kernel<256, 8, 2, 8>);
    optin(gru_cl_fwd_kernel<256, 8, 3, 8>); optin(gru_cl_fwd_kernel<256, 8, 4, 8>);
    optin(gru_cl_bwd_kernel<256, 8, 1>); optin(gru_cl_bwd_kernel<256, 8, 2>); optin(gru_cl_bwd_kernel<256, 8, 3, 2>);
    optin(gru_cl_jvp_bwd_kernel<256, 8, 1>);
    c.f128[0] = max_active_clusters(gru_cl_fwd_kernel<128, 2, 1, 4>, 2, ClFwdSmem<128, 2, 1, 4>::bytes);
    c.f128[1] = max_active_clusters(gru_cl_fwd_kernel<128, 2, 2, 4>, 2, ClFwdSmem<128, 2, 2, 4>::bytes);
    c.b128[0] = max_active_clusters(gru_cl_bwd_kernel<128, 2, 1>, 2, ClBwdSmem<128, 2, 1>::bytes);
    c.b128[1] = max_active_clusters(gru_cl_bwd_kernel<128, 2, 2>, 2, ClBwdSmem<128, 2, 2>::bytes);
    c.j128[0] = max_active_clusters(gru_cl_jvp_bwd_kernel<128, 2, 1>, 2, ClJbSmem<128, 2, 1>::bytes);
    c.j128[1] = max_active_clusters(gru_cl_jvp_bwd_kernel<128, 2, 2>, 2, ClJbSmem<128, 2, 2>::bytes);
    c.f256[0] = max_active_clusters(gru_cl_fwd_kernel<256, 8, 1, 8>, 8, ClFwdSmem<256, 8, 1, 8>::bytes);
    c.f256[1] = max_active_clusters(gru_cl_fwd_kernel<256, 8, 2, 8>, 8, ClFwdSmem<256, 8, 2, 8>::bytes);
    c.f256[2] = max_active_clusters(gru_cl_fwd_kernel<256, 8, 3, 8>, 8, ClFwdSmem<256, 8, 3, 8>::bytes);
    c.f256[3] = max_active_clusters(gru_cl_fwd_kernel<256, 8, 4, 8>, 8, ClFwdSmem<256, 8, 4, 8>::bytes);
    c.b256[0] = max_active_clusters(gru_cl_bwd_kernel<256, 8, 1>, 8, ClBwdSmem<256, 8, 1>::bytes);
    c.b256[1] = max_active_clusters(gru_cl_bwd_kernel<256, 8, 2>, 8, ClBwdSmem<256, 8, 2>::bytes);
    c.b256[2] = max_active_clusters(gru_cl_bwd_kernel<256, 8, 3, 2>, 8, ClBwdSmem<256, 8, 3, 2>::bytes);
    c.j256[0] = max_active_clusters(gru_cl_jvp_bwd_kernel<256, 8, 1>, 8, ClJbSmem<256, 8, 1>::bytes);
    c.ready = true;
  }
  return c;
}

// smallest group count (1-based index into caps) whose cluster count fits one wave; n if none does
int pick_groups(int B, int seqs_per_group, const int* caps, int n) {
  for (int g = 1; g <= n; ++g) {
    const int clusters = (B + seqs_per_group * g - 1) / (seqs_per_group * g);
    if (caps[g - 1] > 0 && clusters <= caps[g - 1]) return g;
  }
  return n;
}

}  // namespace

// diagnostic: resident-cluster capacity of the (H, direction, groups) instantiation; 0 = unknown / not instantiated
extern "C" int tg_cluster_capacity(int H, int backward, int groups) {
  ClCaps& c = cl_caps();
  if (H == 128 && groups >= 1 && groups <= 2) return backward == 2 ? c.j128[groups - 1] : (backward ? c.b128[groups - 1] : c.f128[groups - 1]);
  if (H == 256 && !backward && groups >= 1 && groups <= 4) return c.f256[groups - 1];
  if (H == 256 && backward == 1 && groups >= 1 && groups <= 3) return c.b256[groups - 1];
  if (H == 256 && backward == 2 && groups == 1) return c.j256[0];
  return 0;
}

// Which (hidden size, batch, direction) the cluster kernels take -- chosen from measurements on a B200 (tools/probe_cluster.py,
// profiles/r02_probe_cluster.log; T = 768, us per layer pass, cluster vs. the kernel it replaces):
//   H = 128  forward   B <= 296: 766 vs 747 (gru_fwd.cu, 512 threads, BT = 2)  -> legacy;   B = 512: 1421 vs 1561 -> cluster
//   H = 128  BPTT      B = 256: 1057 vs 1477,  B = 512: 2076 vs 2956                          -> cluster
//   H = 256  forward   B = 256: 4131 (three groups per cluster, 11 clusters = one wave; 5513 with two groups = 16 clusters in
//                      two waves: only 15 eight-CTA clusters are resident at once) vs 9429 (gru_bigh.cu, W_hh from L2)  -> cluster
//   H = 256  BPTT      cluster when the batch fits one wave with <= 3 groups per cluster (shared-memory limit; the third
//                      group costs the input ring two of its four stages): B = 128: 4603 vs 7703, B = 256: see
//                      profiles/r02_probe_cluster.log; two groups = 16 clusters = two waves took 9184 vs 7693 (gru_bigh.cu)
// Exact sizes only: the k-slices are compile-time register arrays.  TIMEGAN_B200_CLUSTER=0 disables them, =2 forces them
// for every H = 128 / 256 launch (tests).
bool tg_cluster_takes(int H, int B, bool backward) {
  const int mode = tg_use_cluster();
  if (mode == 0 || (H != 128 && H != 256)) return false;
  if (mode == 2) return true;
  const int sms = tg_num_sms();
  if (H == 128) return backward || ((B + 3) / 4) * 2 > sms;
  ClCaps& c = cl_caps();
  if (!backward) return true;      // even in two waves it beats the L2-streaming kernel (5.5 vs 9.4 ms at B = 256)
  const int g = pick_groups(B, 8, c.b256, 3);
  return (B + 8 * g - 1) / (8 * g) <= c.b256[g - 1];
}

int tg_gru_cl_fwd(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T, int H,
                  int save) {
  ClFwdParams p{gi, whh, bhh, y, q, B, T, save};
  ClCaps& c = cl_caps();
  // two hidden units per thread (NH = 2) halves the forward's operand fetches too, but the forward was never bound by them
  // (it reads one h vector for three gates): 745 vs 759 us at B = 256, 1380 vs 1367 us at B = 512 -- NH = 1 stays
  if (H == 128 && tg_cluster_no() == 8)
    return pick_groups(B, 4, c.f128, 2) == 1 ? launch_cl_fwd<128, 2, 1, 4, 2>(st, p) : launch_cl_fwd<128, 2, 2, 4, 2>(st, p);
  if (H == 128) return pick_groups(B, 4, c.f128, 2) == 1 ? launch_cl_fwd<128, 2, 1, 4>(st, p) : launch_cl_fwd<128, 2, 2, 4>(st, p);
  if (H == 256) {
    switch (pick_groups(B, 8, c.f256, 4)) {
      case 1: return launch_cl_fwd<256, 8, 1, 8>(st, p);
      case 2: return launch_cl_fwd<256, 8, 2, 8>(st, p);
      case 3: return launch_cl_fwd<256, 8, 3, 8>(st, p);
      default: return launch_cl_fwd<256, 8, 4, 8>(st, p);
    }
  }
  tg_set_error("gru_cl_fwd: hidden size %d not supported", H);
  return TG_ERR_UNSUPPORTED;
}

int tg_gru_cl_bwd(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y, const float* whh,
                  float* dgi, float* dq, int B, int T, int H, int dy_last) {
  ClBwdParams p{dy, rzn, q, y, whh, dgi, dq, B, T, dy_last};
  ClCaps& c = cl_caps();
  if (H == 128 && tg_cluster_no() >= 4 && tg_cluster_dio())
    return pick_groups(B, 4, c.b128, 2) == 1 ? launch_cl_bwd<128, 2, 1, CL_PF, true, 4>(st, p) : launch_cl_bwd<128, 2, 2, CL_PF, true, 4>(st, p);
  if (H == 128 && tg_cluster_no() >= 4)
    return pick_groups(B, 4, c.b128, 2) == 1 ? launch_cl_bwd<128, 2, 1, CL_PF, false, 4>(st, p) : launch_cl_bwd<128, 2, 2, CL_PF, false, 4>(st, p);
  if (H == 128 && tg_cluster_dio())
    return pick_groups(B, 4, c.b128, 2) == 1 ? launch_cl_bwd<128, 2, 1, CL_PF, true>(st, p) : launch_cl_bwd<128, 2, 2, CL_PF, true>(st, p);
  if (H == 128) return pick_groups(B, 4, c.b128, 2) == 1 ? launch_cl_bwd<128, 2, 1>(st, p) : launch_cl_bwd<128, 2, 2>(st, p);
  if (H == 256 && tg_cluster_no() >= 4) {
    switch (pick_groups(B, 8, c.b256, 3)) {
      case 1: return launch_cl_bwd<256, 8, 1, CL_PF, false, 4>(st, p);
      case 2: return launch_cl_bwd<256, 8, 2, CL_PF, false, 4>(st, p);
      default: return launch_cl_bwd<256, 8, 3, 2, false, 4>(st, p);
    }
  }
  if (H == 256) {
    switch (pick_groups(B, 8, c.b256, 3)) {
      case 1: return launch_cl_bwd<256, 8, 1>(st, p);
      case 2: return launch_cl_bwd<256, 8, 2>(st, p);
      default: return launch_cl_bwd<256, 8, 3, 2>(st, p);     // 218 KB of shared memory: a 2-deep input ring
    }
  }
  tg_set_error("gru_cl_bwd: hidden size %d not supported", H);
  return TG_ERR_UNSUPPORTED;
}

// reverse-over-tangent at H = 128: the one-SM kernel (gru_jvp.cu) spills and takes 3.5 ms per call at the c3 shape
// H = 256: one sequence group per cluster (the six all-gathered vectors of 8 sequences already take 104 KB of shared
// memory), so only batches that fit ONE wave of resident clusters (15 x 8 = 120 sequences on a B200) go here: 7.9 vs 15.0 ms
// per pass at B = 120, but 19.2 vs 16.7 ms at B = 256 (three waves; the L2-streaming kernel's time hardly depends on B) --
// profiles/r02_probe_jvp256.log
bool tg_cluster_takes_jvp_bwd(int H, int B) {
  const int mode = tg_use_cluster();
  if (mode == 0) return false;
  if (H == 128) return true;
  if (H != 256 || cl_caps().j256[0] <= 0) return false;
  return mode == 2 || (tg_cluster_jvp256() && (B + 7) / 8 <= cl_caps().j256[0]);
}

int tg_gru_cl_jvp_bwd(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                      const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh, float* gib,
                      float* qb, float* gidb, float* qdb, int B, int T, int H, int last_only) {
  ClJbParams p{hbar, hdbar, rzn, q, ta, qdot, y, ydot, whh, gib, qb, gidb, qdb, B, T, last_only};
  if (H == 128 && tg_cluster_no() >= 4)
    return pick_groups(B, 4, cl_caps().j128, 2) == 1 ? launch_cl_jb<128, 2, 1, 4>(st, p) : launch_cl_jb<128, 2, 2, 4>(st, p);
  if (H == 128) return pick_groups(B, 4, cl_caps().j128, 2) == 1 ? launch_cl_jb<128, 2, 1>(st, p) : launch_cl_jb<128, 2, 2>(st, p);
  if (H == 256) return tg_cluster_no() >= 4 ? launch_cl_jb<256, 8, 1, 4>(st, p) : launch_cl_jb<256, 8, 1>(st, p);
  tg_set_error("gru_cl_jvp_bwd: hidden size %d not supported", H);
  return TG_ERR_UNSUPPORTED;
}

// Recurrent kernels for hidden sizes whose W_hh does not fit one SM's registers (H > 128, e.g. the H = 256 point of
// BASELINE config c4): the same four recurrences as gru_fwd.cu / gru_bwd.cu / gru_jvp.cu (SURVEY.md A.1, A.2, A.4,
// formulas pinned by oracle/gru_math.py), with W_hh streamed from L2 every timestep instead of living on-chip.
//   * one CTA (512 threads) owns BT = 4 sequences; the vector to multiply (h_{t-1}, or the dGH row) sits in shared
//     memory, each warp walks whole weight rows with coalesced 512-byte loads and finishes a row with a shuffle
//     reduction, so a weight element fetched once serves all 4 sequences;
//   * the backward kernels take W_hh^T (H x 3H, made once per call on the host side) so that they walk rows too;
//   * pointwise phases run one (sequence, unit) pair per thread with coalesced global accesses.
// This is the capacity fallback, not the fast path: W_hh (768 KB at H = 256) is re-read from L2 every step.  The
// cluster/DSMEM variant that keeps it distributed over 4 SMs is the planned replacement.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int NT = 512, NW = NT / 32, BT = 4;

// out[b][row] = sum_c W[row][c] * vin[b][c]   (NV input vectors per sequence share every weight load)
// A warp walks RU rows at a time so that the weight loads of RU rows (L2 latency ~600 clk) are in flight together.
// (Measured: a one-item-deep software pipeline across row groups is slower -- fewer loads in flight -- so the simple
// unroll stays.)
template <int NV>
__device__ __forceinline__ void matvec_rows(const float* __restrict__ W, int rows, int cols, const float* vin,
                                            int vin_stride, float* out, int out_stride) {
  constexpr int RU = (NV == 1) ? 8 : 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c4n = cols >> 2;
  for (int row0 = warp * RU; row0 < rows; row0 += NW * RU) {
    float2 acc[RU][NV][BT];          // packed FFMA2: .x sums the even columns of a float4 pair, .y the odd ones
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int b = 0; b < BT; ++b) acc[u][v][b] = make_float2(0.f, 0.f);
    for (int c4 = lane; c4 < c4n; c4 += 32) {
      float4 w4[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int row = min(row0 + u, rows - 1);      // clamped rows are computed and dropped
        w4[u] = __ldg(reinterpret_cast<const float4*>(W + (size_t)row * cols) + c4);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 x4 = reinterpret_cast<const float4*>(vin + (size_t)(v * BT + b) * vin_stride)[c4];
          const float2 x01 = make_float2(x4.x, x4.y), x23 = make_float2(x4.z, x4.w);
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            acc[u][v][b] = __ffma2_rn(make_float2(w4[u].x, w4[u].y), x01, acc[u][v][b]);
            acc[u][v][b] = __ffma2_rn(make_float2(w4[u].z, w4[u].w), x23, acc[u][v][b]);
          }
        }
    }
    // RU * NV * BT == 32 partial sums per lane: a transposing butterfly (16 + 8 + 4 + 2 + 1 shuffles instead of
    // 32 x 5) leaves the total of flat index l = (u * NV + v) * BT + b in lane l -- the shuffle unit, not L2, was
    // the limit of this loop -- and the 32 results go out in one store instruction.
    static_assert(RU * NV * BT == 32, "one total per lane");
    float vals[32];
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int b = 0; b < BT; ++b) vals[(u * NV + v) * BT + b] = acc[u][v][b].x + acc[u][v][b].y;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float send = up ? vals[i] : vals[i + o];
        const float keep = up ? vals[i + o] : vals[i];
        vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    {
      const int u = lane / (NV * BT), v = (lane / BT) % NV, b = lane % BT;
      if (row0 + u < rows) out[(size_t)(v * BT + b) * out_stride + row0 + u] = vals[0];
    }
  }
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(NT, 1) bigh_fwd_kernel(float* gi, const float* __restrict__ whh,
                                                         const float* __restrict__ bhh, float* y, float* q, int B, int T,
                                                         int H, int save) {
  extern __shared__ __align__(16) float sm[];
  float* hs = sm;                 // [BT][H]
  float* gh = sm + BT * H;        // [BT][3H]
  const int b0 = blockIdx.x * BT, nb = min(BT, B - b0);
  for (int i = threadIdx.x; i < BT * H; i += NT) hs[i] = 0.f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    matvec_rows<1>(whh, 3 * H, H, hs, H, gh, 3 * H);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * H; i += NT) {
      const int b = i / H, j = i - b * H;
      const size_t row = (size_t)(b0 + b) * T + t;
      float* g = gi + row * 3 * H;
      const float* a = gh + b * 3 * H;
      const float r = sigmoid_mufu(g[j] + a[j] + bhh[j]);
      const float z = sigmoid_mufu(g[H + j] + a[H + j] + bhh[H + j]);
      const float qv = a[2 * H + j] + bhh[2 * H + j];
      const float n = tanh_mufu(fmaf(r, qv, g[2 * H + j]));
      const float h = fmaf(z, hs[b * H + j] - n, n);
      hs[b * H + j] = h;          // only this thread touches (b, j); the mat-vec reads it after the barrier
      y[row * H + j] = h;
      if (save) { g[j] = r; g[H + j] = z; g[2 * H + j] = n; q[row * H + j] = qv; }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ BPTT
__global__ void __launch_bounds__(NT, 1) bigh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ rzn,
                                                         const float* __restrict__ q, const float* __restrict__ y,
                                                         const float* __restrict__ whh_t, float* dgi, float* dq, int B,
                                                         int T, int H, int dy_last) {
  extern __shared__ __align__(16) float sm[];
  float* dg = sm;                  // [BT][3H]  dGH of this step
  float* mv = sm + BT * 3 * H;     // [BT][H]   dGH W_hh
  float* cr = mv + BT * H;         // [BT][H]   carried dh
  const int b0 = blockIdx.x * BT, nb = min(BT, B - b0);
  for (int i = threadIdx.x; i < BT * 3 * H; i += NT) dg[i] = 0.f;
  for (int i = threadIdx.x; i < BT * H; i += NT) cr[i] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    for (int i = threadIdx.x; i < nb * H; i += NT) {
      const int b = i / H, k = i - b * H;
      const size_t row = (size_t)(b0 + b) * T + t;
      const float* s = rzn + row * 3 * H;
      const float r = s[k], z = s[H + k], n = s[2 * H + k], qv = q[row * H + k];
      const float hp = (t > 0) ? y[(row - 1) * H + k] : 0.f;
      float dyv;
      if (dy_last) dyv = (t == T - 1) ? dy[(size_t)(b0 + b) * H + k] : 0.f;
      else dyv = dy[row * H + k];
      const float dh = dyv + cr[b * H + k];
      const float dan = dh * (1.f - z) * (1.f - n * n);
      const float daz = dh * (hp - n) * z * (1.f - z);
      const float dar = dan * qv * r * (1.f - r);
      const float dqv = dan * r;
      cr[b * H + k] = dh * z;
      float* o = dgi + row * 3 * H;
      o[k] = dar; o[H + k] = daz; o[2 * H + k] = dan;
      dq[row * H + k] = dqv;
      float* d = dg + b * 3 * H;
      d[k] = dar; d[H + k] = daz; d[2 * H + k] = dqv;
    }
    __syncthreads();
    matvec_rows<1>(whh_t, H, 3 * H, dg, 3 * H, mv, H);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * H; i += NT) cr[i] += mv[i];
    // (b, k) pairs are owned by fixed threads, so no barrier is needed before the next pointwise phase
  }
}

// ------------------------------------------------------------------ tangent forward (R1)
__global__ void __launch_bounds__(NT, 1) bigh_jvp_fwd_kernel(float* gid, const float* __restrict__ rzn,
                                                             const float* __restrict__ q, const float* __restrict__ y,
                                                             const float* __restrict__ whh, float* ydot, float* qdot,
                                                             int B, int T, int H) {
  extern __shared__ __align__(16) float sm[];
  float* hd = sm;               // [BT][H]  tangent state
  float* gh = sm + BT * H;      // [BT][3H]
  const int b0 = blockIdx.x * BT, nb = min(BT, B - b0);
  for (int i = threadIdx.x; i < BT * H; i += NT) hd[i] = 0.f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    matvec_rows<1>(whh, 3 * H, H, hd, H, gh, 3 * H);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * H; i += NT) {
      const int b = i / H, j = i - b * H;
      const size_t row = (size_t)(b0 + b) * T + t;
      float* g = gid + row * 3 * H;
      const float* s = rzn + row * 3 * H;
      const float* a = gh + b * 3 * H;
      const float r = s[j], z = s[H + j], n = s[2 * H + j], qv = q[row * H + j];
      const float hp = (t > 0) ? y[(row - 1) * H + j] : 0.f;
      const float a_r = g[j] + a[j], a_z = g[H + j] + a[H + j], qd = a[2 * H + j];
      const float rdot = r * (1.f - r) * a_r, zdot = z * (1.f - z) * a_z;
      const float a_n = g[2 * H + j] + rdot * qv + r * qd;
      const float ndot = (1.f - n * n) * a_n;
      const float v = (1.f - z) * ndot + z * hd[b * H + j] + zdot * (hp - n);
      hd[b * H + j] = v;
      g[j] = a_r; g[H + j] = a_z; g[2 * H + j] = a_n;
      qdot[row * H + j] = qd;
      ydot[row * H + j] = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ reverse over (primal + tangent) (R1)
__global__ void __launch_bounds__(NT, 1)
bigh_jvp_bwd_kernel(const float* __restrict__ hbar, const float* __restrict__ hdbar, const float* __restrict__ rzn,
                    const float* __restrict__ q, const float* __restrict__ ta, const float* __restrict__ qdot,
                    const float* __restrict__ y, const float* __restrict__ ydot, const float* __restrict__ whh_t,
                    float* gib, float* qb_out, float* gidb, float* qdb_out, int B, int T, int H, int last_only) {
  extern __shared__ __align__(16) float sm[];
  float* dg = sm;                      // [2][BT][3H] : primal dGH, tangent dGH
  float* mv = sm + 2 * BT * 3 * H;     // [2][BT][H]
  float* ch = mv + 2 * BT * H;         // [2][BT][H]  carries (h, hdot)
  const int b0 = blockIdx.x * BT, nb = min(BT, B - b0);
  for (int i = threadIdx.x; i < 2 * BT * 3 * H; i += NT) dg[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * BT * H; i += NT) ch[i] = 0.f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    for (int i = threadIdx.x; i < nb * H; i += NT) {
      const int b = i / H, k = i - b * H;
      const size_t row = (size_t)(b0 + b) * T + t;
      const float* s = rzn + row * 3 * H;
      const float* tp = ta + row * 3 * H;
      const float rt = s[k], zt = s[H + k], nt = s[2 * H + k], qt = q[row * H + k];
      const float art = tp[k], azt = tp[H + k], ant = tp[2 * H + k], qdt = qdot[row * H + k];
      const float hp = (t > 0) ? y[(row - 1) * H + k] : 0.f;
      const float hdp = (t > 0) ? ydot[(row - 1) * H + k] : 0.f;
      float hb, hdb;
      if (last_only) {
        hb = (t == T - 1) ? hbar[(size_t)(b0 + b) * H + k] : 0.f;
        hdb = (t == T - 1) ? hdbar[(size_t)(b0 + b) * H + k] : 0.f;
      } else {
        hb = hbar[row * H + k];
        hdb = hdbar[row * H + k];
      }
      hb += ch[b * H + k];
      hdb += ch[(BT + b) * H + k];
      const float sr = rt * (1.f - rt), sz = zt * (1.f - zt), sn = 1.f - nt * nt;
      const float rdot = sr * art, zdot = sz * azt, ndot = sn * ant;
      const float ndb = (1.f - zt) * hdb;
      float zb = hdb * (hdp - ndot);
      const float zdb = hdb * (hp - nt);
      float nb_ = -zdot * hdb;
      float nh = zdot * hdb;
      const float nhd = zt * hdb;
      nb_ += (1.f - zt) * hb;
      zb += hb * (hp - nt);
      nh += zt * hb;
      const float anb_d = sn * ndb;
      nb_ -= 2.f * nt * ant * ndb;
      const float rdb = qt * anb_d;
      float qb = rdot * anb_d;
      float rb = qdt * anb_d;
      const float qdb = rt * anb_d;
      const float anb = sn * nb_;
      rb += qt * anb;
      qb += rt * anb;
      const float azb_d = sz * zdb;
      zb += (1.f - 2.f * zt) * azt * zdb;
      const float arb_d = sr * rdb;
      rb += (1.f - 2.f * rt) * art * rdb;
      const float azb = sz * zb;
      const float arb = sr * rb;
      ch[b * H + k] = nh;
      ch[(BT + b) * H + k] = nhd;
      float* o = gib + row * 3 * H;
      o[k] = arb; o[H + k] = azb; o[2 * H + k] = anb;
      qb_out[row * H + k] = qb;
      float* od = gidb + row * 3 * H;
      od[k] = arb_d; od[H + k] = azb_d; od[2 * H + k] = anb_d;
      qdb_out[row * H + k] = qdb;
      float* d0 = dg + b * 3 * H;
      d0[k] = arb; d0[H + k] = azb; d0[2 * H + k] = qb;
      float* d1 = dg + (BT + b) * 3 * H;
      d1[k] = arb_d; d1[H + k] = azb_d; d1[2 * H + k] = qdb;
    }
    __syncthreads();
    matvec_rows<2>(whh_t, H, 3 * H, dg, 3 * H, mv, H);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * H; i += NT) {
      ch[i] += mv[i];
      ch[BT * H + i] += mv[BT * H + i];
    }
  }
}

int bigh_check(const char* what, int B, int T, int H, size_t smem_floats) {
  TG_REQUIRE(B > 0 && T > 0 && H > 0, TG_ERR_SHAPE, "%s: bad shape B=%d T=%d H=%d", what, B, T, H);
  TG_REQUIRE(H % 4 == 0, TG_ERR_UNSUPPORTED, "%s: hidden size %d > 128 must be a multiple of 4", what, H);
  TG_REQUIRE(smem_floats * 4 <= (size_t)tg_max_optin_smem(), TG_ERR_UNSUPPORTED, "%s: hidden size %d too large", what, H);
  return TG_OK;
}

}  // namespace

int tg_bigh_fwd(cudaStream_t st, float* gi, const float* whh, const float* bhh, float* y, float* q, int B, int T, int H,
                int save) {
  const size_t fl = (size_t)BT * 4 * H;
  int rc = bigh_check("gru_fwd", B, T, H, fl);
  if (rc) return rc;
  TG_OPT_IN_SMEM(bigh_fwd_kernel, "gru_fwd(bigH)");
  bigh_fwd_kernel<<<(B + BT - 1) / BT, NT, fl * 4, st>>>(gi, whh, bhh, y, q, B, T, H, save);
  return tg_check_launch("gru_fwd(bigH)");
}
int tg_bigh_bwd(cudaStream_t st, const float* dy, const float* rzn, const float* q, const float* y, const float* whh_t,
                float* dgi, float* dq, int B, int T, int H, int dy_last) {
  const size_t fl = (size_t)BT * 5 * H;
  int rc = bigh_check("gru_bwd", B, T, H, fl);
  if (rc) return rc;
  TG_OPT_IN_SMEM(bigh_bwd_kernel, "gru_bwd(bigH)");
  bigh_bwd_kernel<<<(B + BT - 1) / BT, NT, fl * 4, st>>>(dy, rzn, q, y, whh_t, dgi, dq, B, T, H, dy_last);
  return tg_check_launch("gru_bwd(bigH)");
}
int tg_bigh_jvp_fwd(cudaStream_t st, float* gid, const float* rzn, const float* q, const float* y, const float* whh,
                    float* ydot, float* qdot, int B, int T, int H) {
  const size_t fl = (size_t)BT * 4 * H;
  int rc = bigh_check("gru_jvp_fwd", B, T, H, fl);
  if (rc) return rc;
  TG_OPT_IN_SMEM(bigh_jvp_fwd_kernel, "gru_jvp_fwd(bigH)");
  bigh_jvp_fwd_kernel<<<(B + BT - 1) / BT, NT, fl * 4, st>>>(gid, rzn, q, y, whh, ydot, qdot, B, T, H);
  return tg_check_launch("gru_jvp_fwd(bigH)");
}
int tg_bigh_jvp_bwd(cudaStream_t st, const float* hbar, const float* hdbar, const float* rzn, const float* q,
                    const float* ta, const float* qdot, const float* y, const float* ydot, const float* whh_t,
                    float* gib, float* qb, float* gidb, float* qdb, int B, int T, int H, int last_only) {
  const size_t fl = (size_t)BT * 10 * H;
  int rc = bigh_check("gru_jvp_bwd", B, T, H, fl);
  if (rc) return rc;
  TG_OPT_IN_SMEM(bigh_jvp_bwd_kernel, "gru_jvp_bwd(bigH)");
  bigh_jvp_bwd_kernel<<<(B + BT - 1) / BT, NT, fl * 4, st>>>(hbar, hdbar, rzn, q, ta, qdot, y, ydot, whh_t, gib, qb, gidb,
                                                            qdb, B, T, H, last_only);
  return tg_check_launch("gru_jvp_bwd(bigH)");
}

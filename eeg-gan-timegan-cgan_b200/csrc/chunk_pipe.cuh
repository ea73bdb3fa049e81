// ChunkPipe: streams per-(sequence,timestep) rows between HBM and a shared-memory ring in chunks of TC
// timesteps, forward or reverse in time, for the persistent recurrent kernels.
//
//  * (B,T,W) batch-first activations: for one sequence a chunk of TC steps is ONE contiguous run of
//    TC*W floats, so every chunk moves with one 1-D bulk async copy (cp.async.bulk, the TMA engine)
//    per sequence per array, completing on an mbarrier (loads) or a bulk group (stores).
//  * "inout" arrays are overwritten in place in shared memory by the compute threads (e.g. gi -> r,z,n)
//    and then stored, so no extra staging buffers are needed.
//  * arrays with shift = -1 deliver row t-1 at step t (h_{t-1}); row -1 is never loaded, the consumer
//    substitutes zero at t == 0.
//  * If H % 4 != 0 or a pointer is not 16-B aligned, bulk copies are illegal; the same interface then
//    falls back to cooperative generic loads/stores by all threads (correct, slower; rare shapes).
#pragma once
#include "common.cuh"

#define TG_STRM_LOAD 1
#define TG_STRM_STORE 2

template <int NS, int BT, int TC, int NST>
struct ChunkPipe {
  float* g[NS];    // load source
  float* gst[NS];  // store destination (may equal g: in place)
  int w[NS];
  int mode[NS];
  int shift[NS];
  int off[NS];
  int bstride[NS];  // floats between consecutive sequences of one array inside a stage (padded, see seq_stride)
  int stage_floats;
  float* stages;
  uint64_t* full;
  int T, nb, b0, NC;
  bool reverse, bulk;

  // Sequence b of an array starts seq_stride() floats after sequence b-1.  The pad makes the stride == 16
  // (mod 32) so that the lanes of a warp that work on two different sequences (same hidden unit) hit
  // different shared-memory banks, and keeps every row 16-B aligned for the bulk copies (pad % 4 == 0
  // whenever TC*w % 4 == 0).
  static __host__ __device__ __forceinline__ int seq_stride(int width) {
    const int n = TC * width;
    return (BT > 1) ? n + ((16 - (n % 32)) + 32) % 32 : n;
  }
  __device__ __forceinline__ void layout() {
    int o = 0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      off[k] = o;
      bstride[k] = seq_stride(w[k]);
      o += BT * bstride[k];
    }
    stage_floats = o;
  }
  static __host__ __device__ int stage_floats_for(const int* widths) {
    int o = 0;
    for (int k = 0; k < NS; ++k) o += BT * seq_stride(widths[k]);
    return o;
  }
  __device__ __forceinline__ int t0_of(int c) const { return (reverse ? (NC - 1 - c) : c) * TC; }
  __device__ __forceinline__ int tcn_of(int c) const {
    int t0 = t0_of(c);
    return min(TC, T - t0);
  }
  // 32-bit shared-window address of a row (for the ld.shared / st.shared helpers of common.cuh)
  __device__ __forceinline__ uint32_t row_addr(int s, int k, int b, int tl) const {
    return smem_u32(stages) + 4u * (uint32_t)(s * stage_floats + off[k] + b * bstride[k] + tl * w[k]);
  }
  __device__ __forceinline__ float* row(int s, int k, int b, int tl) const {
    return stages + (size_t)s * stage_floats + off[k] + b * bstride[k] + tl * w[k];
  }

  // ---- producer side (bulk mode: one thread) ------------------------------------------------
  __device__ __forceinline__ void issue_load(int c) {
    const int s = c % NST;
    const int t0 = t0_of(c), tcn = min(TC, T - t0);
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (mode[k] & TG_STRM_LOAD) {
        int r0 = max(t0 + shift[k], 0);
        int rows = tcn - (r0 - (t0 + shift[k]));
        if (rows > 0) total += (uint32_t)(nb * rows * w[k]) * 4u;
      }
    }
    mbar_expect_tx(&full[s], total);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (mode[k] & TG_STRM_LOAD) {
        int r0 = max(t0 + shift[k], 0);
        int skip = r0 - (t0 + shift[k]);
        int rows = tcn - skip;
        if (rows > 0) {
          for (int b = 0; b < nb; ++b)
            bulk_g2s(row(s, k, b, skip), g[k] + ((size_t)(b0 + b) * T + r0) * w[k], (uint32_t)(rows * w[k]) * 4u,
                     &full[s]);
        }
      }
    }
  }
  __device__ __forceinline__ void issue_store(int c) {
    const int s = c % NST;
    const int t0 = t0_of(c), tcn = min(TC, T - t0);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      if (mode[k] & TG_STRM_STORE) {
        for (int b = 0; b < nb; ++b)
          bulk_s2g(gst[k] + ((size_t)(b0 + b) * T + t0) * w[k], row(s, k, b, 0), (uint32_t)(tcn * w[k]) * 4u);
      }
    }
    bulk_commit();
  }

  // ---- block-level protocol -----------------------------------------------------------------
  // call once by all threads after `full` barriers and `stages` are carved
  __device__ __forceinline__ void start() {
    if (bulk) {
      if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int c = 0; c < NST - 1 && c < NC; ++c) issue_load(c);
      }
    }
  }
  // all threads: block until chunk c is resident in stage c % NST
  __device__ __forceinline__ void acquire(int c) {
    const int s = c % NST;
    if (bulk) {
      mbar_wait(&full[s], (uint32_t)((c / NST) & 1));
    } else {
      const int t0 = t0_of(c), tcn = min(TC, T - t0);
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (mode[k] & TG_STRM_LOAD) {
          int r0 = max(t0 + shift[k], 0);
          int skip = r0 - (t0 + shift[k]);
          int rows = tcn - skip;
          if (rows > 0) {
            int n = rows * w[k];
            for (int b = 0; b < nb; ++b) {
              const float* src = g[k] + ((size_t)(b0 + b) * T + r0) * w[k];
              float* dst = row(s, k, b, skip);
              for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
            }
          }
        }
      }
      __syncthreads();
    }
  }
  // all threads, after the __syncthreads that closes the chunk's last step (in bulk mode every thread
  // must have executed fence_async_smem() before that barrier)
  __device__ __forceinline__ void release(int c) {
    if (bulk) {
      if (threadIdx.x == 0) {
        issue_store(c);
        bulk_wait_read<1>();  // the stage used by chunk c-1 has been read out -> safe to refill
        int cn = c + NST - 1;
        if (cn < NC) issue_load(cn);
      }
    } else {
      const int s = c % NST;
      const int t0 = t0_of(c), tcn = min(TC, T - t0);
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (mode[k] & TG_STRM_STORE) {
          int n = tcn * w[k];
          for (int b = 0; b < nb; ++b) {
            float* dst = gst[k] + ((size_t)(b0 + b) * T + t0) * w[k];
            const float* src = row(s, k, b, 0);
            for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
          }
        }
      }
      __syncthreads();
    }
  }
  __device__ __forceinline__ void drain() {
    if (bulk && threadIdx.x == 0) bulk_wait_all<0>();
  }
};

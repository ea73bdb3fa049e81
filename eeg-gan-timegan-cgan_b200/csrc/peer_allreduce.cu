// One-shot SUM all-reduce over NVLink peer memory (the gradient all-reduce after every
// optimiser step's BPTT and the small batch-statistics all-reduce of the loss forward).
//
// Net-new relative to the reference, which is single-process.  Why not NCCL here: the payloads are tiny
// (<= 1.2 MB per bucket at c2, 4.6 MB at c3), so the cost is launch latency, and an NCCL collective cannot be
// replayed from the CUDA graph that holds the rest of the joint step.  This kernel is an ordinary launch:
//   * it is MULTI-TENSOR: the gradients of one network are gathered straight out of their own tensors and
//     the sums are written straight back (no flatten / all-reduce / scatter-back triple);
//   * every rank owns one cudaMalloc'd "peer region" that all other ranks of the box map through CUDA IPC;
//     a call site (one bucket) has a double-buffered staging area and one flag word per (chunk, source rank);
//   * CTA c stages chunk c (4096 floats) of its rank's contribution, publishes flag[c] = epoch to every peer
//     with a system-scope release store over NVLink, waits for the peers' flag[c], then pulls chunk c from every
//     rank's staging area and adds the W contributions in rank order -- the same order on every rank, so all
//     ranks end up with bit-identical sums (clip + Adam then stay in lock step without a broadcast);
//   * nothing else synchronises: no grid barrier, no host involvement -- so the launch is graph-capturable.  CTA c
//     only ever waits for CTA c of its peers, which waits for nothing else.  The grid is capped at PR_MAX_GRID CTAs
//     (4 per SM: every CTA of the launch can be resident at once); a CTA walks its chunks c, c + grid, c + 2 grid, ...
//     in increasing order, the same order on every rank, so a large bucket (23 MB of gradients at H = 256 = 1 400
//     chunks) does not rely on the order in which the hardware dispatches the CTAs of an oversubscribed grid.
//   * the epoch of a call site lives in device memory and is advanced by the last CTA of each launch, so a
//     replayed graph keeps counting.  Staging is double-buffered by epoch parity: a rank can only overwrite
//     parity p again after every peer has signalled the epoch in between, i.e. has finished reading p.
// A spin that exceeds the timeout (default 10 s, tg_set_option("peer_timeout_ms")) means a peer is gone: the kernel
// sets *status and POISONS its output chunk with NaN instead of summing stale staging data -- the step's losses
// and weights turn non-finite at once and dist.PeerComm.check_status() (polled by the training loop at every
// flush / checkpoint) raises, rather than the replicas silently drifting apart.
#include "../../include/timegan_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "losses.h"
#include <string.h>

namespace {

constexpr int PR_THREADS = 256;
constexpr int PR_CHUNK = 4096;  // floats per chunk
constexpr int PR_MAX_GRID = 592;   // 4 CTAs per SM on a 148-SM part: all of a launch's CTAs are co-resident

struct PeerArgs {
  float* t[TG_MT_MAX];
  long long size[TG_MT_MAX];
  int blk_start[TG_MT_MAX + 1];
  int n;
  float* data[TG_PEER_MAX];          // rank r's staging area of this call site: [2][nchunks * PR_CHUNK]
  unsigned int* flags[TG_PEER_MAX];  // rank r's flag words of this call site: [nchunks][world]
  unsigned int* epoch;               // local {epoch, finished-CTA counter}
  unsigned int* status;              // local error word (0 = ok)
  int rank, world, nchunks;
  long long timeout_clk;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__global__ void __launch_bounds__(PR_THREADS) peer_allreduce_kernel(const __grid_constant__ PeerArgs a) {
  const int tid = threadIdx.x;
  const unsigned int e = *reinterpret_cast<volatile unsigned int*>(a.epoch) + 1u;
  __shared__ int failed;
  for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
    int t = 0;
    while (t + 1 < a.n && c >= a.blk_start[t + 1]) ++t;
    const long long base = (long long)(c - a.blk_start[t]) * PR_CHUNK;
    const int len = (int)min((long long)PR_CHUNK, a.size[t] - base);
    float* __restrict__ tens = a.t[t] + base;
    const size_t slot = (size_t)(e & 1u) * a.nchunks * PR_CHUNK + (size_t)c * PR_CHUNK;
    float* mine = a.data[a.rank] + slot;
    const bool vec = ((reinterpret_cast<uintptr_t>(tens) & 15u) == 0) && (len % 4 == 0);

    // 1. stage this rank's contribution
    if (vec) {
      for (int i = tid * 4; i < len; i += PR_THREADS * 4)
        *reinterpret_cast<float4*>(mine + i) = *reinterpret_cast<const float4*>(tens + i);
    } else {
      for (int i = tid; i < len; i += PR_THREADS) mine[i] = tens[i];
    }
    if (tid == 0) failed = 0;
    __threadfence_system();
    __syncthreads();

    // 2. publish to every peer, then wait for every peer's chunk c
    if (tid < a.world && tid != a.rank) {
      st_release_sys(a.flags[tid] + (size_t)c * a.world + a.rank, e);
      const unsigned int* f = a.flags[a.rank] + (size_t)c * a.world + tid;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(f) - e) < 0) {
        if (clock64() - t0 > a.timeout_clk) {  // a peer is gone; report instead of hanging the device
          atomicExch(a.status, 1u);
          failed = 1;
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();

    // 3. pull and add in rank order (identical on every rank => bit-identical results everywhere); after a timeout
    //    the staging slots of the missing peer hold the data of two epochs ago -- never sum those
    if (failed) {
      const float nan = __int_as_float(0x7fc00000);
      for (int i = tid; i < len; i += PR_THREADS) tens[i] = nan;
    } else if (vec) {
      for (int i = tid * 4; i < len; i += PR_THREADS * 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < a.world; ++r) {
          const float4 v = ld_relaxed_sys_v4(a.data[r] + slot + i);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(tens + i) = s;
      }
    } else {
      for (int i = tid; i < len; i += PR_THREADS) {
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) s += ld_relaxed_sys(a.data[r] + slot + i);
        tens[i] = s;
      }
    }
    __syncthreads();        // `failed` is rewritten by the next chunk
  }

  // 4. the last CTA of the launch advances the call site's epoch
  if (tid == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(a.epoch + 1, 1u);
    if (done == gridDim.x - 1) {
      a.epoch[1] = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(a.epoch) = e;
    }
  }
}

}  // namespace

extern "C" {

int tg_peer_chunk_floats(void) { return PR_CHUNK; }

int tg_peer_alloc(void** ptr, size_t bytes) {
  TG_REQUIRE(ptr && bytes > 0, TG_ERR_ARG, "peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { tg_set_error("peer_alloc(%zu): %s", bytes, cudaGetErrorString(e)); return (int)e; }
  return TG_OK;
}

int tg_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { tg_set_error("peer_free: %s", cudaGetErrorString(e)); return (int)e; }
  return TG_OK;
}

int tg_peer_export(void* ptr, unsigned char* handle64) {
  TG_REQUIRE(ptr && handle64, TG_ERR_ARG, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { tg_set_error("peer_export: %s", cudaGetErrorString(e)); return (int)e; }
  memcpy(handle64, &h, 64);
  return TG_OK;
}

int tg_peer_open(const unsigned char* handle64, void** peer_ptr) {
  TG_REQUIRE(handle64 && peer_ptr, TG_ERR_ARG, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { tg_set_error("peer_open: %s", cudaGetErrorString(e)); return (int)e; }
  return TG_OK;
}

int tg_peer_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) { tg_set_error("peer_close: %s", cudaGetErrorString(e)); return (int)e; }
  return TG_OK;
}

size_t tg_peer_site_bytes(int n, const long long* sizes, int world, size_t* flag_bytes) {
  long long chunks = 0;
  for (int i = 0; i < n; ++i) chunks += tg_ceil_div(sizes[i], PR_CHUNK);
  if (flag_bytes) *flag_bytes = (size_t)chunks * world * sizeof(unsigned int);
  return (size_t)2 * chunks * PR_CHUNK * sizeof(float);
}

int tg_peer_allreduce(void* stream, int rank, int world, void* const* regions, size_t data_off, size_t flag_off,
                      unsigned int* epoch, unsigned int* status, int n, float* const* tensors,
                      const long long* sizes) {
  TG_REQUIRE(world >= 2 && world <= TG_PEER_MAX && rank >= 0 && rank < world, TG_ERR_ARG,
             "peer_allreduce: bad rank/world %d/%d (max %d ranks)", rank, world, TG_PEER_MAX);
  TG_REQUIRE(regions && tensors && sizes && n > 0 && n <= TG_MT_MAX, TG_ERR_ARG,
             "peer_allreduce: bad tensor list (n=%d, max %d per call)", n, TG_MT_MAX);
  TG_REQUIRE(epoch && status, TG_ERR_ARG, "peer_allreduce: null epoch/status pointer");
  TG_REQUIRE(data_off % 16 == 0 && flag_off % 4 == 0, TG_ERR_ALIGN, "peer_allreduce: misaligned offsets");
  PeerArgs a{};
  int blk = 0;
  for (int i = 0; i < n; ++i) {
    TG_REQUIRE(tensors[i] && sizes[i] > 0, TG_ERR_ARG, "peer_allreduce: tensor %d null/empty", i);
    a.t[i] = tensors[i];
    a.size[i] = sizes[i];
    a.blk_start[i] = blk;
    blk += tg_ceil_div(sizes[i], PR_CHUNK);
  }
  a.blk_start[n] = blk;
  a.n = n;
  for (int r = 0; r < world; ++r) {
    TG_REQUIRE(regions[r], TG_ERR_ARG, "peer_allreduce: region of rank %d is null", r);
    a.data[r] = reinterpret_cast<float*>(static_cast<char*>(regions[r]) + data_off);
    a.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(regions[r]) + flag_off);
  }
  a.epoch = epoch;
  a.status = status;
  a.rank = rank; a.world = world; a.nchunks = blk;
  a.timeout_clk = (long long)tg_peer_timeout_ms() * 2000000LL;      // ~2 GHz SM clock
  peer_allreduce_kernel<<<blk < PR_MAX_GRID ? blk : PR_MAX_GRID, PR_THREADS, 0, (cudaStream_t)stream>>>(a);
  return tg_check_launch("peer_allreduce");
}

}  // extern "C"

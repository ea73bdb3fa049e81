// Shared helpers for the TimeGAN sm_100a kernels: error plumbing, PTX wrappers for
// mbarrier / cp.async.bulk (the 1-D TMA path), warp reductions and accurate activations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

// ---- status codes (mirrors include/timegan_b200.h) -------------------------------------------
#define TG_OK 0
#define TG_ERR_ARG (-1)
#define TG_ERR_SHAPE (-2)
#define TG_ERR_ALIGN (-3)
#define TG_ERR_UNSUPPORTED (-4)

void tg_set_error(const char* fmt, ...);
int tg_check_launch(const char* what);   // returns 0 or positive cudaError_t, sets last error

#define TG_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      tg_set_error(__VA_ARGS__);               \
      return (code);                           \
    }                                          \
  } while (0)

// Opt a kernel in to the device's maximum dynamic shared memory, once per (kernel, device).  The attribute is a
// process-wide property of the function: setting it to "what this launch needs" from two host threads (forward on
// the main thread, backward on autograd's) lets the smaller request silently lower the limit for the other.
int tg_max_optin_smem();
#define TG_OPT_IN_SMEM(kern, what)                                                                          \
  do {                                                                                                      \
    static std::atomic<unsigned> _tg_done{0};                                                               \
    int _dev = 0;                                                                                           \
    cudaGetDevice(&_dev);                                                                                   \
    const unsigned _bit = 1u << (_dev & 31);                                                                \
    if (!(_tg_done.load(std::memory_order_acquire) & _bit)) {                                               \
      cudaError_t _e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                            tg_max_optin_smem());                                           \
      if (_e != cudaSuccess) {                                                                              \
        tg_set_error("%s: cannot opt in to %d B of shared memory: %s", what, tg_max_optin_smem(),           \
                     cudaGetErrorString(_e));                                                               \
        return (int)_e;                                                                                     \
      }                                                                                                     \
      _tg_done.fetch_or(_bit, std::memory_order_release);                                                   \
    }                                                                                                       \
  } while (0)

static inline bool tg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 1-D bulk async copy global -> shared (TMA engine, completion on an mbarrier). 16-B aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 1-D bulk async copy shared -> global (bulk-group completion).
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// accurate (non fast-math) activations: fp32 parity target is 1e-4 normwise after 768 recurrent steps
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) { return tanhf(x); }

// shared-memory accesses through 32-bit shared-window addresses: the recurrent kernels keep their ring / state
// addresses in registers instead of re-deriving generic pointers every timestep
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
// one bf16 value from shared memory, widened to fp32 (exact)
__device__ __forceinline__ float lds_bf16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid / tanh as one MUFU.EX2 + one MUFU.RCP each, no range fix-ups (saturate correctly at +-inf)
__device__ __forceinline__ float sigmoid_mufu(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_mufu(float x) {
  return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
}

// fast activations for the recurrent kernels: MUFU.EX2 + MUFU.RCP (relative error ~2^-21); the parity tests
// hold the 1e-4 normwise bound against torch.nn.GRU over 768 steps with these.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

static inline int tg_ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// Counter-based noise (Philox4x32-10) for the training step's random draws.
//
// Reference: sample_noise (train_timegan.py:64-65, U(0,1) of shape (B,T,z)), add_instance_noise
// (tt:46-47, h + std*N(0,1)) and smooth_labels (tt:40-43).  The reference draws from the device's global
// generator; parity runs inject the reference's own CPU draws instead, so these
// kernels only have to be statistically equivalent, reproducible from (seed, offset) and HBM-bound.
#include "common.cuh"
#include "kernels.h"
#include "losses.h"

namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// (0,1]-open-at-zero uniform from 24 random bits -> [0,1) like torch.rand: k / 2^24
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// `ctr` (optional, device memory) is added to the call's offset: a CUDA graph bakes `offset` in, the device
// counter advances between replays, so every replay draws fresh numbers from the same Philox stream.
__global__ void uniform_kernel(float* __restrict__ out, long long n, uint64_t seed, uint64_t offset, float lo, float hi,
                               const unsigned long long* __restrict__ ctr) {
  if (ctr) offset += *ctr;
  const long long n4 = (n + 3) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    uint32_t r[4];
    philox4x32_10(offset + (uint64_t)q, seed, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long i = q * 4 + j;
      if (i < n) out[i] = lo + (hi - lo) * u01(r[j]);
    }
  }
}

__global__ void add_normal_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float std,
                                  uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ ctr,
                                  const float* __restrict__ std_dev) {
  if (ctr) offset += *ctr;
  if (std_dev) std = *std_dev;     // CUDA-graph replay: the instance-noise level decays every step, the graph does not change
  const long long n4 = (n + 3) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    uint32_t r[4];
    philox4x32_10(offset + (uint64_t)q, seed, r);
    // Box-Muller on two pairs
    float z[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float u1 = 1.0f - u01(r[2 * j]);  // (0,1]
      const float u2 = u01(r[2 * j + 1]);
      const float rad = sqrtf(-2.0f * logf(u1));
      float s, c;
      sincospif(2.0f * u2, &s, &c);
      z[2 * j] = rad * c; z[2 * j + 1] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long i = q * 4 + j;
      if (i < n) out[i] = (in ? in[i] : 0.f) + std * z[j];
    }
  }
}

int blocks_for(long long n4) {
  long long b = (n4 + 255) / 256;
  long long cap = (long long)tg_num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int tg_rng_uniform_impl(cudaStream_t st, float* out, long long n, unsigned long long seed, unsigned long long offset,
                        float lo, float hi, const unsigned long long* ctr) {
  TG_REQUIRE(out && n > 0, TG_ERR_ARG, "rng_uniform: bad arguments");
  uniform_kernel<<<blocks_for((n + 3) / 4), 256, 0, st>>>(out, n, seed, offset, lo, hi, ctr);
  return tg_check_launch("rng_uniform");
}

int tg_rng_add_normal_impl(cudaStream_t st, const float* in, float* out, long long n, float std,
                           unsigned long long seed, unsigned long long offset, const unsigned long long* ctr,
                           const float* std_dev) {
  TG_REQUIRE(out && n > 0, TG_ERR_ARG, "rng_add_normal: bad arguments");
  add_normal_kernel<<<blocks_for((n + 3) / 4), 256, 0, st>>>(in, out, n, std, seed, offset, ctr, std_dev);
  return tg_check_launch("rng_add_normal");
}

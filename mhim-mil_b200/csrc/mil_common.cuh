// Shared helpers for libmhimk.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "mhimk.h"

namespace mil {

void set_error(const char* fmt, ...);

#define MIL_CHECK_ARG(cond, ...)                   \
  do {                                             \
    if (!(cond)) {                                 \
      ::mil::set_error(__VA_ARGS__);               \
      return -1;                                   \
    }                                              \
  } while (0)

#define MIL_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::mil::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                      \
    }                                                                                      \
  } while (0)

#define MIL_LAUNCH_CHECK()                                                                 \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::mil::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                      \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case MIL_ACT_RELU: return fmaxf(x, 0.f);
    case MIL_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case MIL_ACT_TANH: return tanhf(x);
    case MIL_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

template <int ACT>
__device__ __forceinline__ float act_apply_t(float x) {
  if (ACT == MIL_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == MIL_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  if (ACT == MIL_ACT_TANH) return tanhf(x);
  if (ACT == MIL_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 keep words of the in-kernel dropout (see mil_umma.cuh / mhimk.h mil_dropout_t)
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__host__ __device__ __forceinline__ uint32_t philox_keep_word(uint32_t row, uint32_t chunk, uint32_t thresh, const uint32_t (&seed)[2], const uint32_t (&off)[2]) {
  uint32_t m = 0;
#pragma unroll
  for (uint32_t q4 = 0; q4 < 4; ++q4) {
    uint32_t c[4] = {row, chunk * 4u + q4, off[0], off[1]};
    philox4x32_10(c, seed[0], seed[1]);
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
      m |= ((c[j] & 0xFFFFu) < thresh ? 1u : 0u) << (q4 * 8u + 2u * j);
      m |= ((c[j] >> 16) < thresh ? 1u : 0u) << (q4 * 8u + 2u * j + 1u);
    }
  }
  return m;
}

int num_sms();

}  // namespace mil

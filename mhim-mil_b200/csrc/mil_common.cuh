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

int num_sms();

}  // namespace mil

// Shared helpers for the graspldm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/graspldm_b200.h"

namespace gldm {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Reports a launch failure through the C ABI instead of the reference's fprintf + exit(-1)
// (R/cuda_utils.cuh:28-37).
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GLDM_ECUDA;
  }
  count_launch();
  return GLDM_OK;
}

#define GLDM_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      gldm::set_error(__VA_ARGS__);  \
      return GLDM_EINVAL;            \
    }                                \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: one opt-in per (kernel, device), checked.
struct SmemOptIn { unsigned long long done = 0; };   // bit d = already set on device d
template <typename K>
inline int opt_in_smem(SmemOptIn& st, K kernel, int bytes, const char* what) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && ((st.done >> dev) & 1ull)) return GLDM_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cannot opt in to %d bytes of dynamic shared memory on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
    return GLDM_ECUDA;
  }
  if (dev >= 0 && dev < 64) st.done |= 1ull << dev;
  return GLDM_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// d = fma(dz, dz, fma(dx, dx, dy*dy)): the association nvcc emits for `dx*dx + dy*dy + dz*dz` in the
// reference kernels (FMUL dy*dy; FFMA dx; FFMA dz - read off the sm_100a SASS of the reference build).
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

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

}  // namespace gldm

// Shared helpers for the cddmsl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cddmsl_b200.h"

namespace cddmsl {

extern unsigned long long g_launch_count;  // host-side counter, see cddmsl_launch_count()

inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

#define CDDMSL_CHECK_LAUNCH()                 \
  do {                                        \
    cudaError_t e__ = cudaGetLastError();     \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

#define CDDMSL_CUDA(call)                     \
  do {                                        \
    cudaError_t e__ = (call);                 \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

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

// Number of SMs of the current device (cached).
int sm_count();

}  // namespace cddmsl

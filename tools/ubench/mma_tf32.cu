// Micro-benchmark: issue rate of the legacy warp-level tensor path (mma.sync.m16n8k8 tf32) on sm_100a.
// Decides whether a dense tensor-core formulation of ROIAlign's horizontal pass is worth building.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) mma_loop(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 11 * threadIdx.x,
           b1 = 13 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
          : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  for (int warps_per_sm : {4, 8, 16, 32}) {
    const int blocks = 148 * warps_per_sm / 8;
    const int iters = 20000;
    mma_loop<<<blocks, 256>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_loop<<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * 8 * iters * 8;
    printf("warps/SM %2d: %.3f ms, %.1f G mma/s, %.1f TFLOP/s tf32 (m16n8k8 = 2048 flop), %.2f cycles per mma per SM sub-partition @1.965 GHz\n",
           warps_per_sm, ms, mmas / ms / 1e6, mmas * 2048 / ms / 1e9, 1.965e6 * ms / (mmas / (148 * 4)));
  }
  return 0;
}

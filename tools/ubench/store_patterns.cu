// Microbenchmark: how fast can a persistent 148 x 512-thread grid write the [R][C][14][14] pooled tensor with
// different per-warp store patterns?  (Picks the output path of roi_align_pr.cu.)
//   0: 2 x 56 B per STG.32 (slot = channel, lanes = 14 bins)        1: 112 B contiguous per STG.32 (2 rows of a channel)
//   2: STG.128, 28 lanes = 4 channels x 112 B                        3: fully coalesced 128 B per STG.32 (ideal)
//   4: [16][28] smem tile + TMA tensor store, no proxy fence         5: same with fence.proxy.async per tile
//   6: smem tile, re-read with LDS.128 and written with STG.128 (pattern 2) + 2 __syncwarp per tile
//   7: STS tile [16 ch][56] (4 rows) + LDS.128/STG.128 224 B per channel
//   8: STG.128, 784 B contiguous per channel                        9: [8 ch][196] tile in smem (STS like the kernel)
//  10: tile 9 + ONE 1-D bulk store of 6272 B (fence + wait)             + LDS.128 / STG.128 of the 6272 contiguous bytes
//  11: two tiles per visit (16 channels) as in 10
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int C = 1024, PER = 196;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* __restrict__ out, const __grid_constant__ CUtensorMap map, int R,
                                            int* counter) {
  extern __shared__ __align__(128) float stage_all[];
  constexpr int TS = MODE == 7 ? 16 * 56 : 16 * 28;  // modes 9-11: 1568 floats per warp
  __shared__ int s_u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane >> 4, p = lane & 15;
  const int ngroups = C / 16;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_u = atomicAdd(counter, 1);
    __syncthreads();
    const int u = s_u;  // unit = (chunk of 256 RoIs, channel group)
    const int nunits = (R / 256) * ngroups;
    if (u >= nunits) break;
    const int j = u / ngroups, cg = u % ngroups;
    for (int i = warp; i < 256; i += 16) {
      const int r = j * 256 + i;
      float* base = out + ((size_t)r * C + cg * 16) * PER;
      const float v = (float)(r + lane);
      if (MODE == 0) {
        for (int ph = 0; ph < 14; ++ph)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            if (p < 14) base[(size_t)(2 * kk + slot) * PER + ph * 14 + p] = v + kk;
      } else if (MODE == 1) {
        for (int pp = 0; pp < 7; ++pp)
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (lane < 28) base[(size_t)c * PER + pp * 28 + lane] = v + c;
      } else if (MODE == 2) {
        for (int pp = 0; pp < 7; ++pp)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (lane < 28)
              *reinterpret_cast<float4*>(base + (size_t)(4 * q + lane / 7) * PER + pp * 28 + 4 * (lane % 7)) =
                  make_float4(v, v + 1, v + 2, v + q);
      } else if (MODE == 3) {
        for (int e = lane; e < 16 * PER; e += 32) base[e] = v;
      } else if (MODE == 4 || MODE == 5) {
        float* st = stage_all + warp * TS;
        for (int pp = 0; pp < 7; ++pp) {
          if (pp) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          for (int row = 0; row < 2; ++row)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (p < 14) st[((kk & 3) + 8 * (kk >> 2) + 4 * slot) * 28 + row * 14 + p] = v + kk;
          if (MODE == 5) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                         "r"((unsigned)__cvta_generic_to_shared(st)), "r"(28 * pp), "r"(cg * 16), "r"(r)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      } else if (MODE == 6) {
        float* st = stage_all + warp * TS;
        for (int pp = 0; pp < 7; ++pp) {
          for (int row = 0; row < 2; ++row)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (p < 14) st[((kk & 3) + 8 * (kk >> 2) + 4 * slot) * 28 + row * 14 + p] = v + kk;
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (lane < 28)
              *reinterpret_cast<float4*>(base + (size_t)(4 * q + lane / 7) * PER + pp * 28 + 4 * (lane % 7)) =
                  *reinterpret_cast<const float4*>(st + q * 112 + 4 * lane);
          __syncwarp();
        }
      } else if (MODE == 8) {
#pragma unroll 1
        for (int c = 0; c < 16; ++c)
          for (int e = lane; e < 49; e += 32)
            *reinterpret_cast<float4*>(base + (size_t)c * PER + 4 * e) = make_float4(v, v + 1, v + 2, v + c);
      } else if (MODE == 9 || MODE == 10 || MODE == 11) {
        float* st = stage_all + warp * 1568;
        for (int half = 0; half < 2; ++half) {
          if (MODE == 10 || MODE == 11) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          for (int ph = 0; ph < 14; ++ph)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              if (p < 14) st[(kk + 4 * slot) * 196 + ph * 14 + p] = v + kk;
          if (MODE == 9) {
            __syncwarp();
            for (int e = lane; e < 392; e += 32)
              *reinterpret_cast<float4*>(base + (size_t)half * 8 * PER + 4 * e) =
                  *reinterpret_cast<const float4*>(st + 4 * e);
            __syncwarp();
          } else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                               base + (size_t)half * 8 * PER),
                           "r"((unsigned)__cvta_generic_to_shared(st)), "r"(6272)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
        if (MODE != 9) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
      } else if (MODE == 7) {
        float* st = stage_all + warp * TS;
        for (int pq = 0; pq < 4; ++pq) {  // 4 rows per tile (last tile: 2 rows)
          const int rows = pq < 3 ? 4 : 2;
          for (int row = 0; row < rows; ++row)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (p < 14) st[((kk & 3) + 8 * (kk >> 2) + 4 * slot) * 56 + row * 14 + p] = v + kk;
          __syncwarp();
          const int per = rows * 14 / 4;  // float4 per channel: 14 or 7
          for (int e = lane; e < 16 * per; e += 32) {
            const int c = e / per, x = e - c * per;
            *reinterpret_cast<float4*>(base + (size_t)c * PER + pq * 56 + 4 * x) =
                *reinterpret_cast<const float4*>(st + c * 56 + 4 * x);
          }
          __syncwarp();
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
void run(float* out, const CUtensorMap& map, int R, int* counter) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaMemset(counter, 0, 4));
    cudaEventRecord(a);
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
    k<MODE><<<148, 512, 104 * 1024>>>(out, map, R, counter);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it && ms < best) best = ms;
  }
  const double bytes = (double)R * C * PER * 4;
  printf("mode %d: %.3f ms  %.2f TB/s\n", MODE, best, bytes / best / 1e9);
}

int main() {
  const int R = 8192;
  float* out;
  int* counter;
  CK(cudaMalloc(&out, (size_t)R * C * PER * 4));
  CK(cudaMalloc(&counter, 4));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  CUtensorMap map;
  cuuint64_t dims[3] = {196, (cuuint64_t)C, (cuuint64_t)R};
  cuuint64_t strides[2] = {196 * 4, (cuuint64_t)C * 196 * 4};
  cuuint32_t box[3] = {28, 16, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  run<0>(out, map, R, counter);
  run<1>(out, map, R, counter);
  run<2>(out, map, R, counter);
  run<3>(out, map, R, counter);
  run<4>(out, map, R, counter);
  run<5>(out, map, R, counter);
  run<6>(out, map, R, counter);
  run<7>(out, map, R, counter);
  run<8>(out, map, R, counter);
  run<9>(out, map, R, counter);
  run<10>(out, map, R, counter);
  return 0;
}

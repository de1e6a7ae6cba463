// Caption-consistency alignment loss (symmetric InfoNCE without temperature) for sm_100a.
//
// Replaces detectron2/modeling/meta_arch/rcnn.py:305-317 (image level) and :455-468 (region level):
//   gather -> x / |x| -> S = A B^T -> (CE(S, arange) + CE(S^T, arange)) / 2,
// with the gradient semantics of backbone/clipcap/gather.py:16-20 (each rank keeps only the gradient of
// its own rows; no reduction).  Normalisation is row-local, so rows are normalised BEFORE the exchange and
// src|tgt are packed into one message: one all-gather per branch instead of the reference's two.
//
// Layout of the gathered buffer: packed_all[world][2][n_local][D], rank-major (what all_gather produces
// from each rank's packed[2][n_local][D]).  Row i of A (global index) lives at rank i / n_local, slot 0;
// row j of B at slot 1.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

// one warp per row; rows [0,n_local) are src, [n_local, 2 n_local) are tgt
__global__ void align_pack_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int n_local, int D,
                                  float* __restrict__ packed, float* __restrict__ norms) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= 2 * n_local) return;
  const float* s = row < n_local ? src + (size_t)row * D : tgt + (size_t)(row - n_local) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) ss = fmaf(s[d], s[d], ss);
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);  // no eps: rcnn.py:308-309 divides by the plain norm
  for (int d = lane; d < D; d += 32) packed[(size_t)row * D + d] = s[d] / nrm;
  if (lane == 0) norms[row] = nrm;
}

__device__ __forceinline__ const float* packed_row(const float* packed_all, int n_local, int D, int which, int i) {
  const int rk = i / n_local, li = i - rk * n_local;
  return packed_all + ((size_t)(rk * 2 + which) * n_local + li) * D;
}

// S[i][j] = A^_i . B^_j ; 32x32 tile per CTA, 256 threads (2x2 micro-tile), K-chunks of 32
__global__ void __launch_bounds__(256)
align_logits_kernel(const float* __restrict__ packed_all, int n, int n_local, int D, float* __restrict__ S) {
  __shared__ float As[32][33], Bs[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < D; k0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int rr = e >> 5, kk = e & 31;
      const int k = k0 + kk;
      As[rr][kk] = (i0 + rr < n && k < D) ? packed_row(packed_all, n_local, D, 0, i0 + rr)[k] : 0.f;
      Bs[rr][kk] = (j0 + rr < n && k < D) ? packed_row(packed_all, n_local, D, 1, j0 + rr)[k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float a0 = As[ty][kk], a1 = As[ty + 16][kk];
      const float b0 = Bs[tx][kk], b1 = Bs[tx + 16][kk];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < n && j < n) S[(size_t)i * n + j] = acc[a][b];
    }
}

// warp w < n: row log-sum-exp of S;  warp n + w: column log-sum-exp
__global__ void align_lse_kernel(const float* __restrict__ S, int n, float* __restrict__ rowlse,
                                 float* __restrict__ collse) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * n) return;
  const bool col = w >= n;
  const int i = col ? w - n : w;
  const size_t stride = col ? (size_t)n : 1, base = col ? (size_t)i : (size_t)i * n;
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, S[base + j * stride]);
  m = warp_max(m);
  float se = 0.f;
  for (int j = lane; j < n; j += 32) se += expf(S[base + j * stride] - m);
  se = warp_sum(se);
  if (lane == 0) (col ? collse : rowlse)[i] = logf(se) + m;
}

__global__ void align_loss_kernel(const float* __restrict__ S, int n, const float* __restrict__ rowlse,
                                  const float* __restrict__ collse, float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = S[(size_t)i * n + i];
    s += (rowlse[i] - d) + (collse[i] - d);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *loss = t / (2.f * (float)n);
  }
}

// CTA b < n_local: gradient of local A row;  b >= n_local: local B row.
//   dS_ij = (softmax_row(S)_ij + softmax_col(S)_ij - 2 [i==j]) / (2n)
//   dA^_i = sum_j dS_ij B^_j ,  dB^_j = sum_i dS_ij A^_i ,  d(x) = (dx^ - x^ (x^ . dx^)) / |x|
__global__ void __launch_bounds__(256)
align_grad_kernel(const float* __restrict__ packed_all, const float* __restrict__ norms_local,
                  const float* __restrict__ S, const float* __restrict__ rowlse, const float* __restrict__ collse,
                  int n, int n_local, int D, int row0, const float* __restrict__ grad_scale, float* __restrict__ da,
                  float* __restrict__ db) {
  extern __shared__ float coef[];  // [n]
  __shared__ float red[8];
  const bool bside = (int)blockIdx.x >= n_local;
  const int li = bside ? blockIdx.x - n_local : blockIdx.x;
  float* out = bside ? db : da;
  if (!out) return;
  const int gi = row0 + li;
  const float scale = (grad_scale ? *grad_scale : 1.f) * 0.5f / (float)n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    // a-side walks row gi of S, b-side walks column gi
    const float s = bside ? S[(size_t)j * n + gi] : S[(size_t)gi * n + j];
    const float pr = expf(s - (bside ? rowlse[j] : rowlse[gi]));
    const float pc = expf(s - (bside ? collse[gi] : collse[j]));
    coef[j] = scale * (pr + pc - (j == gi ? 2.f : 0.f));
  }
  __syncthreads();
  const float* self = packed_row(packed_all, n_local, D, bside ? 1 : 0, gi);
  const float nrm = norms_local[(bside ? n_local : 0) + li];
  // D is processed in slabs of blockDim.x columns; the projection needs the full dot first
  float dot_part = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(coef[j], packed_row(packed_all, n_local, D, bside ? 0 : 1, j)[d], acc);
    out[(size_t)li * D + d] = acc;  // raw dx^ parked in the output, fixed up below
    dot_part = fmaf(acc, self[d], dot_part);
  }
  dot_part = warp_sum(dot_part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot_part;
  __syncthreads();
  float dot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) dot += red[i];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g = out[(size_t)li * D + d];
    out[(size_t)li * D + d] = (g - self[d] * dot) / nrm;
  }
}

struct AlignWs {
  float* S;
  float* rowlse;
  float* collse;
  size_t total;
};
static AlignWs align_carve(void* base, int n) {
  AlignWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t m = (size_t)(n > 0 ? n : 1);
  w.S = (float*)take(m * m * 4);
  w.rowlse = (float*)take(m * 4);
  w.collse = (float*)take(m * 4);
  w.total = off;
  return w;
}

}  // namespace cddmsl

using namespace cddmsl;

extern "C" int cddmsl_align_pack_normalized(const float* src, const float* tgt, int n_local, int D, float* packed,
                                            float* norms, cddmsl_stream_t stream) {
  if (n_local < 0 || D <= 0) return CDDMSL_EINVAL;
  if (n_local == 0) return CDDMSL_OK;
  if (!src || !tgt || !packed || !norms) return CDDMSL_EINVAL;
  align_pack_kernel<<<ceil_div(2 * n_local * 32, 256), 256, 0, (cudaStream_t)stream>>>(src, tgt, n_local, D, packed,
                                                                                       norms);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

extern "C" size_t cddmsl_align_loss_workspace_bytes(int world, int n_local, int D) {
  (void)D;
  return align_carve(nullptr, world * n_local).total;
}

extern "C" int cddmsl_align_loss(const float* packed_all, const float* norms_local, int world, int n_local, int D,
                                 int rank, const float* grad_scale, float* loss, float* da, float* db,
                                 void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (world <= 0 || n_local < 0 || D <= 0 || rank < 0 || rank >= world || !loss) return CDDMSL_EINVAL;
  const int n = world * n_local;
  if (n == 0) {  // F.cross_entropy on an empty batch is NaN in the reference
    const float nanv = __builtin_nanf("");
    CDDMSL_CUDA(cudaMemcpyAsync(loss, &nanv, sizeof(float), cudaMemcpyHostToDevice, stream));
    return CDDMSL_OK;
  }
  if (!packed_all || !workspace || ((da || db) && !norms_local)) return CDDMSL_EINVAL;
  if (n > 16384) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return CDDMSL_EALIGN;
  AlignWs w = align_carve(workspace, n);
  if (w.total > workspace_bytes) return CDDMSL_EWORKSPACE;
  const int tiles = ceil_div(n, 32);
  align_logits_kernel<<<dim3(tiles, tiles), 256, 0, stream>>>(packed_all, n, n_local, D, w.S);
  align_lse_kernel<<<ceil_div(2 * n * 32, 256), 256, 0, stream>>>(w.S, n, w.rowlse, w.collse);
  align_loss_kernel<<<1, 256, 0, stream>>>(w.S, n, w.rowlse, w.collse, loss);
  count_launch(3);
  if (da || db) {
    const int smem = n * 4;
    if (smem > 48 * 1024)
      CDDMSL_CUDA(cudaFuncSetAttribute(align_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    align_grad_kernel<<<2 * n_local, 256, smem, stream>>>(packed_all, norms_local, w.S, w.rowlse, w.collse, n,
                                                          n_local, D, rank * n_local, grad_scale, da, db);
    count_launch();
  }
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

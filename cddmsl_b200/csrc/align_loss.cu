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

// S[i][j] = A^_i . B^_j ; 64x64 tile per CTA, 256 threads (4x4 micro-tile), K-chunks of 16 staged k-major so
// the inner product reads its four rows / four columns with one LDS.128 each.
constexpr int kLT = 64, kLK = 16;
__global__ void __launch_bounds__(256)
align_logits_kernel(const float* __restrict__ packed_all, int n, int n_local, int D, float logit_scale,
                    float* __restrict__ S) {
  __shared__ __align__(16) float As[kLK][kLT + 4], Bs[kLK][kLT + 4];
  const int i0 = blockIdx.y * kLT, j0 = blockIdx.x * kLT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: one row, four consecutive k
  const float* arow = i0 + lrow < n ? packed_row(packed_all, n_local, D, 0, i0 + lrow) : nullptr;
  const float* brow = j0 + lrow < n ? packed_row(packed_all, n_local, D, 1, j0 + lrow) : nullptr;
  const bool vec = (D & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float av[4], bv[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) av[c] = bv[c] = 0.f;
    const int k = k0 + lk;
    if (vec) {
      if (k < D) {
        if (arow) {
          const float4 t = *reinterpret_cast<const float4*>(arow + k);
          av[0] = t.x, av[1] = t.y, av[2] = t.z, av[3] = t.w;
        }
        if (brow) {
          const float4 t = *reinterpret_cast<const float4*>(brow + k);
          bv[0] = t.x, bv[1] = t.y, bv[2] = t.z, bv[3] = t.w;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (arow && k + c < D) av[c] = arow[k + c];
        if (brow && k + c < D) bv[c] = brow[k + c];
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < D; k0 += kLK) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      As[lk + c][lrow] = av[c];
      Bs[lk + c][lrow] = bv[c];
    }
    __syncthreads();
    if (k0 + kLK < D) fetch(k0 + kLK);  // next chunk's global loads fly under this chunk's FMAs
#pragma unroll
    for (int kk = 0; kk < kLK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = i0 + ty * 4 + x, j = j0 + tx * 4;
    if (i >= n) continue;
    if ((n & 3) == 0 && j + 3 < n) {
      *reinterpret_cast<float4*>(S + (size_t)i * n + j) =
          make_float4(acc[x][0] * logit_scale, acc[x][1] * logit_scale, acc[x][2] * logit_scale,
                      acc[x][3] * logit_scale);
    } else {
#pragma unroll
      for (int y = 0; y < 4; ++y)
        if (j + y < n) S[(size_t)i * n + j + y] = acc[x][y] * logit_scale;
    }
  }
}

// blocks [0, row_blocks): one warp per row of S (coalesced along the row);
// blocks [row_blocks, ..): one CTA per 32 columns -- lane = column, the 32 warps stride over the rows (eight loads in
// flight each) with an online (max, sum-exp) pair and merge through shared memory: the column pass is coalesced too.
constexpr int kLseWarps = 32;
__global__ void __launch_bounds__(kLseWarps * 32)
align_lse_kernel(const float* __restrict__ S, int n, int row_blocks, float* __restrict__ rowlse,
                 float* __restrict__ collse) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < row_blocks) {
    const int i = blockIdx.x * kLseWarps + warp;
    if (i >= n) return;
    const float* row = S + (size_t)i * n;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float se = 0.f;
    for (int j = lane; j < n; j += 32) se += expf(row[j] - m);
    se = warp_sum(se);
    if (lane == 0) rowlse[i] = logf(se) + m;
    return;
  }
  __shared__ float sm[kLseWarps][32], ss[kLseWarps][32];
  const int j = (blockIdx.x - row_blocks) * 32 + lane;
  float m = -INFINITY, se = 0.f;
  if (j < n) {
    for (int i = warp; i < n; i += kLseWarps * 8) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int ii = i + u * kLseWarps;
        x[u] = ii < n ? S[(size_t)ii * n + j] : -INFINITY;
      }
      float cm = x[0];
#pragma unroll
      for (int u = 1; u < 8; ++u) cm = fmaxf(cm, x[u]);
      if (cm > m) {
        se *= expf(m - cm);  // expf(-inf) == 0 covers the first batch
        m = cm;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) se += expf(x[u] - m);
    }
  }
  sm[warp][lane] = m;
  ss[warp][lane] = se;
  __syncthreads();
  if (warp == 0 && j < n) {
    float mm = sm[0][lane];
    for (int w = 1; w < kLseWarps; ++w) mm = fmaxf(mm, sm[w][lane]);
    float tot = 0.f;
    for (int w = 0; w < kLseWarps; ++w) tot += ss[w][lane] == 0.f ? 0.f : ss[w][lane] * expf(sm[w][lane] - mm);
    collse[j] = logf(tot) + mm;
  }
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

// Gradient of this rank's rows.  CTA (x, y): x < gb owns kGR consecutive local src rows, x >= gb tgt rows; y picks a
// 64-column slab of D.  Every row of the other side streamed from L2 feeds kGR accumulators; the 256 threads are
// 64 columns x 4 interleaved j-partitions (eight loads in flight per thread), merged through shared memory.
//   dS_ij = (softmax_row(S)_ij + softmax_col(S)_ij - 2 [i==j]) / (2n)
//   dA^_i = sum_j dS_ij B^_j ,  dB^_j = sum_i dS_ij A^_i
// The coefficients of a chunk of kGJ opposite rows are staged in shared memory as coef[j][r] (two LDS.128 per j).
// Writes the raw dx^; align_grad_finish_kernel applies d(x) = (dx^ - x^ (x^ . dx^)) / |x|.
constexpr int kGR = 8, kGJ = 1024, kGD = 64, kGP = 4;
__global__ void __launch_bounds__(256)
align_grad_kernel(const float* __restrict__ packed_all, const float* __restrict__ S,
                  const float* __restrict__ rowlse, const float* __restrict__ collse, int n, int n_local, int D,
                  int row0, const float* __restrict__ grad_scale, float grad_mult, float* __restrict__ da,
                  float* __restrict__ db) {
  __shared__ __align__(16) float coef[kGJ][kGR];
  __shared__ unsigned rowoff[kGJ];  // float offset of opposite row j inside packed_all (no division in the hot loop)
  const int gb = gridDim.x >> 1;
  const bool bside = (int)blockIdx.x >= gb;
  const int l0 = (bside ? blockIdx.x - gb : blockIdx.x) * kGR;
  float* out = bside ? db : da;
  if (!out) return;
  const int rows = min(kGR, n_local - l0);
  const int g0 = row0 + l0;
  const float scale = (grad_scale ? *grad_scale : 1.f) * grad_mult * 0.5f / (float)n;
  const int tid = threadIdx.x, dl = tid & (kGD - 1), jp = tid / kGD;
  const int d = blockIdx.y * kGD + dl;
  const bool dok = d < D;
  const int other = bside ? 0 : 1;

  float acc[kGR];
#pragma unroll
  for (int r = 0; r < kGR; ++r) acc[r] = 0.f;
  for (int jc = 0; jc < n; jc += kGJ) {
    const int jn = min(kGJ, n - jc);
    __syncthreads();
    if (bside) {
      // column block S[j][g0 .. g0+7]: eight lanes share one 32-byte sector
      const int r = tid & 7;
      const float cl = r < rows ? collse[g0 + r] : 0.f;
      for (int jb = tid >> 3; jb < jn; jb += 32 * 4) {
        float sv[4], rl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = jc + min(jb + u * 32, jn - 1);
          sv[u] = r < rows ? S[(size_t)j * n + g0 + r] : 0.f;
          rl[u] = rowlse[j];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = jb + u * 32, j = jc + jj;
          if (jj < jn)
            coef[jj][r] = r < rows ? scale * (expf(sv[u] - rl[u]) + expf(sv[u] - cl) - (j == g0 + r ? 2.f : 0.f)) : 0.f;
        }
      }
    } else {
      for (int jj = tid; jj < jn; jj += 256) {
        const int j = jc + jj;
        const float cl = collse[j];
#pragma unroll
        for (int r = 0; r < kGR; ++r) {
          float c = 0.f;
          if (r < rows) {
            const float s = S[(size_t)(g0 + r) * n + j];
            c = scale * (expf(s - rowlse[g0 + r]) + expf(s - cl) - (j == g0 + r ? 2.f : 0.f));
          }
          coef[jj][r] = c;
        }
      }
    }
    for (int jj = tid; jj < jn; jj += 256)
      rowoff[jj] = (unsigned)(packed_row(packed_all, n_local, D, other, jc + jj) - packed_all);
    __syncthreads();
    if (dok) {
      const float* col = packed_all + d;
      for (int jj = jp; jj < jn; jj += kGP * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(col + rowoff[min(jj + u * kGP, jn - 1)]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j2 = jj + u * kGP;
          const float x = j2 < jn ? v[u] : 0.f;
          const int jq = min(j2, jn - 1);
          const float4 c0 = *reinterpret_cast<const float4*>(&coef[jq][0]);
          const float4 c1 = *reinterpret_cast<const float4*>(&coef[jq][4]);
          acc[0] = fmaf(c0.x, x, acc[0]);
          acc[1] = fmaf(c0.y, x, acc[1]);
          acc[2] = fmaf(c0.z, x, acc[2]);
          acc[3] = fmaf(c0.w, x, acc[3]);
          acc[4] = fmaf(c1.x, x, acc[4]);
          acc[5] = fmaf(c1.y, x, acc[5]);
          acc[6] = fmaf(c1.z, x, acc[6]);
          acc[7] = fmaf(c1.w, x, acc[7]);
        }
      }
    }
  }
  // merge the four j-partitions (coef is free again after the barrier)
  __syncthreads();
  float* red = &coef[0][0];  // [kGP][kGR][kGD]
#pragma unroll
  for (int r = 0; r < kGR; ++r) red[(jp * kGR + r) * kGD + dl] = acc[r];
  __syncthreads();
  for (int e = tid; e < kGR * kGD; e += 256) {
    const int r = e / kGD, c = e % kGD;
    const int dd = blockIdx.y * kGD + c;
    if (r < rows && dd < D) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < kGP; ++q) t += red[(q * kGR + r) * kGD + c];
      out[(size_t)(l0 + r) * D + dd] = t;
    }
  }
}

// one warp per local row (rows [0,n_local): src, then tgt): projection through the normalisation, in place
__global__ void align_grad_finish_kernel(const float* __restrict__ packed_all, const float* __restrict__ norms_local,
                                         int n_local, int D, int row0, float* __restrict__ da,
                                         float* __restrict__ db) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 2 * n_local) return;
  const bool bside = w >= n_local;
  const int li = bside ? w - n_local : w;
  float* out = bside ? db : da;
  if (!out) return;
  out += (size_t)li * D;
  const float* self = packed_row(packed_all, n_local, D, bside ? 1 : 0, row0 + li);
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(out[d], self[d], dot);
  dot = warp_sum(dot);
  const float nrm = norms_local[w];
  for (int d = lane; d < D; d += 32) out[d] = (out[d] - self[d] * dot) / nrm;
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

extern "C" int cddmsl_contrastive_loss(const float* packed_all, const float* norms_local, int world, int n_local,
                                       int D, int rank, float logit_scale, float grad_mult, const float* grad_scale,
                                       float* loss, float* da, float* db, void* workspace, size_t workspace_bytes,
                                       cddmsl_stream_t stream_);

extern "C" int cddmsl_align_loss(const float* packed_all, const float* norms_local, int world, int n_local, int D,
                                 int rank, const float* grad_scale, float* loss, float* da, float* db,
                                 void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  return cddmsl_contrastive_loss(packed_all, norms_local, world, n_local, D, rank, 1.f, 1.f, grad_scale, loss, da, db,
                                 workspace, workspace_bytes, stream_);
}

extern "C" int cddmsl_contrastive_loss(const float* packed_all, const float* norms_local, int world, int n_local,
                                       int D, int rank, float logit_scale, float grad_mult, const float* grad_scale,
                                       float* loss, float* da, float* db, void* workspace, size_t workspace_bytes,
                                       cddmsl_stream_t stream_) {
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
  const int tiles = ceil_div(n, kLT);
  const int row_blocks = ceil_div(n, kLseWarps);
  align_logits_kernel<<<dim3(tiles, tiles), 256, 0, stream>>>(packed_all, n, n_local, D, logit_scale, w.S);
  align_lse_kernel<<<row_blocks + ceil_div(n, 32), kLseWarps * 32, 0, stream>>>(w.S, n, row_blocks, w.rowlse,
                                                                               w.collse);
  align_loss_kernel<<<1, 256, 0, stream>>>(w.S, n, w.rowlse, w.collse, loss);
  count_launch(3);
  if ((da || db) && n_local > 0) {
    align_grad_kernel<<<dim3(2 * ceil_div(n_local, kGR), ceil_div(D, kGD)), 256, 0, stream>>>(
        packed_all, w.S, w.rowlse, w.collse, n, n_local, D, rank * n_local, grad_scale, grad_mult * logit_scale, da,
        db);
    align_grad_finish_kernel<<<ceil_div(2 * n_local * 32, 256), 256, 0, stream>>>(packed_all, norms_local, n_local, D,
                                                                                  rank * n_local, da, db);
    count_launch(2);
  }
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

// ------------------------------------------------------------------------------------------------
// KD regulariser of the image-level branch (detectron2/modeling/meta_arch/rcnn.py:265-272):
// kd = L1Loss()(teacher.detach(), student) = mean |teacher - student| over all elements; the gradient
// d kd / d student = sign(student - teacher) / numel comes out of the same pass.  Deterministic: fixed-order
// per-block partial sums, summed in order by the last block.
// ------------------------------------------------------------------------------------------------
namespace cddmsl {
constexpr int kKdThreads = 256, kKdMaxBlocks = 1024;

__global__ void __launch_bounds__(kKdThreads) kd_l1_kernel(const float* __restrict__ teacher,
                                                          const float* __restrict__ student, long long n, float inv_n,
                                                          const float* __restrict__ grad_scale, float* __restrict__ loss,
                                                          float* __restrict__ dstudent, float* __restrict__ partial,
                                                          unsigned int* __restrict__ ticket) {
  __shared__ float red[kKdThreads / 32];
  __shared__ bool last;
  const float gs = grad_scale ? __ldg(grad_scale) * inv_n : inv_n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * kKdThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKdThreads) {
    const float d = __ldg(student + i) - __ldg(teacher + i);
    acc += fabsf(d);
    // torch's L1 backward: sign(input - target) * grad / numel, sign(0) = 0
    if (dstudent) dstudent[i] = d > 0.f ? gs : d < 0.f ? -gs : 0.f;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kKdThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) t += *((volatile float*)partial + b);
    *loss = t * inv_n;
    *ticket = 0u;  // workspace is reusable without another memset
  }
}
}  // namespace cddmsl

extern "C" size_t cddmsl_kd_l1_loss_workspace_bytes(void) { return (cddmsl::kKdMaxBlocks + 64) * sizeof(float); }

extern "C" int cddmsl_kd_l1_loss(const float* teacher, const float* student, int64_t numel, const float* grad_scale,
                                 float* loss, float* dstudent, void* workspace, size_t workspace_bytes,
                                 cddmsl_stream_t stream_) {
  using namespace cddmsl;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (numel < 0 || !loss) return CDDMSL_EINVAL;
  if (numel == 0) {  // mean over nothing: NaN like torch's L1Loss on empty inputs
    const float nanv = __builtin_nanf("");
    CDDMSL_CUDA(cudaMemcpyAsync(loss, &nanv, sizeof(float), cudaMemcpyHostToDevice, stream));
    return CDDMSL_OK;
  }
  if (!teacher || !student || !workspace) return CDDMSL_EINVAL;
  if (workspace_bytes < cddmsl_kd_l1_loss_workspace_bytes()) return CDDMSL_EWORKSPACE;
  float* partial = (float*)workspace;
  unsigned int* ticket = (unsigned int*)(partial + kKdMaxBlocks);
  CDDMSL_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream));
  long long blocks = (numel + kKdThreads - 1) / kKdThreads;
  if (blocks > kKdMaxBlocks) blocks = kKdMaxBlocks;
  kd_l1_kernel<<<(unsigned)blocks, kKdThreads, 0, stream>>>(teacher, student, (long long)numel, 1.0f / (float)numel,
                                                            grad_scale, loss, dstudent, partial, ticket);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row-softmax losses against a DENSE target matrix (RegionCLIP pretraining):
//   mode 0  KL distillation   detectron2/modeling/meta_arch/clip_rcnn.py:597-600:
//           F.kl_div(softmax(s).log(), t, 'batchmean') = sum_ik t_ik (log t_ik - log p_ik) / R   (0 log 0 = 0)
//   mode 1  MIL cross-entropy detectron2/utils/comm.py:332-355 (avg_positives=False, weights=None):
//           mean_i -log(sum_k t_ik p_ik)
// logits [R, ld] (the first K columns count: the cosine-logit op appends a background column), target [R, K].
// One warp per row: online max / sum-exp pass, then loss and d loss / d logits (columns >= K get 0).  The mean is a
// fixed-order two-stage sum.
// ------------------------------------------------------------------------------------------------
namespace cddmsl {
constexpr int kTgtWarps = 8;

__global__ void __launch_bounds__(kTgtWarps * 32)
softmax_target_rows_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ target, int R, int K,
                           int mode, const float* __restrict__ grad_scale, float* __restrict__ row_loss,
                           float* __restrict__ dlogits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kTgtWarps + warp;
  if (i >= R) return;
  const float* s = logits + (size_t)i * ld;
  const float* t = target + (size_t)i * K;
  float m = -INFINITY;
  for (int k = lane; k < K; k += 32) m = fmaxf(m, s[k]);
  m = warp_max(m);
  float z = 0.f, st = 0.f, stp = 0.f, stl = 0.f, sts = 0.f;  // sum exp, sum t, sum t*exp, sum t log t, sum t*s
  for (int k = lane; k < K; k += 32) {
    const float e = expf(s[k] - m), tk = t[k];
    z += e;
    st += tk;
    stp += tk * e;
    if (tk > 0.f) stl += tk * logf(tk);
    sts += tk * (s[k] - m);
  }
  z = warp_sum(z);
  st = warp_sum(st);
  stp = warp_sum(stp);
  stl = warp_sum(stl);
  sts = warp_sum(sts);
  const float logz = logf(z);
  // KL: sum t (log t - (s - m - log z)) ;  MIL: -log(sum t p) = log z - log(sum t e)
  const float loss = mode == 0 ? (stl - sts + st * logz) : (logz - logf(stp));
  if (lane == 0) row_loss[i] = loss;
  if (dlogits) {
    const float gs = (grad_scale ? __ldg(grad_scale) : 1.f) / (float)R;
    float* d = dlogits + (size_t)i * ld;
    const float inv_z = 1.f / z, inv_stp = 1.f / stp;
    for (int k = lane; k < ld; k += 32) {
      float g = 0.f;
      if (k < K) {
        const float p = expf(s[k] - m) * inv_z;
        g = mode == 0 ? (p * st - t[k]) : (p - t[k] * expf(s[k] - m) * inv_stp);
      }
      d[k] = g * gs;
    }
  }
}

__global__ void __launch_bounds__(256) mean_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += v[i];  // fixed order per thread, fixed tree below
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    *out = t / (float)n;
  }
}
}  // namespace cddmsl

extern "C" size_t cddmsl_softmax_target_loss_workspace_bytes(int R) { return ((size_t)(R > 0 ? R : 0) + 64) * 4; }

extern "C" int cddmsl_softmax_target_loss(const float* logits, int ld, const float* target, int R, int K, int mode,
                                          const float* grad_scale, float* loss, float* dlogits, void* workspace,
                                          size_t workspace_bytes, cddmsl_stream_t stream_) {
  using namespace cddmsl;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (R < 0 || K <= 0 || ld < K || !loss || (mode != 0 && mode != 1)) return CDDMSL_EINVAL;
  if (R == 0) {  // mean over nothing
    const float nanv = __builtin_nanf("");
    CDDMSL_CUDA(cudaMemcpyAsync(loss, &nanv, sizeof(float), cudaMemcpyHostToDevice, stream));
    return CDDMSL_OK;
  }
  if (!logits || !target || !workspace) return CDDMSL_EINVAL;
  if (workspace_bytes < cddmsl_softmax_target_loss_workspace_bytes(R)) return CDDMSL_EWORKSPACE;
  float* row_loss = (float*)workspace;
  softmax_target_rows_kernel<<<ceil_div(R, kTgtWarps), kTgtWarps * 32, 0, stream>>>(logits, ld, target, R, K, mode,
                                                                                   grad_scale, row_loss, dlogits);
  mean_rows_kernel<<<1, 256, 0, stream>>>(row_loss, R, loss);
  count_launch(2);
  return (int)cudaGetLastError();
}

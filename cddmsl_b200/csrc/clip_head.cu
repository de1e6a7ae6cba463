// CLIP box-predictor head for sm_100a: cosine logits against concept text embeddings + background
// embedding, temperature, softmax, focal / (weighted) cross-entropy, gradient w.r.t. the region embeddings
// and the classification statistics — one fused warp-per-row kernel for the VOC/Cityscapes-size vocabularies.
//
// Replaces detectron2/modeling/roi_heads/fast_rcnn.py:543-565 (F.normalize, two matmuls, cat, /T),
// :624-644 (focal_loss), :611-615 + layers/wrappers.py:26-33 (CE / weighted CE) and :100-127 (stats),
// which are ~20 ATen launches in the reference.
//
// Per row r (one warp): x is streamed once for the norm and the K+1 dot products (8 classes at a time so
// the row stays in L1), the logits live in shared memory, softmax/CE/focal and dL/dlogits are evaluated
// there, and a second sweep forms dx = (sum_k g_k w^_k - x^ sum_k g_k c_k) / (T |x|).
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

constexpr int kHeadThreads = 256;
constexpr int kHeadWarps = kHeadThreads / 32;
constexpr int kKT = 8;  // classes per sweep over the row
constexpr float kNormEps = 1e-12f;  // F.normalize eps (fast_rcnn.py:547)

enum { HEAD_OP_SCORES = 0, HEAD_OP_LOSS = 1, HEAD_OP_SCORES_BWD = 2 };

// rows 0..K-1: w_k / max(|w_k|, eps);  row K: the raw background embedding (an nn.Linear, fast_rcnn.py:560)
__global__ void clip_head_prep_kernel(const float* __restrict__ w, const float* __restrict__ w_bg, int K, int D,
                                      float* __restrict__ wall) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp > K) return;
  const float* src = warp < K ? w + (size_t)warp * D : w_bg;
  float ss = 0.f;
  if (warp < K) {
    for (int d = lane; d < D; d += 32) ss = fmaf(src[d], src[d], ss);
    ss = warp_sum(ss);
  }
  const float inv = warp < K ? 1.f / fmaxf(sqrtf(ss), kNormEps) : 1.f;
  for (int d = lane; d < D; d += 32) wall[(size_t)warp * D + d] = warp < K ? src[d] * inv : src[d];
}

// norm[0] = sum_i weight(gt_i) for the weighted-CE mean (F.cross_entropy(weight=...) semantics)
__global__ void clip_head_wsum_kernel(const int64_t* __restrict__ gt, int R, int K, float bg_weight,
                                      float* __restrict__ norm) {
  __shared__ int red[32];
  int nbg = 0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) nbg += (gt[i] == K);
  for (int o = 16; o > 0; o >>= 1) nbg += __shfl_xor_sync(0xffffffffu, nbg, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nbg;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *norm = (float)(R - t) + bg_weight * (float)t;
  }
}

struct HeadArgs {
  const float* x;
  const float* wall;       // [(K+1), D] prepared
  const int64_t* gt;
  const float* dscores;    // HEAD_OP_SCORES_BWD
  const float* grad_scale; // nullable
  const float* norm_dev;   // nullable: denominator on the device (weighted CE)
  float norm_host;
  int R, D, K;
  float inv_T;
  int op, loss_mode;
  float gamma, bg_weight;
  int strict_nan;
  float* scores;
  float* dx;
  float* partial;  // [gridDim.x] per-block loss sums
  int32_t* stats;
};

__global__ void __launch_bounds__(kHeadThreads) clip_head_kernel(HeadArgs a) {
  extern __shared__ float sm[];  // [kHeadWarps][K1p]
  const int K1 = a.K + 1;
  const int K1p = (K1 + 3) & ~3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* lg = sm + warp * K1p;
  const int D4 = a.D >> 2;
  const float4* wall4 = reinterpret_cast<const float4*>(a.wall);
  float loss_acc = 0.f;
  int st_acc = 0, st_fg = 0, st_fgacc = 0, st_fn = 0;
  const float gscale = a.grad_scale ? *a.grad_scale : 1.f;
  const float inv_norm_den = 1.f / (a.norm_dev ? *a.norm_dev : a.norm_host);

  for (int r = blockIdx.x * kHeadWarps + warp; r < a.R; r += gridDim.x * kHeadWarps) {
    const float4* x4 = reinterpret_cast<const float4*>(a.x + (size_t)r * a.D);
    // ---- sweep 1: |x|^2 and the K+1 dot products
    float ss = 0.f;
    for (int k0 = 0; k0 < K1; k0 += kKT) {
      float acc[kKT];
#pragma unroll
      for (int kk = 0; kk < kKT; ++kk) acc[kk] = 0.f;
      for (int v = lane; v < D4; v += 32) {
        const float4 xv = x4[v];
        if (k0 == 0) ss = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, ss))));
#pragma unroll
        for (int kk = 0; kk < kKT; ++kk) {
          if (k0 + kk < K1) {
            const float4 wv = __ldg(wall4 + (size_t)(k0 + kk) * D4 + v);
            acc[kk] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[kk]))));
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < kKT; ++kk) {
        const float s = warp_sum(acc[kk]);
        if (lane == 0 && k0 + kk < K1) lg[k0 + kk] = s;
      }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float inv_n = 1.f / fmaxf(nrm, kNormEps);
    __syncwarp();
    // logits l_k = (x^ . w_k) / T
    for (int k = lane; k < K1; k += 32) {
      const float l = lg[k] * inv_n * a.inv_T;
      lg[k] = l;
      if (a.scores) a.scores[(size_t)r * K1 + k] = l;
    }
    __syncwarp();
    if (a.op == HEAD_OP_SCORES) continue;

    float coef_s = 0.f;  // s = sum_k g_k c_k
    if (a.op == HEAD_OP_LOSS) {
      // ---- softmax / CE / focal on the K+1 logits
      float m = -INFINITY;
      int am = 0x7fffffff;
      for (int k = lane; k < K1; k += 32) {
        const float l = lg[k];
        if (l > m) {
          m = l;
          am = k;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // max with first-index tie break (torch.argmax on CPU)
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) {
          m = om;
          am = oa;
        }
      }
      float se = 0.f;
      for (int k = lane; k < K1; k += 32) se += expf(lg[k] - m);
      se = warp_sum(se);
      const int t = min(max((int)a.gt[r], 0), a.K);  // gt must lie in [0, K]; clamp guards shared memory
      const float lt = lg[t];
      const float ce = (logf(se) + m) - lt;
      const float et = expf(lt - m);
      const float pt = et / se;
      const bool is_bg = (t == a.K);
      float wi = 1.f, li, coef;
      if (a.loss_mode == CDDMSL_LOSS_FOCAL) {
        if (a.bg_weight >= 0.f && is_bg) wi = a.bg_weight;
        const float om = 1.f - pt;
        const float f = powf(om, a.gamma);
        li = ce * f;
        float t2;
        if (om > 0.f) t2 = a.gamma * ce * pt * (f / om);
        else t2 = a.strict_nan ? __int_as_float(0x7fc00000) : 0.f;
        coef = f + t2;
      } else {
        if (a.loss_mode == CDDMSL_LOSS_WEIGHTED_CE && is_bg) wi = a.bg_weight;
        li = ce;
        coef = 1.f;
      }
      if (lane == 0) {
        loss_acc += wi * li;
        st_acc += (am == t);
        if (t >= 0 && t < a.K) {
          st_fg += 1;
          st_fgacc += (am == t);
          st_fn += (am == a.K);
        }
      }
      if (!a.dx) continue;
      // g_k = coef * w_i / den * (p_k - [k == t]) * upstream
      const float cg = coef * wi * inv_norm_den * gscale;
      const float inv_se = 1.f / se;
      __syncwarp();
      for (int k = lane; k < K1; k += 32) {
        const float l = lg[k];
        const float pk = expf(l - m) * inv_se;
        const float gk = cg * (pk - (k == t ? 1.f : 0.f));
        coef_s = fmaf(gk, l, coef_s);  // l = c_k / T  ->  s = T * sum g_k l_k
        lg[k] = gk;
      }
    } else {  // HEAD_OP_SCORES_BWD: g_k given
      for (int k = lane; k < K1; k += 32) {
        const float gk = a.dscores[(size_t)r * K1 + k];
        coef_s = fmaf(gk, lg[k], coef_s);
        lg[k] = gk;
      }
    }
    coef_s = warp_sum(coef_s);  // sum_k g_k l_k = s / T
    __syncwarp();
    // ---- sweep 2: dx = (u - x^ s) / (T |x|),  u = sum_k g_k w^_k
    const bool clamped = nrm < kNormEps;  // x^ = x / eps: no gradient through the norm
    const float sx = clamped ? 0.f : coef_s * inv_n * inv_n;           // (s/T) / |x|^2 applied to x
    const float su = inv_n * a.inv_T;
    float4* dx4 = reinterpret_cast<float4*>(a.dx + (size_t)r * a.D);
    for (int v = lane; v < D4; v += 32) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < K1; ++k) {
        const float gk = lg[k];
        const float4 wv = __ldg(wall4 + (size_t)k * D4 + v);
        u.x = fmaf(gk, wv.x, u.x);
        u.y = fmaf(gk, wv.y, u.y);
        u.z = fmaf(gk, wv.z, u.z);
        u.w = fmaf(gk, wv.w, u.w);
      }
      const float4 xv = x4[v];
      float4 o;
      o.x = u.x * su - xv.x * sx;
      o.y = u.y * su - xv.y * sx;
      o.z = u.z * su - xv.z * sx;
      o.w = u.w * su - xv.w * sx;
      dx4[v] = o;
    }
    __syncwarp();
  }
  if (a.op != HEAD_OP_LOSS) return;
  // ---- block reduction of the loss (fixed order -> deterministic) and of the statistics
  __shared__ float red[kHeadWarps];
  if (lane == 0) red[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kHeadWarps; ++i) s += red[i];
    a.partial[blockIdx.x] = s;
  }
  if (a.stats && lane == 0) {
    if (st_acc) atomicAdd(a.stats + 0, st_acc);
    if (st_fg) atomicAdd(a.stats + 1, st_fg);
    if (st_fgacc) atomicAdd(a.stats + 2, st_fgacc);
    if (st_fn) atomicAdd(a.stats + 3, st_fn);
  }
}

__global__ void clip_head_finish_kernel(const float* __restrict__ partial, int n, const float* __restrict__ norm_dev,
                                        float norm_host, float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    *loss = t / (norm_dev ? *norm_dev : norm_host);
  }
}

struct HeadWs {
  float* wall;
  float* partial;
  float* norm;
  size_t total;
  int grid;
};

static HeadWs head_carve(void* base, int R, int D, int K) {
  HeadWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.grid = min(max(ceil_div(R, kHeadWarps), 1), sm_count() * 8);
  w.wall = (float*)take((size_t)(K + 1) * D * 4);
  w.partial = (float*)take((size_t)w.grid * 4);
  w.norm = (float*)take(4);
  w.total = off;
  return w;
}

bool clip_head_tc_eligible(int R, int D, int K);
size_t clip_head_tc_workspace_bytes(int R, int D, int K);
int clip_head_tc_run(int op, const float* x, const float* w, const float* w_bg, const int64_t* gt,
                     const float* dscores, int R, int D, int K, float temperature, int loss_mode, float gamma,
                     float bg_weight, const float* grad_scale, int strict_nan, float* scores, float* loss, float* dx,
                     int32_t* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream);

static int head_launch(HeadArgs a, const float* w, const float* w_bg, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, float* loss) {
  if (a.R < 0 || a.D <= 0 || a.K < 0 || (a.D & 3)) return CDDMSL_EINVAL;
  if (!w_bg || (a.K > 0 && !w) || !workspace) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return CDDMSL_EALIGN;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) || (reinterpret_cast<uintptr_t>(a.dx) & 15)) return CDDMSL_EALIGN;
  HeadWs ws = head_carve(workspace, a.R, a.D, a.K);
  if (ws.total > workspace_bytes) return CDDMSL_EWORKSPACE;
  const int K1p = (a.K + 1 + 3) & ~3;
  const int smem = kHeadWarps * K1p * 4;
  if (smem > 200 * 1024) return CDDMSL_EINVAL;
  a.wall = ws.wall;
  a.partial = ws.partial;
  if (a.op == HEAD_OP_LOSS) {
    if (a.stats) CDDMSL_CUDA(cudaMemsetAsync(a.stats, 0, 4 * sizeof(int32_t), stream));
    if (a.R == 0) {  // gradient-connected zero of wrappers.py:31-32
      CDDMSL_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
      return CDDMSL_OK;
    }
  }
  if (a.R == 0) return CDDMSL_OK;
  // LVIS-scale vocabularies: the logits are a dense contraction -> tcgen05 path (clip_head_tc.cuh)
  if (a.op != HEAD_OP_SCORES_BWD && clip_head_tc_eligible(a.R, a.D, a.K) &&
      workspace_bytes >= clip_head_tc_workspace_bytes(a.R, a.D, a.K))
    return clip_head_tc_run(a.op == HEAD_OP_SCORES ? 0 : 1, a.x, w, w_bg, a.gt, nullptr, a.R, a.D, a.K, 1.f / a.inv_T,
                            a.loss_mode, a.gamma, a.bg_weight, a.grad_scale, a.strict_nan, a.scores, loss, a.dx,
                            a.stats, workspace, workspace_bytes, stream);
  clip_head_prep_kernel<<<ceil_div((a.K + 1) * 32, 256), 256, 0, stream>>>(w, w_bg, a.K, a.D, ws.wall);
  count_launch();
  a.norm_host = (float)a.R;
  a.norm_dev = nullptr;
  if (a.op == HEAD_OP_LOSS && a.loss_mode == CDDMSL_LOSS_WEIGHTED_CE) {
    clip_head_wsum_kernel<<<1, 1024, 0, stream>>>(a.gt, a.R, a.K, a.bg_weight, ws.norm);
    count_launch();
    a.norm_dev = ws.norm;
  }
  if (smem > 48 * 1024)
    CDDMSL_CUDA(cudaFuncSetAttribute(clip_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  clip_head_kernel<<<ws.grid, kHeadThreads, smem, stream>>>(a);
  count_launch();
  if (a.op == HEAD_OP_LOSS) {
    clip_head_finish_kernel<<<1, 256, 0, stream>>>(ws.partial, ws.grid, a.norm_dev, a.norm_host, loss);
    count_launch();
  }
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

}  // namespace cddmsl

#include "clip_head_tc.cuh"

using namespace cddmsl;

extern "C" size_t cddmsl_clip_head_workspace_bytes(int R, int D, int K) {
  const int r = R > 0 ? R : 0;
  size_t b = head_carve(nullptr, r, D, K).total;
  if (r > 0 && clip_head_tc_eligible(r, D, K)) b = b > clip_head_tc_workspace_bytes(r, D, K) ? b : clip_head_tc_workspace_bytes(r, D, K);
  return b;
}

extern "C" int cddmsl_clip_head_scores(const float* x, const float* w, const float* w_bg, int R, int D, int K,
                                       float temperature, float* scores, void* workspace, size_t workspace_bytes,
                                       cddmsl_stream_t stream) {
  if (R > 0 && (!x || !scores)) return CDDMSL_EINVAL;
  HeadArgs a = {};
  a.x = x;
  a.R = R;
  a.D = D;
  a.K = K;
  a.inv_T = 1.f / temperature;
  a.op = HEAD_OP_SCORES;
  a.scores = scores;
  return head_launch(a, w, w_bg, workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}

extern "C" int cddmsl_clip_head_scores_bwd(const float* x, const float* w, const float* w_bg, const float* dscores,
                                           int R, int D, int K, float temperature, float* dx, void* workspace,
                                           size_t workspace_bytes, cddmsl_stream_t stream) {
  if (R > 0 && (!x || !dscores || !dx)) return CDDMSL_EINVAL;
  HeadArgs a = {};
  a.x = x;
  a.dscores = dscores;
  a.R = R;
  a.D = D;
  a.K = K;
  a.inv_T = 1.f / temperature;
  a.op = HEAD_OP_SCORES_BWD;
  a.dx = dx;
  return head_launch(a, w, w_bg, workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}

extern "C" int cddmsl_clip_head_loss(const float* x, const float* w, const float* w_bg, const int64_t* gt, int R,
                                     int D, int K, float temperature, int loss_mode, float gamma, float bg_weight,
                                     const float* grad_scale, int strict_nan, float* scores, float* loss, float* dx,
                                     int32_t* stats, void* workspace, size_t workspace_bytes,
                                     cddmsl_stream_t stream) {
  if (!loss) return CDDMSL_EINVAL;
  if (R > 0 && (!x || !gt)) return CDDMSL_EINVAL;
  if (loss_mode < CDDMSL_LOSS_FOCAL || loss_mode > CDDMSL_LOSS_WEIGHTED_CE) return CDDMSL_EINVAL;
  if (loss_mode == CDDMSL_LOSS_WEIGHTED_CE && bg_weight < 0.f) return CDDMSL_EINVAL;
  HeadArgs a = {};
  a.x = x;
  a.gt = gt;
  a.grad_scale = grad_scale;
  a.R = R;
  a.D = D;
  a.K = K;
  a.inv_T = 1.f / temperature;
  a.op = HEAD_OP_LOSS;
  a.loss_mode = loss_mode;
  a.gamma = gamma;
  a.bg_weight = bg_weight;
  a.strict_nan = strict_nan;
  a.scores = scores;
  a.dx = dx;
  a.stats = stats;
  return head_launch(a, w, w_bg, workspace, workspace_bytes, (cudaStream_t)stream, loss);
}

// Box-regression loss of the box predictor, fused (fast_rcnn.py:646-689 + box_regression.py:42-75 +
// fvcore.nn.smooth_l1_loss): sum over foreground rows of smooth-L1(pred_deltas, get_deltas(proposal, gt)) / R.
//
// The reference evaluates this with ~50 tiny PyTorch kernels and two device->host syncs (nonzero() for the
// foreground selection, the validity assert of get_deltas).  Here: one pass, one thread per row, per-CTA partial
// sums reduced in a fixed order by a second tiny kernel (deterministic), gradient written in the same pass.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

struct BoxRegArgs {
  const float* prop;   // [R,4] xyxy
  const float* gtb;    // [R,4]
  const float* pred;   // [R,4] (class agnostic) or [R,4K]
  const int64_t* gt;   // [R]
  int R, K, agnostic;
  float wx, wy, ww, wh, beta;
  const float* grad_scale;
  float* partial;      // [gridDim.x]
  float* dpred;        // nullable, same shape as pred; must be zero-filled by the caller when not agnostic
};

__global__ void __launch_bounds__(256) box_reg_loss_kernel(BoxRegArgs a) {
  __shared__ float red[8];
  const int r = blockIdx.x * 256 + threadIdx.x;
  float sum = 0.f;
  if (r < a.R) {
    const long long c = a.gt[r];
    const bool fg = c >= 0 && c < a.K;
    const int stride = a.agnostic ? 4 : 4 * a.K;
    float* dp = a.dpred ? a.dpred + (size_t)r * stride + (a.agnostic || !fg ? 0 : 4 * (int)c) : nullptr;
    if (fg) {
      const float4 p = *reinterpret_cast<const float4*>(a.prop + (size_t)r * 4);
      const float4 t = *reinterpret_cast<const float4*>(a.gtb + (size_t)r * 4);
      const float sw = p.z - p.x, sh = p.w - p.y;
      const float sx = p.x + 0.5f * sw, sy = p.y + 0.5f * sh;
      const float tw = t.z - t.x, th = t.w - t.y;
      const float tx = t.x + 0.5f * tw, ty = t.y + 0.5f * th;
      float tgt[4];
      tgt[0] = a.wx * (tx - sx) / sw;
      tgt[1] = a.wy * (ty - sy) / sh;
      tgt[2] = a.ww * logf(tw / sw);
      tgt[3] = a.wh * logf(th / sh);
      const float* pr = a.pred + (size_t)r * stride + (a.agnostic ? 0 : 4 * (int)c);
      const float gs = (a.grad_scale ? *a.grad_scale : 1.f) / (float)a.R;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = pr[k] - tgt[k];
        const float n = fabsf(d);
        float l, g;
        if (a.beta < 1e-5f) {
          l = n;
          g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        } else if (n < a.beta) {
          l = 0.5f * n * n / a.beta;
          g = d / a.beta;
        } else {
          l = n - 0.5f * a.beta;
          g = d > 0.f ? 1.f : -1.f;
        }
        sum += l;
        if (dp) dp[k] = g * gs;
      }
    } else if (dp && a.agnostic) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dp[k] = 0.f;
    }
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    a.partial[blockIdx.x] = t;
  }
}

__global__ void box_reg_finish_kernel(const float* __restrict__ partial, int nblocks, int R, float* __restrict__ loss) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    *loss = t / (float)(R > 0 ? R : 1);
  }
}

}  // namespace cddmsl

using namespace cddmsl;

extern "C" size_t cddmsl_box_reg_loss_workspace_bytes(int R) {
  return align_up((size_t)(R > 0 ? ceil_div(R, 256) : 1) * 4, 256);
}

extern "C" int cddmsl_box_reg_loss(const float* proposal_boxes, const float* gt_boxes, const float* pred_deltas,
                                   const int64_t* gt_classes, int R, int K, int cls_agnostic, float wx, float wy,
                                   float ww, float wh, float beta, const float* grad_scale, float* loss,
                                   float* dpred, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (R < 0 || K <= 0 || !loss) return CDDMSL_EINVAL;
  if (R == 0) {  // empty batch: 0 / max(R, 1)
    CDDMSL_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
    return CDDMSL_OK;
  }
  if (!proposal_boxes || !gt_boxes || !pred_deltas || !gt_classes || !workspace) return CDDMSL_EINVAL;
  if (((reinterpret_cast<uintptr_t>(proposal_boxes) | reinterpret_cast<uintptr_t>(gt_boxes)) & 15) != 0)
    return CDDMSL_EALIGN;
  if (cddmsl_box_reg_loss_workspace_bytes(R) > workspace_bytes) return CDDMSL_EWORKSPACE;
  const int nb = ceil_div(R, 256);
  if (dpred && !cls_agnostic)
    CDDMSL_CUDA(cudaMemsetAsync(dpred, 0, (size_t)R * 4 * K * sizeof(float), stream));
  BoxRegArgs a{proposal_boxes, gt_boxes, pred_deltas, gt_classes, R,     K, cls_agnostic, wx, wy, ww,
               wh,             beta,     grad_scale,  (float*)workspace, dpred};
  box_reg_loss_kernel<<<nb, 256, 0, stream>>>(a);
  box_reg_finish_kernel<<<1, 256, 0, stream>>>((const float*)workspace, nb, R, loss);
  count_launch(2);
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

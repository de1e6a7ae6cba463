// Proposal pre-processing between the RPN head and NMS, one kernel for a batch:
//   RPN._decode_proposals            detectron2/modeling/proposal_generator/rpn.py:514-533
//   Box2BoxTransform.apply_deltas    detectron2/modeling/box_regression.py:77-117
//   finite check / Boxes.clip / Boxes.nonempty / boolean selection   proposal_generator/proposal_utils.py:95-114
// applied to the pre-NMS top-k candidates of every image (decoding commutes with the top-k selection, so only the
// kept candidates are decoded).  One CTA per image walks its K candidates in score order, decodes, clips, tests and
// writes the survivors COMPACTED in their original (score) order -- the stable partition the reference's boolean
// indexing performs -- together with the per-image count and a batch-wide "everything finite" flag (the reference
// raises FloatingPointError in training, proposal_utils.py:100-105).  The arithmetic is the reference's, operation
// for operation, in unfused fp32 (expf = the function torch.exp runs on the device).
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

constexpr int kDecThreads = 1024;

__global__ void __launch_bounds__(kDecThreads) rpn_decode_topk_kernel(
    const float* __restrict__ anchors /* [A,4] */, const float* __restrict__ deltas /* [N,A,4] */,
    const int64_t* __restrict__ topk_idx /* [N,K] */, const float* __restrict__ topk_scores /* [N,K] */,
    const float* __restrict__ image_hw /* [N,2] */, long long A, int K, float wx, float wy, float ww, float wh,
    float scale_clamp, float min_box_size, float* __restrict__ boxes_out /* [N,K,4] */,
    float* __restrict__ scores_out /* [N,K] */, int* __restrict__ counts /* [N] */, int* __restrict__ all_finite) {
  __shared__ int s_warp[kDecThreads / 32];
  __shared__ int s_base;
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float ih = image_hw[2 * n], iw = image_hw[2 * n + 1];
  const float* dl = deltas + (size_t)n * A * 4;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  bool finite_all = true;
  for (int k0 = 0; k0 < K; k0 += kDecThreads) {
    const int k = k0 + threadIdx.x;
    bool sel = false;
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, sc = 0.f;
    if (k < K) {
      const long long a = topk_idx[(size_t)n * K + k];
      sc = topk_scores[(size_t)n * K + k];
      const float4 an = *reinterpret_cast<const float4*>(anchors + a * 4);
      const float4 d = *reinterpret_cast<const float4*>(dl + a * 4);
      // box_regression.py:89-113, term for term (no FMA contraction)
      const float widths = __fsub_rn(an.z, an.x), heights = __fsub_rn(an.w, an.y);
      const float ctr_x = __fadd_rn(an.x, __fmul_rn(0.5f, widths)), ctr_y = __fadd_rn(an.y, __fmul_rn(0.5f, heights));
      const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
      const float dw = fminf(__fdiv_rn(d.z, ww), scale_clamp), dh = fminf(__fdiv_rn(d.w, wh), scale_clamp);
      const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x), pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
      const float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
      x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
      y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
      y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
      // torch.clamp(max=) propagates NaN; fminf does not: a NaN delta must stay NaN for the finite check
      if (d.z != d.z) x1 = x2 = d.z;
      if (d.w != d.w) y1 = y2 = d.w;
      const bool fin = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(sc);
      finite_all = finite_all && fin;
      // Boxes.clip (structures/boxes.py:192-206) then Boxes.nonempty (:208-222)
      x1 = fminf(fmaxf(x1, 0.f), iw);
      y1 = fminf(fmaxf(y1, 0.f), ih);
      x2 = fminf(fmaxf(x2, 0.f), iw);
      y2 = fminf(fmaxf(y2, 0.f), ih);
      sel = fin && (__fsub_rn(x2, x1) > min_box_size) && (__fsub_rn(y2, y1) > min_box_size);
    }
    // stable compaction of this chunk behind the survivors of the previous chunks
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w2 = 0; w2 < warp; ++w2) off += s_warp[w2];
    if (sel) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      *reinterpret_cast<float4*>(boxes_out + ((size_t)n * K + pos) * 4) = make_float4(x1, y1, x2, y2);
      scores_out[(size_t)n * K + pos] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w2 = 0; w2 < kDecThreads / 32; ++w2) t += s_warp[w2];
      s_base += t;
    }
    __syncthreads();
  }
  // rows behind the survivors: zero boxes, -inf scores (never read: NMS takes `counts`)
  for (int k = s_base + threadIdx.x; k < K; k += kDecThreads) {
    *reinterpret_cast<float4*>(boxes_out + ((size_t)n * K + k) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    scores_out[(size_t)n * K + k] = -INFINITY;
  }
  if (threadIdx.x == 0) counts[n] = s_base;
  if (!__syncthreads_and(finite_all) && threadIdx.x == 0) atomicAnd(all_finite, 0);
}

}  // namespace cddmsl

extern "C" int cddmsl_rpn_decode_topk(const float* anchors, const float* deltas, const int64_t* topk_idx,
                                      const float* topk_scores, const float* image_hw, int N, int64_t A, int K,
                                      float wx, float wy, float ww, float wh, float scale_clamp, float min_box_size,
                                      float* boxes_out, float* scores_out, int32_t* counts, int32_t* all_finite,
                                      cddmsl_stream_t stream_) {
  using namespace cddmsl;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || A < 0 || K < 0 || !all_finite) return CDDMSL_EINVAL;
  const int one = 1;
  CDDMSL_CUDA(cudaMemcpyAsync(all_finite, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
  if (N == 0) return CDDMSL_OK;
  if (!counts) return CDDMSL_EINVAL;
  if (K == 0) {
    CDDMSL_CUDA(cudaMemsetAsync(counts, 0, (size_t)N * sizeof(int), stream));
    return CDDMSL_OK;
  }
  if (!anchors || !deltas || !topk_idx || !topk_scores || !image_hw || !boxes_out || !scores_out) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(anchors) & 15) || (reinterpret_cast<uintptr_t>(deltas) & 15) ||
      (reinterpret_cast<uintptr_t>(boxes_out) & 15))
    return CDDMSL_EALIGN;
  rpn_decode_topk_kernel<<<N, kDecThreads, 0, stream>>>(anchors, deltas, topk_idx, topk_scores, image_hw,
                                                        (long long)A, K, wx, wy, ww, wh, scale_clamp, min_box_size,
                                                        boxes_out, scores_out, counts, all_finite);
  count_launch();
  return (int)cudaGetLastError();
}

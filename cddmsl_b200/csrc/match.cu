// Proposal <-> ground-truth matching for sm_100a: pairwise IoU + Matcher in one pass, all images of a batch at once.
//
// Replaces, per image, detectron2/structures/boxes.py:346-368 (pairwise_iou) + detectron2/modeling/matcher.py:63-127
// (Matcher.__call__, set_low_quality_matches_) as called from roi_heads.py:285-289 (label_and_sample_proposals) and
// rpn.py (label_and_sample_anchors).  The reference materialises the [G, M] IoU matrix, reduces it twice and loops over
// the images in Python; here one thread owns one candidate box, walks the image's ground-truth boxes from shared
// memory and emits (argmax gt, label) directly -- the matrix never exists.  IoU arithmetic is the reference's, unfused
// fp32: inter = clamp(min(x2)-max(x1),0) * clamp(min(y2)-max(y1),0); iou = inter > 0 ? inter / (a1 + a2 - inter) : 0,
// so the argmax (first maximum, like torch.max on CPU) and the threshold labels are bit-identical.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

constexpr int kMatchChunk = 1024;  // ground-truth boxes staged per pass
constexpr int kMaxThr = 4;

struct MatchArgs {
  const float4* gt;        // [B][Gmax]
  const int32_t* gcounts;  // [B]
  const float4* boxes;     // [B][Mmax]
  const int32_t* counts;   // [B] or null (every image holds Mmax boxes)
  int B, Gmax, Mmax;
  int nthr;                // thresholds low..high: (-inf, t0, ..., t_{n-1}, +inf)
  float thr[kMaxThr];
  int labels[kMaxThr + 1];
  int64_t* matches;        // [B][Mmax]
  int8_t* match_labels;    // [B][Mmax]
  float* matched_vals;     // [B][Mmax], nullable
  float* gt_best;          // [B][Gmax], nullable: max IoU of every gt over the candidates (low-quality pass)
};

__device__ __forceinline__ float box_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

__device__ __forceinline__ float pair_iou(const float4 g, float area_g, const float4 p, float area_p) {
  const float w = fmaxf(__fsub_rn(fminf(g.z, p.z), fmaxf(g.x, p.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, p.w), fmaxf(g.y, p.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_g, area_p), inter)) : 0.f;
}

// grid (ceil(Mmax/256), B)
__global__ void __launch_bounds__(256) match_kernel(MatchArgs a) {
  __shared__ float4 gbox[kMatchChunk];
  __shared__ float garea[kMatchChunk];
  const int img = blockIdx.y;
  const int G = min(max(a.gcounts[img], 0), a.Gmax);
  const int M = a.counts ? min(max(a.counts[img], 0), a.Mmax) : a.Mmax;
  const int m = blockIdx.x * 256 + threadIdx.x;
  const bool ok = m < M;
  const float4 p = ok ? a.boxes[(size_t)img * a.Mmax + m] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_p = box_area(p);
  float best = -1.f;  // every IoU is >= 0 (or NaN, which never wins -- like torch.max's first-max rule on finite data)
  int best_g = 0;
  for (int g0 = 0; g0 < G; g0 += kMatchChunk) {
    const int n = min(kMatchChunk, G - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) {
      const float4 g = a.gt[(size_t)img * a.Gmax + g0 + i];
      gbox[i] = g;
      garea[i] = box_area(g);
    }
    __syncthreads();
    if (ok)
      for (int i = 0; i < n; ++i) {
        const float v = pair_iou(gbox[i], garea[i], p, area_p);
        if (v > best) {  // strict: the first maximum wins
          best = v;
          best_g = g0 + i;
        }
      }
  }
  if (!ok) return;
  const size_t o = (size_t)img * a.Mmax + m;
  int label;
  if (G == 0) {  // matcher.py:78-86: empty matrix -> match 0, label of the lowest stratum
    best = 0.f;
    best_g = 0;
    label = a.labels[0];
  } else {
    label = 1;  // matcher.py:93 initial fill; every finite value falls in exactly one stratum below
    float low = -INFINITY;
    for (int t = 0; t <= a.nthr; ++t) {
      const float high = t < a.nthr ? a.thr[t] : INFINITY;
      if (best >= low && best < high) label = a.labels[t];
      low = high;
    }
  }
  a.matches[o] = best_g;
  a.match_labels[o] = (int8_t)label;
  if (a.matched_vals) a.matched_vals[o] = best;
}

// grid (Gmax, B): highest IoU of each ground-truth box over the image's candidates (matcher.py:117)
__global__ void __launch_bounds__(256) match_gt_best_kernel(MatchArgs a) {
  __shared__ float red[8];
  const int img = blockIdx.y, g = blockIdx.x;
  const int G = min(max(a.gcounts[img], 0), a.Gmax);
  if (g >= G) return;
  const int M = a.counts ? min(max(a.counts[img], 0), a.Mmax) : a.Mmax;
  const float4 gb = a.gt[(size_t)img * a.Gmax + g];
  const float ag = box_area(gb);
  float best = -INFINITY;
  for (int m = threadIdx.x; m < M; m += 256) {
    const float4 p = a.boxes[(size_t)img * a.Mmax + m];
    best = fmaxf(best, pair_iou(gb, ag, p, box_area(p)));
  }
  best = warp_max(best);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    a.gt_best[(size_t)img * a.Gmax + g] = t;
  }
}

// grid (ceil(Mmax/256), B): label 1 for every candidate that attains some gt's highest IoU, ties included (:118-127)
__global__ void __launch_bounds__(256) match_low_quality_kernel(MatchArgs a) {
  __shared__ float4 gbox[kMatchChunk];
  __shared__ float garea[kMatchChunk];
  __shared__ float gbest[kMatchChunk];
  const int img = blockIdx.y;
  const int G = min(max(a.gcounts[img], 0), a.Gmax);
  const int M = a.counts ? min(max(a.counts[img], 0), a.Mmax) : a.Mmax;
  const int m = blockIdx.x * 256 + threadIdx.x;
  const bool ok = m < M;
  const float4 p = ok ? a.boxes[(size_t)img * a.Mmax + m] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_p = box_area(p);
  bool hit = false;
  for (int g0 = 0; g0 < G; g0 += kMatchChunk) {
    const int n = min(kMatchChunk, G - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) {
      const float4 g = a.gt[(size_t)img * a.Gmax + g0 + i];
      gbox[i] = g;
      garea[i] = box_area(g);
      gbest[i] = a.gt_best[(size_t)img * a.Gmax + g0 + i];
    }
    __syncthreads();
    if (ok)
      for (int i = 0; i < n; ++i) hit |= pair_iou(gbox[i], garea[i], p, area_p) == gbest[i];
  }
  if (ok && hit) a.match_labels[(size_t)img * a.Mmax + m] = 1;
}

}  // namespace cddmsl

using namespace cddmsl;

extern "C" size_t cddmsl_match_boxes_workspace_bytes(int B, int Gmax) {
  return align_up((size_t)(B > 0 ? B : 1) * (size_t)(Gmax > 0 ? Gmax : 1) * 4, 256);
}

extern "C" int cddmsl_match_boxes(const float* gt_boxes, const int32_t* gt_counts, const float* boxes,
                                  const int32_t* counts, int B, int Gmax, int Mmax, const float* thresholds,
                                  const int32_t* labels, int num_thresholds, int allow_low_quality_matches,
                                  int64_t* matches, int8_t* match_labels, float* matched_vals, void* workspace,
                                  size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B < 0 || Gmax < 0 || Mmax < 0 || num_thresholds < 1 || num_thresholds > kMaxThr || !thresholds || !labels)
    return CDDMSL_EINVAL;
  if (B == 0 || Mmax == 0) return CDDMSL_OK;
  if (!gt_counts || !boxes || !matches || !match_labels || (Gmax > 0 && !gt_boxes)) return CDDMSL_EINVAL;
  if (B > 65535 || Gmax > 65535) return CDDMSL_EINVAL;
  if (((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(gt_boxes)) & 15) != 0) return CDDMSL_EALIGN;
  MatchArgs a{};
  a.gt = reinterpret_cast<const float4*>(gt_boxes);
  a.gcounts = gt_counts;
  a.boxes = reinterpret_cast<const float4*>(boxes);
  a.counts = counts;
  a.B = B;
  a.Gmax = Gmax;
  a.Mmax = Mmax;
  a.nthr = num_thresholds;
  for (int t = 0; t < num_thresholds; ++t) {
    a.thr[t] = thresholds[t];
    if (t > 0 && thresholds[t] < thresholds[t - 1]) return CDDMSL_EINVAL;  // matcher.py:52
  }
  if (!(thresholds[0] > 0.f)) return CDDMSL_EINVAL;                       // matcher.py:49
  for (int t = 0; t <= num_thresholds; ++t) {
    if (labels[t] < -1 || labels[t] > 1) return CDDMSL_EINVAL;            // matcher.py:53
    a.labels[t] = labels[t];
  }
  a.matches = matches;
  a.match_labels = match_labels;
  a.matched_vals = matched_vals;
  a.gt_best = nullptr;
  const dim3 grid(ceil_div(Mmax, 256), B);
  match_kernel<<<grid, 256, 0, stream>>>(a);
  count_launch();
  if (allow_low_quality_matches && Gmax > 0) {
    if (!workspace || cddmsl_match_boxes_workspace_bytes(B, Gmax) > workspace_bytes) return CDDMSL_EWORKSPACE;
    a.gt_best = (float*)workspace;
    match_gt_best_kernel<<<dim3(Gmax, B), 256, 0, stream>>>(a);
    match_low_quality_kernel<<<grid, 256, 0, stream>>>(a);
    count_launch(2);
  }
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

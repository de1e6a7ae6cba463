// Class-aware greedy NMS for sm_100a, bit-exact with the CPU kernel the oracle runs.
//
// Replaces torchvision.ops.boxes.batched_nms / torchvision::nms as called from
// detectron2/layers/nms.py:19-39 (call sites proposal_utils.py:116, fast_rcnn.py:184).
//
// Pipeline, all on the caller's stream, no host synchronisation:
//   1. (coordinate-trick mode only) max over all coordinates               -> nms_max_kernel
//   2. stable descending radix sort of (score, index)                      -> cub::DeviceRadixSort
//   3. gather boxes (+ class offset) / class ids into sorted order          -> nms_gather_kernel
//   4. 64x64-tile IoU bitmask, upper triangle only; diagonal tiles also emit the column view of the
//      suppression relation with warp ballots                              -> nms_mask_kernel
//   5. greedy scan on the device: per 64-box block a warp-ballot fix-point resolves the diagonal tile, the
//      kept rows (mask words prefetched one block ahead) are OR-ed into the removed-bitmap, kept indices are
//      emitted in order                                                    -> nms_scan_kernel
// Every kernel takes an image index from the grid (blockIdx.z, or .x for the one-CTA-per-image kernels) and a
// per-image box count read from device memory, so ONE launch sequence serves the B images of an RPN batch
// (cddmsl_nms_batched: padded [B][Mmax] layout, counts on the device, no host sync; the B serial greedy scans
// run concurrently on B SMs).  cddmsl_nms is the B = 1 case.
// IoU arithmetic is fp32 without FMA contraction ((area_i + area_j) - inter, IEEE division) and the
// threshold test promotes the fp32 IoU to double, exactly like the CPU kernel (oracle/c/nms_ref.c).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace cddmsl {

constexpr int kTile = 64;

// per-image box count: device array (batched entry point) or the host-known M
__device__ __forceinline__ int image_count(const int32_t* __restrict__ counts, int img, int m_fixed, int m_max) {
  return counts ? min(max(counts[img], 0), m_max) : m_fixed;
}

__global__ void nms_max_kernel(const float* __restrict__ boxes_all, const int32_t* __restrict__ counts, int m_fixed,
                               int m_max, float* __restrict__ out_max_all) {
  // one block per image; M*4 values.  -inf start; result is the exact max (order-independent).
  __shared__ float red[32];
  const int img = blockIdx.x;
  const float* boxes = boxes_all + (size_t)img * m_max * 4;
  float* out_max = out_max_all + img;
  const int64_t n4 = (int64_t)image_count(counts, img, m_fixed, m_max) * 4;
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) m = fmaxf(m, boxes[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    m = warp_max(m);
    if (threadIdx.x == 0) *out_max = m;
  }
}

__global__ void nms_iota_kernel(const float* __restrict__ scores, float* __restrict__ keys, int* __restrict__ vals,
                                const int32_t* __restrict__ counts, int m_fixed, int m_max,
                                int* __restrict__ seg_begin, int* __restrict__ seg_end) {
  const int img = blockIdx.z;
  const int M = image_count(counts, img, m_fixed, m_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && seg_begin) {  // segment [img*m_max, img*m_max + M) of the padded arrays
    seg_begin[img] = img * m_max;
    seg_end[img] = img * m_max + M;
  }
  if (i < M) {
    const size_t g = (size_t)img * m_max + i;
    float s = scores[g];
    keys[g] = (s == 0.f) ? 0.f : s;  // -0.0 and +0.0 compare equal in the reference's sort
    vals[g] = i;
  }
}

// scores already non-increasing per image (the RPN path hands over its sorted top-k): the sorted order is the identity
__global__ void nms_identity_kernel(int* __restrict__ order, const int32_t* __restrict__ counts, int m_fixed,
                                    int m_max) {
  const int img = blockIdx.z;
  const int M = image_count(counts, img, m_fixed, m_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) order[(size_t)img * m_max + i] = i;
}

__global__ void nms_gather_kernel(const float4* __restrict__ boxes_all, const int64_t* __restrict__ idxs_all,
                                  const int* __restrict__ order_all, const float* __restrict__ max_coord_all,
                                  int coord_trick, float4* __restrict__ sboxes_all, int* __restrict__ scls_all,
                                  const int32_t* __restrict__ counts, int m_fixed, int m_max) {
  const int img = blockIdx.z;
  const int M = image_count(counts, img, m_fixed, m_max);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const size_t ib = (size_t)img * m_max;
  const float4* boxes = boxes_all + ib;
  const int64_t* idxs = idxs_all ? idxs_all + ib : nullptr;
  const int* order = order_all + ib;
  const float* max_coord = max_coord_all + img;
  float4* sboxes = sboxes_all + ib;
  int* scls = scls_all + ib;
  const int o = order[i];
  float4 b = boxes[o];
  int cls = 0;
  if (idxs) {
    const int64_t id = idxs[o];
    if (coord_trick) {
      // offsets = idxs.to(boxes) * (max_coordinate + 1);  boxes + offsets[:, None]   (fp32, rn)
      const float off = __fmul_rn((float)id, __fadd_rn(*max_coord, 1.0f));
      b.x = __fadd_rn(b.x, off);
      b.y = __fadd_rn(b.y, off);
      b.z = __fadd_rn(b.z, off);
      b.w = __fadd_rn(b.w, off);
    } else {
      cls = (int)id;
    }
  }
  sboxes[i] = b;
  scls[i] = cls;
}

__device__ __forceinline__ bool iou_over(const float4 a, const float4 b, float area_a, float area_b, float thr_dn) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  // disjoint boxes (the vast majority of pairs): 0 / union is 0, or NaN for 0 / 0 -- never above a non-negative
  // threshold.  Same verdict as the division below, without the IEEE divide.
  if (inter == 0.f && thr_dn >= 0.f) return false;
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  // (double)ovr > thr  <=>  ovr > thr_dn with thr_dn = the largest float <= thr (round_down_threshold): exact, and
  // keeps the comparison off the FP64 pipe.  NaN (0/0) compares false: never suppresses.
  return ovr > thr_dn;
}

// grid (col_blocks, row_blocks), 64 threads; only tiles with col_block >= row_block are computed.
// Row masks: mask[i][cb] bit j = box (cb*64+j) is suppressed by box i (j > i inside a diagonal tile).
// Diagonal tiles additionally emit the COLUMN view through warp ballots: coldiag[rb*64+j] bit i = row i (< j, same
// tile) suppresses j — what the scan kernel's ballot fix-point needs.
// one 64 x 64 tile (rb, cb) of image img; M = boxes of the image that take part.  Called by all 64 threads of a CTA.
__device__ __forceinline__ void nms_mask_tile(const float4* __restrict__ sboxes_all, const int* __restrict__ scls_all,
                                              int m_max, float thr, int col_blocks,
                                              unsigned long long* __restrict__ mask_all,
                                              unsigned long long* __restrict__ coldiag_all, int img, int M, int rb,
                                              int cb) {
  // col_blocks = ceil(m_max / 64): row stride of the mask for every image
  if (cb < rb || cb * kTile >= M) return;
  const float4* sboxes = sboxes_all + (size_t)img * m_max;
  const int* scls = scls_all + (size_t)img * m_max;
  unsigned long long* mask = mask_all + (size_t)img * m_max * col_blocks;
  unsigned long long* coldiag = coldiag_all + (size_t)img * col_blocks * kTile;
  __shared__ float4 cbox[kTile];
  __shared__ float carea[kTile];
  __shared__ int ccls[kTile];
  __shared__ unsigned int colpart[2][kTile];
  const int ncol = min(M - cb * kTile, kTile);
  if (threadIdx.x < ncol) {
    const float4 b = sboxes[cb * kTile + threadIdx.x];
    cbox[threadIdx.x] = b;
    carea[threadIdx.x] = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    ccls[threadIdx.x] = scls[cb * kTile + threadIdx.x];
  }
  __syncthreads();
  const int i = rb * kTile + threadIdx.x;
  const bool row_ok = i < M;
  const float4 a = row_ok ? sboxes[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const int cls_a = row_ok ? scls[i] : -1;
  unsigned long long bits = 0ull;
  if (rb != cb) {
    if (!row_ok) return;
    for (int j = 0; j < ncol; ++j)
      if (ccls[j] == cls_a && iou_over(a, cbox[j], area_a, carea[j], thr)) bits |= 1ull << j;
    mask[(size_t)i * col_blocks + cb] = bits;
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < kTile; ++j) {  // uniform trip count: every lane takes part in every ballot
    const bool p = row_ok && j < ncol && j > (int)threadIdx.x && ccls[j] == cls_a &&
                   iou_over(a, cbox[j], area_a, carea[j], thr);
    if (p) bits |= 1ull << j;
    const unsigned int m = __ballot_sync(0xffffffffu, p);
    if (lane == 0) colpart[warp][j] = m;
  }
  if (row_ok) mask[(size_t)i * col_blocks + cb] = bits;
  __syncthreads();
  coldiag[(size_t)rb * kTile + threadIdx.x] =
      (unsigned long long)colpart[0][threadIdx.x] | ((unsigned long long)colpart[1][threadIdx.x] << 32);
}


// m_limit: only the m_limit best-scoring boxes of an image take part (top-k-limited NMS, see nms_run)
__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float4* __restrict__ sboxes_all, const int* __restrict__ scls_all,
                const int32_t* __restrict__ counts, int m_fixed, int m_max, float thr, int col_blocks,
                unsigned long long* __restrict__ mask_all, unsigned long long* __restrict__ coldiag_all, int m_limit) {
  const int img = blockIdx.z;
  const int M = min(image_count(counts, img, m_fixed, m_max), m_limit);
  nms_mask_tile(sboxes_all, scls_all, m_max, thr, col_blocks, mask_all, coldiag_all, img, M, blockIdx.y, blockIdx.x);
}

// The full pass behind a top-k-limited first pass: normally no image needs it, and a grid of col_blocks^2 x B CTAs that
// all exit at once still costs 0.29 ms (RPN batch 16 x 12 000).  So a small fixed grid looks at the per-image flags
// and walks the tiles of the images that do need the full mask.
__global__ void __launch_bounds__(kTile)
nms_mask_needed_kernel(const float4* __restrict__ sboxes_all, const int* __restrict__ scls_all,
                       const int32_t* __restrict__ counts, int m_fixed, int m_max, float thr, int col_blocks,
                       unsigned long long* __restrict__ mask_all, unsigned long long* __restrict__ coldiag_all, int B,
                       const int32_t* __restrict__ need) {
  for (int img = 0; img < B; ++img) {
    if (!need[img]) continue;
    const int M = image_count(counts, img, m_fixed, m_max);
    const int nblk = (M + kTile - 1) / kTile;
    for (int t = blockIdx.x; t < nblk * nblk; t += gridDim.x) {
      nms_mask_tile(sboxes_all, scls_all, m_max, thr, col_blocks, mask_all, coldiag_all, img, M, t / nblk, t % nblk);
      __syncthreads();  // the tile's shared arrays are reused by the next one
    }
  }
}

// One CTA, 1024 threads.  remv[] (shared) = bitmap of suppressed boxes.  Per 64-box block b:
//   * warp 0 resolves the block with a ballot fix-point on the column masks: kept_{n+1} = alive & ~{ j : some kept_n
//     row i < j suppresses j } — its unique fixed point is the greedy result, reached in (longest suppression chain
//     + 1) rounds of two ballots instead of up to 64 dependent steps;
//   * every thread owns (16 rows, 1 column word) of the block's rows and has ALREADY loaded those 16 mask words one
//     block ahead (the loads do not depend on the scan state), so applying the kept rows to the removed-bitmap is
//     register work; column words beyond the first 256 are fetched in batches of 16 independent loads.
__global__ void __launch_bounds__(1024)
nms_scan_kernel(const unsigned long long* __restrict__ mask_all, const unsigned long long* __restrict__ coldiag_all,
                const int* __restrict__ order_all, const int32_t* __restrict__ counts, int m_fixed, int m_max,
                int stride_blocks, int64_t* __restrict__ keep_all, int32_t* __restrict__ num_keep_all, int m_limit,
                int max_keep, const int32_t* __restrict__ need_in, int32_t* __restrict__ need_out) {
  extern __shared__ unsigned long long remv[];  // [stride_blocks]
  const int img = blockIdx.x;
  if (need_in && !need_in[img]) return;   // the limited pass already produced max_keep boxes for this image
  const int m_full = image_count(counts, img, m_fixed, m_max);
  const int M = min(m_full, m_limit);
  const int col_blocks = (M + kTile - 1) / kTile;  // this image's blocks; rows are stride_blocks words apart
  const unsigned long long* mask = mask_all + (size_t)img * m_max * stride_blocks;
  const unsigned long long* coldiag = coldiag_all + (size_t)img * stride_blocks * kTile;
  const int* order = order_all + (size_t)img * m_max;
  int64_t* keep = keep_all + (size_t)img * m_max;
  int32_t* num_keep = num_keep_all + img;
  __shared__ unsigned long long kept_s;
  __shared__ int count_s;
  for (int w = threadIdx.x; w < col_blocks; w += blockDim.x) remv[w] = 0ull;
  if (threadIdx.x == 0) count_s = 0;
  const int slot = threadIdx.x >> 8;    // 0..3: rows t with t % 4 == slot
  const int wlane = threadIdx.x & 255;  // column word within a pass of 256
  const int lane = threadIdx.x & 31;
  unsigned int* remv32 = reinterpret_cast<unsigned int*>(remv);

  // v[u] = mask word (row b*64 + 4u + slot, column word w), 0 outside the matrix
#define NMS_LOAD_ROWS(B, WORD, V)                                                             \
  _Pragma("unroll") for (int u = 0; u < 16; ++u) {                                            \
    const int row_ = (B) * kTile + 4 * u + slot;                                              \
    (V)[u] = ((WORD) < col_blocks && row_ < M) ? mask[(size_t)row_ * stride_blocks + (WORD)] : 0ull; \
  }
  unsigned long long v[16];  // prefetched for the block about to be resolved
  NMS_LOAD_ROWS(0, 1 + wlane, v)
  unsigned long long cm_lo = 0ull, cm_hi = 0ull;  // column masks of block 0 for warp 0
  if (threadIdx.x < 32 && col_blocks > 0) {
    cm_lo = coldiag[lane];
    cm_hi = coldiag[lane + 32];
  }
  __syncthreads();
  for (int b = 0; b < col_blocks; ++b) {
    const int nb = min(M - b * kTile, kTile);
    if (threadIdx.x < 32) {
      unsigned long long alive = ~remv[b];
      if (nb < kTile) alive &= (1ull << nb) - 1ull;
      unsigned long long kept = alive;
      for (int it = 0; it < kTile + 1; ++it) {
        const unsigned int s_lo = __ballot_sync(0xffffffffu, (cm_lo & kept) != 0ull);
        const unsigned int s_hi = __ballot_sync(0xffffffffu, (cm_hi & kept) != 0ull);
        const unsigned long long nk = alive & ~((unsigned long long)s_lo | ((unsigned long long)s_hi << 32));
        if (nk == kept) break;
        kept = nk;
      }
      if (lane == 0) kept_s = kept;
      if (b + 1 < col_blocks) {  // next block's column masks, consumed one iteration later
        cm_lo = coldiag[(size_t)(b + 1) * kTile + lane];
        cm_hi = coldiag[(size_t)(b + 1) * kTile + lane + 32];
      }
    }
    __syncthreads();
    const unsigned long long kept = kept_s;
    const int base = count_s;
    if (threadIdx.x < kTile && ((kept >> threadIdx.x) & 1ull)) {
      const int pos = base + __popcll(kept & ((1ull << threadIdx.x) - 1ull));
      keep[pos] = (int64_t)order[b * kTile + threadIdx.x];
    }
    {  // first pass of column words: the prefetched registers
      const int w = b + 1 + wlane;
      unsigned long long acc = 0ull;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if ((kept >> (4 * u + slot)) & 1ull) acc |= v[u];
      if (acc && w < col_blocks) {
        atomicOr(remv32 + 2 * w, (unsigned int)acc);
        atomicOr(remv32 + 2 * w + 1, (unsigned int)(acc >> 32));
      }
    }
    for (int w = b + 1 + wlane + 256; w < col_blocks; w += 256) {  // wider problems: 16 independent loads per pass
      NMS_LOAD_ROWS(b, w, v)
      unsigned long long acc = 0ull;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if ((kept >> (4 * u + slot)) & 1ull) acc |= v[u];
      if (acc) {
        atomicOr(remv32 + 2 * w, (unsigned int)acc);
        atomicOr(remv32 + 2 * w + 1, (unsigned int)(acc >> 32));
      }
    }
    // the next block's rows do not depend on the scan state: fetch them now, use them after the next resolve
    if (b + 1 < col_blocks) {
      NMS_LOAD_ROWS(b + 1, b + 2 + wlane, v)
    }
    __syncthreads();
    if (threadIdx.x == 0) count_s = base + __popcll(kept);
    // (count_s and kept_s are next touched after the barrier above / the next iteration's barrier)
    if (max_keep > 0 && base + __popcll(kept) >= max_keep) break;  // uniform: the caller wants no more than max_keep
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = count_s;
    *num_keep = max_keep > 0 ? min(total, max_keep) : total;
    // the limited pass ran out of boxes before max_keep were kept: the full pass has to redo this image
    if (need_out) need_out[img] = (max_keep > 0 && total < max_keep && M < m_full) ? 1 : 0;
  }
}

struct NmsWs {
  float* keys_in;
  float* keys_out;
  int* vals_in;
  int* order;
  float4* sboxes;
  int* scls;
  float* max_coord;
  int* seg_begin;
  int* seg_end;
  unsigned long long* mask;
  unsigned long long* coldiag;
  int32_t* need;           // top-k-limited NMS: images the limited pass could not finish
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static NmsWs carve(void* base, int64_t B, int64_t M) {
  NmsWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t b = (size_t)(B > 0 ? B : 1);
  const size_t m = (size_t)(M > 0 ? M : 1);
  const size_t cb = (m + kTile - 1) / kTile;
  w.keys_in = (float*)take(b * m * 4);
  w.keys_out = (float*)take(b * m * 4);
  w.vals_in = (int*)take(b * m * 4);
  w.order = (int*)take(b * m * 4);
  w.sboxes = (float4*)take(b * m * 16);
  w.scls = (int*)take(b * m * 4);
  w.max_coord = (float*)take(b * 4);
  w.seg_begin = (int*)take(b * 4);
  w.seg_end = (int*)take(b * 4);
  w.mask = (unsigned long long*)take(b * m * cb * 8);
  w.coldiag = (unsigned long long*)take(b * cb * kTile * 8);
  w.need = (int32_t*)take(b * 4);
  w.cub_bytes = 0;
  if (b == 1) {
    cub::DeviceRadixSort::SortPairsDescending(nullptr, w.cub_bytes, (const float*)nullptr, (float*)nullptr,
                                              (const int*)nullptr, (int*)nullptr, (int)m, 0, 32, (cudaStream_t)0);
  } else {
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, w.cub_bytes, (const float*)nullptr, (float*)nullptr,
                                                       (const int*)nullptr, (int*)nullptr, (int)(b * m), (int)b,
                                                       (const int*)nullptr, (const int*)nullptr, 0, 32,
                                                       (cudaStream_t)0);
  }
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

// The reference compares the fp32 IoU with the threshold in DOUBLE.  For a float v and a double t:
// (double)v > t  <=>  v > f, where f is the largest float <= t (if t is a float, f == t; otherwise f < t < next(f) and a
// float exceeds t exactly when it is >= next(f), i.e. > f).  NaN / +-inf thresholds keep their meaning.
static float round_down_threshold(double t) {
  float f = (float)t;  // round to nearest
  if ((double)f > t) f = nextafterf(f, -INFINITY);
  return f;
}

// B images, padded to m_max boxes each; counts == nullptr: every image holds exactly m_max boxes (the B = 1 call)
// max_keep > 0: only the first max_keep entries of the keep list are wanted (find_top_rpn_proposals keeps
// post_nms_topk = 1000 / 2000 of up to 12 000 candidates, proposal_utils.py:116-118).  Greedy NMS never looks ahead,
// so those entries depend on the best-scoring boxes only: a first pass works on the top 2 * max_keep boxes (mask work
// falls with the square of that), and a second pass over everything runs only for the images where the first one ran
// out of boxes before max_keep were kept (decided on the device: no host synchronisation).
static int nms_run(const float* boxes, const float* scores, const int64_t* idxs, const int32_t* counts, int B,
                   int m_max, double iou_threshold, int coord_trick, int64_t* keep, int32_t* num_keep, const NmsWs& w,
                   cudaStream_t stream, int max_keep = 0, int presorted = 0) {
  const int col_blocks = ceil_div(m_max, kTile);
  if ((size_t)col_blocks * 8 > 200 * 1024) return CDDMSL_EINVAL;  // removed-bitmap must fit shared memory
  if (idxs && coord_trick) {
    nms_max_kernel<<<B, 1024, 0, stream>>>(boxes, counts, m_max, m_max, w.max_coord);
    count_launch();
  }
  if (presorted) {
    // a stable sort of an already sorted sequence is the identity: skip the radix sort (0.15 of 0.48 ms for an RPN batch)
    nms_identity_kernel<<<dim3(ceil_div(m_max, 256), 1, B), 256, 0, stream>>>(w.order, counts, m_max, m_max);
    count_launch();
  } else {
    nms_iota_kernel<<<dim3(ceil_div(m_max, 256), 1, B), 256, 0, stream>>>(scores, w.keys_in, w.vals_in, counts, m_max,
                                                                          m_max, counts ? w.seg_begin : nullptr,
                                                                          w.seg_end);
    count_launch();
    size_t cub_bytes = w.cub_bytes;
    if (!counts) {  // B == 1, host-known size
      CDDMSL_CUDA(cub::DeviceRadixSort::SortPairsDescending(w.cub_temp, cub_bytes, w.keys_in, w.keys_out, w.vals_in,
                                                            w.order, m_max, 0, 32, stream));
    } else {
      CDDMSL_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(w.cub_temp, cub_bytes, w.keys_in, w.keys_out,
                                                                     w.vals_in, w.order, B * m_max, B, w.seg_begin,
                                                                     w.seg_end, 0, 32, stream));
    }
    count_launch(3);  // histogram + onesweep passes (CUB internal; counted as one logical sort)
  }
  nms_gather_kernel<<<dim3(ceil_div(m_max, 256), 1, B), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(boxes), idxs, w.order, w.max_coord, coord_trick, w.sboxes, w.scls, counts,
      m_max, m_max);
  count_launch();
  const float thr = round_down_threshold(iou_threshold);
  const int smem = col_blocks * 8;
  if (smem > 48 * 1024)
    CDDMSL_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int m_first = max_keep > 0 ? min(m_max, ceil_div(2 * max_keep, kTile) * kTile) : m_max;
  const bool limited = m_first < m_max;
  if (limited) {
    const int cb1 = ceil_div(m_first, kTile);
    nms_mask_kernel<<<dim3(cb1, cb1, B), kTile, 0, stream>>>(w.sboxes, w.scls, counts, m_max, m_max, thr, col_blocks,
                                                             w.mask, w.coldiag, m_first);
    nms_scan_kernel<<<B, 1024, smem, stream>>>(w.mask, w.coldiag, w.order, counts, m_max, m_max, col_blocks, keep,
                                               num_keep, m_first, max_keep, nullptr, w.need);
    count_launch(2);
  }
  if (limited)
    nms_mask_needed_kernel<<<sm_count() * 16, kTile, 0, stream>>>(w.sboxes, w.scls, counts, m_max, m_max, thr,
                                                                 col_blocks, w.mask, w.coldiag, B, w.need);
  else
    nms_mask_kernel<<<dim3(col_blocks, col_blocks, B), kTile, 0, stream>>>(w.sboxes, w.scls, counts, m_max, m_max, thr,
                                                                           col_blocks, w.mask, w.coldiag, m_max);
  count_launch();
  nms_scan_kernel<<<B, 1024, smem, stream>>>(w.mask, w.coldiag, w.order, counts, m_max, m_max, col_blocks, keep,
                                             num_keep, m_max, max_keep, limited ? w.need : nullptr, nullptr);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

}  // namespace cddmsl

using namespace cddmsl;

extern "C" size_t cddmsl_nms_workspace_bytes(int64_t M) { return carve(nullptr, 1, M).total; }

extern "C" int cddmsl_nms_topk(const float* boxes, const float* scores, const int64_t* idxs, int64_t M,
                               double iou_threshold, int coord_trick, int max_keep, int presorted, int64_t* keep,
                               int32_t* num_keep, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_);

extern "C" int cddmsl_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t M,
                          double iou_threshold, int coord_trick, int64_t* keep, int32_t* num_keep, void* workspace,
                          size_t workspace_bytes, cddmsl_stream_t stream_) {
  return cddmsl_nms_topk(boxes, scores, idxs, M, iou_threshold, coord_trick, 0, 0, keep, num_keep, workspace,
                         workspace_bytes, stream_);
}

extern "C" int cddmsl_nms_topk(const float* boxes, const float* scores, const int64_t* idxs, int64_t M,
                               double iou_threshold, int coord_trick, int max_keep, int presorted, int64_t* keep,
                               int32_t* num_keep, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M < 0 || !num_keep) return CDDMSL_EINVAL;
  if (M == 0) {
    CDDMSL_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return CDDMSL_OK;
  }
  if (M > (1 << 24)) return CDDMSL_EINVAL;
  if (!boxes || !scores || !keep || !workspace) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return CDDMSL_EALIGN;
  NmsWs w = carve(workspace, 1, M);
  if (w.total > workspace_bytes) return CDDMSL_EWORKSPACE;
  return nms_run(boxes, scores, idxs, nullptr, 1, (int)M, iou_threshold, coord_trick, keep, num_keep, w, stream,
                 max_keep, presorted);
}

extern "C" size_t cddmsl_nms_batched_workspace_bytes(int B, int64_t Mmax) {
  return carve(nullptr, B < 2 ? 2 : B, Mmax).total;
}

extern "C" int cddmsl_nms_batched_topk(const float* boxes, const float* scores, const int64_t* idxs,
                                       const int32_t* counts, int B, int64_t Mmax, double iou_threshold,
                                       int coord_trick, int max_keep, int presorted, int64_t* keep,
                                       int32_t* num_keep, void* workspace, size_t workspace_bytes,
                                       cddmsl_stream_t stream_);

extern "C" int cddmsl_nms_batched(const float* boxes, const float* scores, const int64_t* idxs,
                                  const int32_t* counts, int B, int64_t Mmax, double iou_threshold, int coord_trick,
                                  int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                                  cddmsl_stream_t stream_) {
  return cddmsl_nms_batched_topk(boxes, scores, idxs, counts, B, Mmax, iou_threshold, coord_trick, 0, 0, keep,
                                 num_keep, workspace, workspace_bytes, stream_);
}

extern "C" int cddmsl_nms_batched_topk(const float* boxes, const float* scores, const int64_t* idxs,
                                       const int32_t* counts, int B, int64_t Mmax, double iou_threshold,
                                       int coord_trick, int max_keep, int presorted, int64_t* keep,
                                       int32_t* num_keep, void* workspace, size_t workspace_bytes,
                                       cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B < 0 || Mmax < 0 || (B > 0 && !num_keep)) return CDDMSL_EINVAL;
  if (B == 0) return CDDMSL_OK;
  if (Mmax == 0) {
    CDDMSL_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * B, stream));
    return CDDMSL_OK;
  }
  if (Mmax > (1 << 24) || (int64_t)B * Mmax > (1ll << 30) || B > 65535) return CDDMSL_EINVAL;
  if (!boxes || !scores || !keep || !workspace || !counts) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return CDDMSL_EALIGN;
  const int bw = B < 2 ? 2 : B;  // the segmented sort also serves B == 1 with a device-side count
  NmsWs w = carve(workspace, bw, Mmax);
  if (w.total > workspace_bytes) return CDDMSL_EWORKSPACE;
  return nms_run(boxes, scores, idxs, counts, B, (int)Mmax, iou_threshold, coord_trick, keep, num_keep, w, stream,
                 max_keep, presorted);
}

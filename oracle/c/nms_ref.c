/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C restatement of greedy NMS as the reference reaches it through
 * `detectron2/layers/nms.py:19-39` -> `torchvision.ops.boxes.batched_nms` -> `torchvision.ops.nms`
 * (CPU kernel; torchvision 0.26.0 in this image, un-pinned by the reference).  The in-tree twin that
 * documents the same algorithm is `detectron2/layers/csrc/nms_rotated/nms_rotated_cuda.cu:21-143`.
 *
 * Parity-defining details restated here:
 *   - order = stable sort by score, descending (ties: lower index first);
 *   - area = (x2-x1)*(y2-y1) in fp32; inter = max(0,xx2-xx1)*max(0,yy2-yy1);
 *     ovr = inter / (area_i + area_j - inter), no FMA contraction;
 *   - suppression test is `ovr > iou_threshold` with the fp32 ovr promoted to DOUBLE and compared
 *     against the double threshold (the CPU kernel receives a double);  NaN ovr (0/0) never suppresses;
 *   - batched_nms: classes never suppress each other.  `coord_trick` != 0 reproduces
 *     `_batched_nms_coordinate_trick` (boxes + idx*(max_coord+1) in fp32, torchvision/ops/boxes.py),
 *     otherwise the per-class loop `_batched_nms_vanilla` (raw coordinates);
 *   - output: kept indices ordered by score descending, ties by index ascending.
 *
 * Pinned by tests/test_oracle.py against the installed torchvision CPU op (fixtures in tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float s;
  int64_t i;
} item_t;

static int cmp_desc_stable(const void* a, const void* b) {
  const item_t* x = (const item_t*)a;
  const item_t* y = (const item_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* boxes [M,4] xyxy fp32, scores [M], idxs [M] int64 or NULL; keep [M] int64 out; returns #kept. */
int64_t oracle_batched_nms(const float* boxes_in, const float* scores, const int64_t* idxs, int64_t M,
                           double iou_threshold, int coord_trick, int64_t* keep) {
  if (M <= 0) return 0;
  float* boxes = (float*)malloc((size_t)M * 4 * sizeof(float));
  memcpy(boxes, boxes_in, (size_t)M * 4 * sizeof(float));
  if (idxs && coord_trick) {
    float mx = boxes[0];
    for (int64_t k = 1; k < 4 * M; k++) mx = boxes[k] > mx ? boxes[k] : mx;
    const float step = mx + 1.0f;
    for (int64_t i = 0; i < M; i++) {
      const float off = (float)idxs[i] * step;
      for (int k = 0; k < 4; k++) boxes[4 * i + k] = boxes[4 * i + k] + off;
    }
  }
  const int class_aware = idxs && !coord_trick;
  item_t* order = (item_t*)malloc((size_t)M * sizeof(item_t));
  float* area = (float*)malloc((size_t)M * sizeof(float));
  uint8_t* sup = (uint8_t*)calloc((size_t)M, 1);
  for (int64_t i = 0; i < M; i++) {
    order[i].s = scores[i];
    order[i].i = i;
    area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  qsort(order, (size_t)M, sizeof(item_t), cmp_desc_stable);
  int64_t nk = 0;
  for (int64_t a = 0; a < M; a++) {
    const int64_t i = order[a].i;
    if (sup[i]) continue;
    keep[nk++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = area[i];
    for (int64_t b = a + 1; b < M; b++) {
      const int64_t j = order[b].i;
      if (sup[j]) continue;
      if (class_aware && idxs[i] != idxs[j]) continue;
      const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
      const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
      const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
      const float inter = w * h;
      const float ovr = inter / (iarea + area[j] - inter);
      if ((double)ovr > iou_threshold) sup[j] = 1;
    }
  }
  free(order);
  free(area);
  free(sup);
  free(boxes);
  return nk;
}

/*
 * cddmsl_b200 — C ABI of the B200-native region-level vision-language hot path of CDDMSL.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point enqueues work on the given CUDA stream and returns immediately (no hidden device synchronise);
 * all buffers, including workspaces, are owned by the caller (PyTorch's caching allocator in practice) and
 * outputs are fully overwritten.  All floating tensors are fp32, contiguous; indices are int64.
 *
 * Return value: 0 on success; otherwise a negative CDDMSL_E* shape/argument code or a positive
 * cudaError_t.  `cddmsl_error_string()` renders either.  The Python host layer raises RuntimeError.
 *
 * Reference interfaces replaced (paths relative to the upstream checkout):
 *   cddmsl_roi_align_fwd / _bwd   torchvision::roi_align / ::_roi_align_backward as called from
 *                                 detectron2/layers/roi_align.py:49-65 (autograd backward of the same op)
 *   cddmsl_nms                    torchvision.ops.boxes.batched_nms / torchvision::nms as called from
 *                                 detectron2/layers/nms.py:19-39
 *   cddmsl_clip_head_*            detectron2/modeling/roi_heads/fast_rcnn.py:543-565 (cosine logits),
 *                                 :624-644 + layers/wrappers.py:26-33 (focal / weighted CE), :100-127 (stats)
 *   cddmsl_align_loss_*           detectron2/modeling/meta_arch/rcnn.py:305-317, :455-468 (+ gather.py:5-20)
 */
#ifndef CDDMSL_B200_H_
#define CDDMSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cddmsl_stream_t; /* a cudaStream_t */

enum {
  CDDMSL_OK = 0,
  CDDMSL_EINVAL = -1,     /* bad shape / null pointer / unsupported size */
  CDDMSL_EWORKSPACE = -2, /* workspace too small */
  CDDMSL_EALIGN = -3      /* pointer not aligned as required */
};

int cddmsl_abi_version(void);
const char* cddmsl_error_string(int code);
/* number of kernels this library has launched in this process (all entry points); bench.py's
 * `gpu_launches` is the difference of two reads. */
uint64_t cddmsl_launch_count(void);

/* ---------------------------------------------------------------- piece 1: ROIAlign ------------- */
/* in [N,C,H,W], rois [R,5] = (batch_idx, x0, y0, x1, y1) in image coordinates, out [R,C,PH,PW]. */
/* workspace: scratch for the channels-last copy of the map used by the 14x14 fast path; with a NULL / too
 * small workspace the generic NCHW kernel runs instead (same results, slower). */
size_t cddmsl_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int R);
int cddmsl_roi_align_fwd(const float* in, const float* rois, float* out, int N, int C, int H, int W, int R,
                         int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, void* workspace,
                         size_t workspace_bytes, cddmsl_stream_t stream);

/* gout [R,C,PH,PW] -> gin [N,C,H,W] (fully overwritten; zeroing happens inside). */
size_t cddmsl_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R);
int cddmsl_roi_align_bwd(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R,
                         int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, void* workspace,
                         size_t workspace_bytes, cddmsl_stream_t stream);

/* The same RoIs on TWO feature maps of the same shape: the source / target pair of the region-level consistency
 * branch, detectron2/modeling/roi_heads/clip_roi_heads.py:117-132 (`forward_get_features` pools features_src and
 * features_trgt with identical proposal boxes; the reference runs ROIAlign twice).  One plan (per-RoI bands, image
 * buckets) serves both maps; results are bit-identical to two single calls.  Workspace: the single-map sizes. */
int cddmsl_roi_align_fwd2(const float* in_a, const float* in_b, const float* rois, float* out_a, float* out_b, int N,
                          int C, int H, int W, int R, int PH, int PW, float spatial_scale, int sampling_ratio,
                          int aligned, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);
int cddmsl_roi_align_bwd2(const float* gout_a, const float* gout_b, const float* rois, float* gin_a, float* gin_b,
                          int N, int C, int H, int W, int R, int PH, int PW, float spatial_scale, int sampling_ratio,
                          int aligned, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);

/* ---------------------------------------------------------------- piece 2: NMS ------------------ */
/* boxes [M,4] xyxy, scores [M], idxs [M] class/level ids or NULL.  keep [M] receives the kept ORIGINAL
 * indices ordered by score descending (ties: lower index first); *num_keep (device int32) their count.
 * iou_threshold is a double and the fp32 IoU is promoted before the strict `>` test, like the CPU kernel
 * the oracle runs.  coord_trick != 0 reproduces torchvision's `_batched_nms_coordinate_trick`
 * (boxes + idx * (max_coord + 1) in fp32), 0 the per-class loop (`_batched_nms_vanilla`, nms.py:32-39). */
size_t cddmsl_nms_workspace_bytes(int64_t M);
int cddmsl_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t M, double iou_threshold,
               int coord_trick, int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
               cddmsl_stream_t stream);
/* cddmsl_nms with an upper bound on the wanted keep list (`keep[:topk_per_image]`, fast_rcnn.py:186-187): the first
 * max_keep entries of the full result, bit-identical; see cddmsl_nms_batched_topk.  max_keep <= 0: everything. */
int cddmsl_nms_topk(const float* boxes, const float* scores, const int64_t* idxs, int64_t M, double iou_threshold,
                    int coord_trick, int max_keep, int presorted, int64_t* keep, int32_t* num_keep, void* workspace,
                    size_t workspace_bytes, cddmsl_stream_t stream);


/* The B images of one RPN batch in one call (the loop of proposal_utils.py:42-66 calls batched_nms once per image).
 * Padded layout: boxes [B][Mmax][4], scores [B][Mmax], idxs [B][Mmax] (nullable), counts int32[B] ON THE DEVICE
 * (image b holds counts[b] <= Mmax boxes; the rest of its slice is ignored), keep int64[B][Mmax] (indices local to
 * the image, score order), num_keep int32[B].  Result per image is bit-identical to cddmsl_nms on that image's
 * first counts[b] boxes; the coordinate-trick offset uses that image's own max coordinate (boxes.py:67-73).
 * No host synchronisation; the B greedy scans run concurrently. */
size_t cddmsl_nms_batched_workspace_bytes(int B, int64_t Mmax);
int cddmsl_nms_batched(const float* boxes, const float* scores, const int64_t* idxs, const int32_t* counts, int B,
                       int64_t Mmax, double iou_threshold, int coord_trick, int64_t* keep, int32_t* num_keep,
                       void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);
/* The same with an upper bound on the wanted keep list: only the first `max_keep` kept boxes of every image are
 * produced (num_keep[b] <= max_keep; max_keep <= 0: all of them).  This is what find_top_rpn_proposals needs
 * (`keep = keep[:post_nms_topk]`, detectron2/modeling/proposal_generator/proposal_utils.py:116-118): greedy NMS never
 * looks ahead, so the first max_keep entries depend on the best-scoring boxes only -- a first pass works on the top
 * 2*max_keep candidates, the full pass runs only for images where that was not enough (decided on the device).
 * Results are bit-identical to the first max_keep entries of cddmsl_nms_batched.  Same workspace.
 * presorted != 0: the caller guarantees that the scores of every image are already non-increasing (the RPN path hands
 * over its sorted top-k, proposal_utils.py:77-79); the stable sort is then the identity and is skipped. */
int cddmsl_nms_batched_topk(const float* boxes, const float* scores, const int64_t* idxs, const int32_t* counts, int B,
                            int64_t Mmax, double iou_threshold, int coord_trick, int max_keep, int presorted,
                            int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                            cddmsl_stream_t stream);


/* Proposal pre-processing between the RPN head and NMS for a batch: RPN._decode_proposals
 * (proposal_generator/rpn.py:514-533) = Box2BoxTransform.apply_deltas (box_regression.py:77-117) on the pre-NMS
 * top-k candidates of every image, then the finite check, Boxes.clip, Boxes.nonempty and the boolean selection of
 * proposal_utils.py:95-114 as a stable compaction.  anchors [A,4], deltas [N,A,4], topk_idx int64 [N,K] (anchor
 * indices in descending score order), topk_scores [N,K], image_hw [N,2] = (h, w).  Out: boxes_out [N,K,4] /
 * scores_out [N,K] with the survivors first, in score order; counts int32 [N]; *all_finite = 0 if any decoded box or
 * score of the batch is Inf/NaN (the reference raises FloatingPointError in training).  Feeds cddmsl_nms_batched. */
int cddmsl_rpn_decode_topk(const float* anchors, const float* deltas, const int64_t* topk_idx,
                           const float* topk_scores, const float* image_hw, int N, int64_t A, int K, float wx,
                           float wy, float ww, float wh, float scale_clamp, float min_box_size, float* boxes_out,
                           float* scores_out, int32_t* counts, int32_t* all_finite, cddmsl_stream_t stream);

/* ---------------------------------------------------------------- piece 3: CLIP box predictor --- */
/* loss modes */
enum { CDDMSL_LOSS_FOCAL = 0, CDDMSL_LOSS_CE = 1, CDDMSL_LOSS_WEIGHTED_CE = 2 };

/* scores[R,K+1] = [ x^ . w^_k | x^ . w_bg ] / T   (x^ = x / max(|x|, 1e-12), w^ likewise).
 * w [K,D] raw concept embeddings, w_bg [D] raw background embedding. */
size_t cddmsl_clip_head_workspace_bytes(int R, int D, int K);
int cddmsl_clip_head_scores(const float* x, const float* w, const float* w_bg, int R, int D, int K,
                            float temperature, float* scores, void* workspace, size_t workspace_bytes,
                            cddmsl_stream_t stream);
/* dscores [R,K+1] -> dx [R,D] through the cosine logits. */
int cddmsl_clip_head_scores_bwd(const float* x, const float* w, const float* w_bg, const float* dscores, int R,
                                int D, int K, float temperature, float* dx, void* workspace,
                                size_t workspace_bytes, cddmsl_stream_t stream);
/* Fused logits -> softmax -> (focal | CE | weighted CE) loss -> gradient.
 *   gt [R] int64 in [0,K] (K = background);  gamma: focal exponent (mode FOCAL);  bg_weight: weight of
 *   class K (< 0: no class weighting);  grad_scale: device pointer to the upstream dL/dloss (NULL = 1).
 *   scores (nullable) [R,K+1];  loss: device float;  dx (nullable) [R,D];  stats (nullable) device
 *   int32[4] = {num_accurate, num_fg, fg_num_accurate, num_false_negative} (fast_rcnn.py:100-127).
 *   strict_nan != 0 reproduces autograd's NaN gradient for rows whose softmax saturates to p_t == 1
 *   under the focal loss; 0 emits the analytic limit (0). */
int cddmsl_clip_head_loss(const float* x, const float* w, const float* w_bg, const int64_t* gt, int R, int D,
                          int K, float temperature, int loss_mode, float gamma, float bg_weight,
                          const float* grad_scale, int strict_nan, float* scores, float* loss, float* dx,
                          int32_t* stats, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);

/* Box-regression loss of the same predictor (fast_rcnn.py:646-689): smooth-L1 (fvcore.nn.smooth_l1_loss, beta)
 * between pred_deltas and Box2BoxTransform.get_deltas(proposal, gt) (box_regression.py:42-75, weights wx..wh),
 * summed over foreground rows (0 <= gt_classes < K) and divided by max(R, 1).  pred_deltas is [R,4]
 * (cls_agnostic != 0) or [R,4K]; dpred (nullable) receives d loss / d pred_deltas * (*grad_scale).  Unlike the
 * reference there is no device->host sync (no nonzero(), no validity assert): rows outside the foreground are never
 * evaluated, exactly as upstream never passes them to get_deltas. */
size_t cddmsl_box_reg_loss_workspace_bytes(int R);
int cddmsl_box_reg_loss(const float* proposal_boxes, const float* gt_boxes, const float* pred_deltas,
                        const int64_t* gt_classes, int R, int K, int cls_agnostic, float wx, float wy, float ww,
                        float wh, float beta, const float* grad_scale, float* loss, float* dpred, void* workspace,
                        size_t workspace_bytes, cddmsl_stream_t stream);

/* ---------------------------------------------------------------- next row: proposal matching -- */
/* pairwise_iou (detectron2/structures/boxes.py:346-368) + Matcher.__call__ / set_low_quality_matches_
 * (detectron2/modeling/matcher.py:63-127) for the B images of a batch in one pass, as label_and_sample_proposals
 * (roi_heads.py:285-289) and the RPN's anchor labelling use them; the [G, M] IoU matrix is never materialised.
 * Padded layout: gt_boxes [B][Gmax][4] with gt_counts int32[B] on the device, boxes [B][Mmax][4] with counts int32[B]
 * on the device (nullable: every image holds Mmax boxes).  thresholds (host, ascending, first > 0) and labels (host,
 * num_thresholds + 1 values in {-1,0,1}) are the Matcher's constructor arguments.  Outputs per candidate: matches
 * int64 (argmax gt, first maximum; 0 when the image has no gt), match_labels int8, matched_vals float (nullable).
 * Bit-identical to the reference arithmetic (unfused fp32 IoU).  workspace: only with allow_low_quality_matches. */
size_t cddmsl_match_boxes_workspace_bytes(int B, int Gmax);
int cddmsl_match_boxes(const float* gt_boxes, const int32_t* gt_counts, const float* boxes, const int32_t* counts,
                       int B, int Gmax, int Mmax, const float* thresholds, const int32_t* labels, int num_thresholds,
                       int allow_low_quality_matches, int64_t* matches, int8_t* match_labels, float* matched_vals,
                       void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);

/* ---------------------------------------------------------------- piece 4: alignment loss ------- */
/* Row-normalise src|tgt [n_local,D] each (x / |x|, no eps — rcnn.py:308-309, :458-459) and pack them as
 * packed[2][n_local][D] for ONE all-gather per branch (the reference issues two, rcnn.py:455-456;
 * normalisation is row-local so it commutes with the gather).  norms[2][n_local] keeps |row| for the
 * backward. */
int cddmsl_align_pack_normalized(const float* src, const float* tgt, int n_local, int D, float* packed,
                                 float* norms, cddmsl_stream_t stream);
/* packed_all [world][2][n_local][D]: what all_gather makes of every rank's `packed` (world = 1: the local
 * buffer itself).  loss = (CE(S, I) + CE(S^T, I)) / 2 with S = A^ B^T over all n = world*n_local rows, no
 * temperature.  Gradients w.r.t. the UN-normalised local rows (rank*n_local ...) only — GatherLayer.backward
 * keeps the local slice, gather.py:16-20 — into da, db [n_local,D] (both nullable).  grad_scale: device
 * pointer to the upstream dL/dloss (NULL = 1). */
size_t cddmsl_align_loss_workspace_bytes(int world, int n_local, int D);
int cddmsl_align_loss(const float* packed_all, const float* norms_local, int world, int n_local, int D, int rank,
                      const float* grad_scale, float* loss, float* da, float* db, void* workspace,
                      size_t workspace_bytes, cddmsl_stream_t stream);

/* cddmsl_align_loss with a temperature and the gradient semantics of RegionCLIP's `gather_tensors`
 * (detectron2/utils/comm.py:268-322, diffdist all_gather: every rank's identical loss back-propagates into every
 * rank's rows, so the local rows receive `world` times the gradient GatherLayer keeps): the image-text matching loss
 * of the pretraining model, detectron2/modeling/meta_arch/clip_rcnn.py:608-640 --
 * logits = logit_scale * A^ B^T (logit_scale = 1 / matching_temp), loss = (CE(S, I) + CE(S^T, I)) / 2, gradients of
 * the local rows multiplied by grad_mult.  (logit_scale = grad_mult = 1: cddmsl_align_loss.)  Same workspace. */
int cddmsl_contrastive_loss(const float* packed_all, const float* norms_local, int world, int n_local, int D, int rank,
                            float logit_scale, float grad_mult, const float* grad_scale, float* loss, float* da,
                            float* db, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);

/* Row-softmax losses against a dense target matrix, RegionCLIP pretraining (`region_concept_matching`,
 * clip_rcnn.py:583-606): mode 0 = KL distillation `F.kl_div(softmax(s).log(), t, 'batchmean')` (:597-600),
 * mode 1 = MILCrossEntropy(s, label_mtx) (detectron2/utils/comm.py:332-355, sum over positives, mean over rows).
 * logits [R, ld] (first K columns count), target [R, K]; dlogits (nullable) [R, ld] = d loss / d logits * *grad_scale
 * (device pointer, NULL = 1), zero in the columns >= K. */
size_t cddmsl_softmax_target_loss_workspace_bytes(int R);
int cddmsl_softmax_target_loss(const float* logits, int ld, const float* target, int R, int K, int mode,
                               const float* grad_scale, float* loss, float* dlogits, void* workspace,
                               size_t workspace_bytes, cddmsl_stream_t stream);

/* KD regulariser of the image-level branch, detectron2/modeling/meta_arch/rcnn.py:265-272:
 * loss = L1Loss()(teacher.detach(), student) = mean |teacher - student| over `numel` elements ([B,768] V2L
 * features); dstudent (nullable, [numel]) = sign(student - teacher) * scale / numel with scale = *grad_scale
 * (device pointer, NULL = 1) -- the gradient autograd would produce, from the same pass.  Deterministic. */
size_t cddmsl_kd_l1_loss_workspace_bytes(void);
int cddmsl_kd_l1_loss(const float* teacher, const float* student, int64_t numel, const float* grad_scale, float* loss,
                      float* dstudent, void* workspace, size_t workspace_bytes, cddmsl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CDDMSL_B200_H_ */

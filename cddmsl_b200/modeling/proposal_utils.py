"""`find_top_rpn_proposals` (detectron2/modeling/proposal_generator/proposal_utils.py:22-130): the RPN call
site of `batched_nms`.  Per feature level take the pre-NMS top-k by objectness, then per image: finite
check, clip, drop empty boxes, level-aware NMS, post-NMS top-k.  Same control flow as upstream (including
the FloatingPointError in training); only the NMS underneath is the sm_100a kernel.

On CUDA inputs the per-image loop is evaluated for all images at once (`_per_image_batched`): masks instead of
data-dependent indexing, a stable partition instead of boolean selection, ONE batched NMS call and ONE device->host
read for the whole batch (upstream: three syncs and one NMS per image).  Results are identical to the loop."""
from __future__ import annotations

from typing import List, Tuple

import torch

from ..layers import batched_nms, batched_nms_images, cat
from .. import ops
from ..structures import Boxes, Instances

# The container types the results are built with.  When `find_top_rpn_proposals` is patched into the reference's
# loop, set these to the reference's own classes -- its `add_ground_truth_to_proposals` calls
# `Instances.cat` / `Boxes.cat` on what this function returns:
#     pu.CONTAINERS.update(Instances=detectron2.structures.Instances, Boxes=detectron2.structures.Boxes)
CONTAINERS = {"Instances": Instances, "Boxes": Boxes}


BATCHED_IMAGES = True  # False: the upstream-shaped per-image loop (one NMS launch sequence + syncs per image)


def _per_image_batched(topk_proposals, topk_scores, level_ids, image_sizes, nms_thresh, post_nms_topk, min_box_size,
                       training):
    """proposal_utils.py:42-66 for every image at once.  topk_proposals [N,M,4], topk_scores [N,M], level_ids [M]."""
    n_img, m = topk_scores.shape
    device = topk_scores.device
    boxes = topk_proposals.float()
    valid = torch.isfinite(boxes).all(dim=2) & torch.isfinite(topk_scores)
    hw = torch.tensor([[float(h), float(w)] for h, w in image_sizes], device=device)        # Boxes.clip, boxes.py:192-206
    h, w = hw[:, 0:1], hw[:, 1:2]
    zero = torch.zeros((), device=device)
    x1 = torch.minimum(torch.maximum(boxes[..., 0], zero), w)
    y1 = torch.minimum(torch.maximum(boxes[..., 1], zero), h)
    x2 = torch.minimum(torch.maximum(boxes[..., 2], zero), w)
    y2 = torch.minimum(torch.maximum(boxes[..., 3], zero), h)
    boxes = torch.stack((x1, y1, x2, y2), dim=-1)
    nonempty = ((x2 - x1) > min_box_size) & ((y2 - y1) > min_box_size)                      # Boxes.nonempty, boxes.py:208-222
    sel = valid & nonempty
    # boolean selection == stable partition (selected rows first, original order) + a count
    perm = torch.argsort((~sel).to(torch.int8), dim=1, stable=True)
    boxes = torch.gather(boxes, 1, perm.unsqueeze(-1).expand(-1, -1, 4)).contiguous()
    scores = torch.gather(topk_scores, 1, perm)
    lvl = level_ids.to(torch.int64)[perm]
    counts = sel.sum(dim=1).to(torch.int32)
    keep, num_keep = batched_nms_images(boxes, scores, lvl, counts, nms_thresh, max_keep=post_nms_topk)
    host = torch.cat([num_keep, valid.all().reshape(1).to(torch.int32)]).tolist()          # the one sync of the batch
    if training and not host[-1]:
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    results: List[Instances] = []
    for n, image_size in enumerate(image_sizes):
        k = keep[n, : min(int(host[n]), post_nms_topk)]
        res = CONTAINERS["Instances"](image_size)
        res.proposal_boxes = CONTAINERS["Boxes"](boxes[n][k])
        res.objectness_logits = scores[n][k]
        results.append(res)
    return results


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                           post_nms_topk: int, min_box_size: float, training: bool):
    """
    Args:
        proposals: L tensors [N, Hi*Wi*A, 4];  pred_objectness_logits: L tensors [N, Hi*Wi*A]
        image_sizes: N (h, w) pairs
    Returns: list of N Instances with `proposal_boxes`, `objectness_logits` (sorted by score).
    """
    num_images = len(image_sizes)
    device = proposals[0].device
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images, device=device)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        hi_wi_a = logits_i.shape[1]
        num_proposals_i = min(hi_wi_a, pre_nms_topk)
        logits_i, idx = logits_i.sort(descending=True, dim=1)
        topk_scores_i = logits_i.narrow(1, 0, num_proposals_i)
        topk_idx = idx.narrow(1, 0, num_proposals_i)
        topk_proposals.append(proposals_i[batch_idx[:, None], topk_idx])
        topk_scores.append(topk_scores_i)
        level_ids.append(torch.full((num_proposals_i,), level_id, dtype=torch.int64, device=device))
    topk_scores = cat(topk_scores, dim=1)
    topk_proposals = cat(topk_proposals, dim=1)
    level_ids = cat(level_ids, dim=0)

    if BATCHED_IMAGES and device.type == "cuda" and num_images > 0:
        return _per_image_batched(topk_proposals, topk_scores, level_ids, image_sizes, nms_thresh, post_nms_topk,
                                  min_box_size, training)
    results: List[Instances] = []
    for n, image_size in enumerate(image_sizes):
        boxes = CONTAINERS["Boxes"](topk_proposals[n])
        scores_per_img = topk_scores[n]
        lvl = level_ids
        valid_mask = torch.isfinite(boxes.tensor).all(dim=1) & torch.isfinite(scores_per_img)
        if not valid_mask.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes, scores_per_img, lvl = boxes[valid_mask], scores_per_img[valid_mask], lvl[valid_mask]
        boxes.clip(image_size)
        keep = boxes.nonempty(threshold=min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores_per_img, lvl = boxes[keep], scores_per_img[keep], lvl[keep]
        keep = batched_nms(boxes.tensor, scores_per_img, lvl, nms_thresh)
        keep = keep[:post_nms_topk]  # already sorted by score
        res = CONTAINERS["Instances"](image_size)
        res.proposal_boxes = boxes[keep]
        res.objectness_logits = scores_per_img[keep]
        results.append(res)
    return results


def decode_proposals(anchors: List, pred_anchor_deltas: List[torch.Tensor], box2box_transform) -> List[torch.Tensor]:
    """`RPN._decode_proposals` (proposal_generator/rpn.py:514-533): anchors (one `Boxes` per level, shared by the
    images) + predicted deltas [N, Hi*Wi*A, B] -> proposals [N, Hi*Wi*A, B] per level."""
    n = pred_anchor_deltas[0].shape[0]
    proposals = []
    for anchors_i, deltas_i in zip(anchors, pred_anchor_deltas):
        a = anchors_i.tensor if hasattr(anchors_i, "tensor") else anchors_i
        b = a.size(1)
        deltas_i = deltas_i.reshape(-1, b)
        a = a.unsqueeze(0).expand(n, -1, -1).reshape(-1, b)
        proposals.append(box2box_transform.apply_deltas(deltas_i, a).view(n, -1, b))
    return proposals


@torch.no_grad()
def predict_proposals(anchors: List, pred_objectness_logits: List[torch.Tensor], pred_anchor_deltas: List[torch.Tensor],
                      image_sizes: List[Tuple[int, int]], *, box2box_transform, nms_thresh: float, pre_nms_topk: int,
                      post_nms_topk: int, min_box_size: float, training: bool):
    """`RPN.predict_proposals` (proposal_generator/rpn.py:482-512): decode, top-k, clip, drop empty boxes, NMS,
    post-NMS top-k; a list of N Instances (`proposal_boxes`, `objectness_logits`, score-descending).

    One RPN level on CUDA (the C4 detector of both CDDMSL configs): `sort` (the segmented top-k) -> ONE kernel for
    decode + finite flag + clip + non-empty + stable compaction of the top-k candidates of every image
    (csrc/rpn_decode.cu; only the kept candidates are decoded) -> ONE batched NMS call -> ONE device->host read
    (kept counts + finite flag).  Several levels / CPU tensors: the reference-shaped path."""
    single = len(pred_objectness_logits) == 1 and pred_objectness_logits[0].is_cuda and BATCHED_IMAGES
    if not single or len(image_sizes) == 0 or len(box2box_transform.weights) != 4:
        props = decode_proposals(anchors, pred_anchor_deltas, box2box_transform)
        return find_top_rpn_proposals(props, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk,
                                      post_nms_topk, min_box_size, training)
    logits = pred_objectness_logits[0]
    n, a_tot = logits.shape
    k = min(a_tot, pre_nms_topk)
    sorted_logits, idx = logits.sort(descending=True, dim=1)        # proposal_utils.py:77-79
    topk_scores, topk_idx = sorted_logits.narrow(1, 0, k), idx.narrow(1, 0, k)
    an = anchors[0].tensor if hasattr(anchors[0], "tensor") else anchors[0]
    hw = torch.tensor([[float(h), float(w)] for h, w in image_sizes], device=logits.device)
    boxes, scores, counts, fin = ops.rpn_decode_topk(an, pred_anchor_deltas[0].reshape(n, a_tot, 4), topk_idx,
                                                     topk_scores, hw, [float(v) for v in box2box_transform.weights],
                                                     float(box2box_transform.scale_clamp), float(min_box_size))
    # `topk_scores` are sorted (proposal_utils.py:77-79) and the decode kernel compacts stably: no second sort
    keep, num_keep = batched_nms_images(boxes, scores, None, counts, nms_thresh, max_keep=post_nms_topk, presorted=True)
    host = torch.cat([num_keep.to(torch.int32), fin]).tolist()       # the one sync of the batch
    if training and not host[-1]:
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    results = []
    for i, image_size in enumerate(image_sizes):
        kk = keep[i, : min(int(host[i]), post_nms_topk)]
        res = CONTAINERS["Instances"](image_size)
        res.proposal_boxes = CONTAINERS["Boxes"](boxes[i][kk])
        res.objectness_logits = scores[i][kk]
        results.append(res)
    return results

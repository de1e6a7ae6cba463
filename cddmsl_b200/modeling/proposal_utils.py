"""`find_top_rpn_proposals` (detectron2/modeling/proposal_generator/proposal_utils.py:22-130): the RPN call
site of `batched_nms`.  Per feature level take the pre-NMS top-k by objectness, then per image: finite
check, clip, drop empty boxes, level-aware NMS, post-NMS top-k.  Same control flow as upstream (including
the FloatingPointError in training); only the NMS underneath is the sm_100a kernel."""
from __future__ import annotations

from typing import List, Tuple

import torch

from ..layers import batched_nms, cat
from ..structures import Boxes, Instances


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

    results: List[Instances] = []
    for n, image_size in enumerate(image_sizes):
        boxes = Boxes(topk_proposals[n])
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
        res = Instances(image_size)
        res.proposal_boxes = boxes[keep]
        res.objectness_logits = scores_per_img[keep]
        results.append(res)
    return results

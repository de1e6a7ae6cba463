"""`Matcher` (detectron2/modeling/matcher.py:8-127) and `pairwise_iou` (detectron2/structures/boxes.py:346-368).

`Matcher.__call__(match_quality_matrix)` keeps the reference's signature and arithmetic (plain tensor ops, any
device).  `Matcher.match_boxes(gt_boxes, gt_counts, boxes, counts)` is the fused form the training path uses: IoU and
matching of all images of a batch in one kernel (csrc/match.cu), the [G, M] matrix never materialised."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import ops
from ..layers import nonzero_tuple


def pairwise_iou(boxes1, boxes2) -> torch.Tensor:
    """boxes.py:346-368 on `Boxes` or plain [N,4] / [M,4] tensors -> [N,M]."""
    b1 = boxes1.tensor if hasattr(boxes1, "tensor") else boxes1
    b2 = boxes2.tensor if hasattr(boxes2, "tensor") else boxes2
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (area1[:, None] + area2 - inter),
                       torch.zeros(1, dtype=inter.dtype, device=inter.device))


class Matcher:
    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        thresholds = thresholds[:]
        assert thresholds[0] > 0
        self._user_thresholds = thresholds[:]
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = labels
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert match_quality_matrix.dim() == 2
        if match_quality_matrix.numel() == 0:
            n = match_quality_matrix.size(1)
            return (match_quality_matrix.new_full((n,), 0, dtype=torch.int64),
                    match_quality_matrix.new_full((n,), self.labels[0], dtype=torch.int8))
        assert torch.all(match_quality_matrix >= 0)
        matched_vals, matches = match_quality_matrix.max(dim=0)
        match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for (l, low, high) in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            match_labels[(matched_vals >= low) & (matched_vals < high)] = l
        if self.allow_low_quality_matches:
            self.set_low_quality_matches_(match_labels, match_quality_matrix)
        return matches, match_labels

    def set_low_quality_matches_(self, match_labels, match_quality_matrix):
        highest_quality_foreach_gt, _ = match_quality_matrix.max(dim=1)
        _, pred_inds = nonzero_tuple(match_quality_matrix == highest_quality_foreach_gt[:, None])
        match_labels[pred_inds] = 1

    def match_boxes(self, gt_boxes: torch.Tensor, gt_counts: torch.Tensor, boxes: torch.Tensor,
                    counts: Optional[torch.Tensor] = None):
        """Fused `pairwise_iou` + `__call__` for a padded batch: gt_boxes [B,G,4] (+ gt_counts [B]), boxes [B,M,4]
        (+ counts [B] or None).  Returns (matches int64 [B,M], match_labels int8 [B,M], matched_vals [B,M]); rows
        beyond an image's count are zero.  No device->host sync, no validity assert on the IoUs."""
        return ops.match_boxes(gt_boxes, gt_counts, boxes, counts, [float(t) for t in self._user_thresholds],
                               [int(l) for l in self.labels], bool(self.allow_low_quality_matches))

"""The proposal labelling / sampling step that feeds ROIAlign during training: `ROIHeads._sample_proposals` and
`ROIHeads.label_and_sample_proposals` (detectron2/modeling/roi_heads/roi_heads.py:196-319) and
`add_ground_truth_to_proposals` (proposal_generator/proposal_utils.py:133-200).

SURVEY.md §8f row 3.  The reference loops over the images in Python: per image an IoU matrix, two reductions, two
`nonzero`, two `randperm`, a `.sum().item()` sync.  On CUDA inputs this mirror evaluates the whole batch at once:
one fused IoU+matcher kernel over a padded [B, M] layout (csrc/match.cu), mask-based sampling with random keys
(`sampling.subsample_labels_batched`), and ONE device->host read (the per-image sample counts) per batch.  The
deterministic part -- matched ground-truth index, label, assigned class -- is bit-identical to the reference; the
random subset follows the same law (uniform subset of the positives / negatives of the same sizes) but not the same
random stream (upstream draws `randperm` on the device for positives and on the CPU for negatives)."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from ..structures import Boxes, Instances
from . import fast_rcnn as _fr
from .matcher import Matcher, pairwise_iou
from .sampling import subsample_labels, subsample_labels_batched

BATCHED_IMAGES = True  # False: the upstream-shaped per-image loop


def add_ground_truth_to_proposals_single_image(gt, proposals: Instances) -> Instances:
    """proposal_utils.py:162-200: gt boxes join the proposals with objectness logit ~ +inf (P(object) = 1 - 1e-10)."""
    if isinstance(gt, Boxes):
        gt = Instances(proposals.image_size, gt_boxes=gt)
    gt_boxes = gt.gt_boxes
    device = proposals.objectness_logits.device
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
    out = Instances(proposals.image_size)
    out.proposal_boxes = Boxes.cat([proposals.proposal_boxes, gt_boxes])
    out.objectness_logits = torch.cat([proposals.objectness_logits, gt_logits])
    return out


def add_ground_truth_to_proposals(gt: List, proposals: List[Instances]) -> List[Instances]:
    assert gt is not None
    if len(proposals) != len(gt):
        raise ValueError("proposals and gt should have the same length as the number of images!")
    if len(proposals) == 0:
        return proposals
    return [add_ground_truth_to_proposals_single_image(g, p) for g, p in zip(gt, proposals)]


class ROIHeads:
    """The sampling half of `ROIHeads` (roi_heads.py:118-319); ctor arguments = the cfg keys of `from_config`
    (`ROI_HEADS.NUM_CLASSES`, `BATCH_SIZE_PER_IMAGE`, `POSITIVE_FRACTION`, `IOU_THRESHOLDS`, `IOU_LABELS`,
    `PROPOSAL_APPEND_GT`)."""

    def __init__(self, *, num_classes: int, batch_size_per_image: int = 512, positive_fraction: float = 0.25,
                 proposal_matcher: Matcher = None, proposal_append_gt: bool = True):
        self.num_classes = num_classes
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.proposal_matcher = proposal_matcher or Matcher([0.5], [0, 1], allow_low_quality_matches=False)
        self.proposal_append_gt = proposal_append_gt

    # ------------------------------------------------------------------ reference-shaped per-image pieces
    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes) -> Tuple[torch.Tensor, torch.Tensor]:
        """roi_heads.py:196-235."""
        has_gt = gt_classes.numel() > 0
        if has_gt:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes
            gt_classes[matched_labels == -1] = -1
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        fg, bg = subsample_labels(gt_classes, self.batch_size_per_image, self.positive_fraction, self.num_classes)
        sampled_idxs = torch.cat([fg, bg], dim=0)
        return sampled_idxs, gt_classes[sampled_idxs]

    def _finish_image(self, proposals_per_image, targets_per_image, sampled_idxs, gt_classes, matched_idxs):
        out = proposals_per_image[sampled_idxs]
        out.gt_classes = gt_classes
        if len(targets_per_image) > 0:
            sampled_targets = matched_idxs[sampled_idxs]
            for name, value in targets_per_image.get_fields().items():
                if name.startswith("gt_") and not out.has(name):
                    out.set(name, value[sampled_targets])
        return out

    # ------------------------------------------------------------------ the public step
    @torch.no_grad()
    def label_and_sample_proposals(self, proposals: List[Instances], targets: List[Instances]) -> List[Instances]:
        if self.proposal_append_gt:
            proposals = add_ground_truth_to_proposals(targets, proposals)
        on_cuda = len(proposals) > 0 and proposals[0].proposal_boxes.tensor.is_cuda
        if BATCHED_IMAGES and on_cuda:
            return self._label_and_sample_batched(proposals, targets)
        out, num_fg, num_bg = [], [], []
        for p, t in zip(proposals, targets):
            mqm = pairwise_iou(t.gt_boxes, p.proposal_boxes)
            matched_idxs, matched_labels = self.proposal_matcher(mqm)
            sampled_idxs, gt_classes = self._sample_proposals(matched_idxs, matched_labels, t.gt_classes)
            out.append(self._finish_image(p, t, sampled_idxs, gt_classes, matched_idxs))
            num_bg.append((gt_classes == self.num_classes).sum().item())
            num_fg.append(gt_classes.numel() - num_bg[-1])
        if out:
            _fr._put_scalar("roi_head/num_fg_samples", sum(num_fg) / len(num_fg))
            _fr._put_scalar("roi_head/num_bg_samples", sum(num_bg) / len(num_bg))
        return out

    def _label_and_sample_batched(self, proposals: List[Instances], targets: List[Instances]) -> List[Instances]:
        nb = len(proposals)
        dev = proposals[0].proposal_boxes.tensor.device
        m_len = [len(p) for p in proposals]                     # host-known: Instances carry their length
        g_len = [len(t) for t in targets]
        m_max, g_max = max(m_len), max(max(g_len), 1)
        boxes = torch.zeros((nb, m_max, 4), device=dev)
        gtb = torch.zeros((nb, g_max, 4), device=dev)
        gtc = torch.zeros((nb, g_max), dtype=torch.int64, device=dev)
        for b, (p, t) in enumerate(zip(proposals, targets)):
            boxes[b, : m_len[b]] = p.proposal_boxes.tensor
            if g_len[b]:
                gtb[b, : g_len[b]] = t.gt_boxes.tensor
                gtc[b, : g_len[b]] = t.gt_classes
        counts = torch.tensor(m_len, dtype=torch.int32).to(dev, non_blocking=True)
        gcounts = torch.tensor(g_len, dtype=torch.int32).to(dev, non_blocking=True)
        matches, mlabels, _ = self.proposal_matcher.match_boxes(gtb, gcounts, boxes, counts)
        # roi_heads.py:216-224: class of the matched gt, background for label 0, ignore for -1; no gt -> background
        cls = gtc.gather(1, matches)
        cls = torch.where(mlabels == 0, torch.full_like(cls, self.num_classes), cls)
        cls = torch.where(mlabels == -1, torch.full_like(cls, -1), cls)
        cls = torch.where((gcounts > 0)[:, None], cls, torch.full_like(cls, self.num_classes))
        valid = torch.arange(m_max, device=dev)[None, :] < counts[:, None]
        pos_idx, neg_idx, num_pos, num_neg = subsample_labels_batched(cls, valid, self.batch_size_per_image,
                                                                      self.positive_fraction, self.num_classes)
        host = torch.stack([num_pos, num_neg], dim=1).tolist()  # the one device->host read of the batch
        out = []
        for b, (p, t) in enumerate(zip(proposals, targets)):
            n_pos, n_neg = host[b]
            sampled = torch.cat([pos_idx[b, :n_pos], neg_idx[b, :n_neg]])
            out.append(self._finish_image(p, t, sampled, cls[b][sampled], matches[b]))
        if out:
            _fr._put_scalar("roi_head/num_fg_samples", sum(h[0] for h in host) / nb)
            _fr._put_scalar("roi_head/num_bg_samples", sum(h[1] for h in host) / nb)
        return out

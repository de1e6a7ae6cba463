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
from .. import ops as _ops
from . import fast_rcnn as _fr
from .matcher import Matcher, pairwise_iou
from .sampling import subsample_labels, subsample_labels_batched

BATCHED_IMAGES = True  # False: the upstream-shaped per-image loop


def add_ground_truth_to_proposals_single_image(gt, proposals):
    """proposal_utils.py:162-200: gt boxes join the proposals with objectness logit ~ +inf (P(object) = 1 - 1e-10).
    Results are built with the CALLER's container types (`type(proposals)`, `type(proposals.proposal_boxes)`), so the
    function also serves the reference's own `Instances` / `Boxes` when it is patched into the upstream loop."""
    inst_cls, box_cls = type(proposals), type(proposals.proposal_boxes)
    if not hasattr(gt, "gt_boxes"):       # a bare Boxes
        gt = inst_cls(proposals.image_size, gt_boxes=gt)
    gt_boxes = gt.gt_boxes
    device = proposals.objectness_logits.device
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
    out = inst_cls(proposals.image_size)
    out.proposal_boxes = box_cls.cat([proposals.proposal_boxes, gt_boxes])
    out.objectness_logits = torch.cat([proposals.objectness_logits, gt_logits])
    return out


def add_ground_truth_to_proposals(gt: List, proposals: List) -> List:
    assert gt is not None
    if len(proposals) != len(gt):
        raise ValueError("proposals and gt should have the same length as the number of images!")
    if len(proposals) == 0:
        return proposals
    return [add_ground_truth_to_proposals_single_image(g, p) for g, p in zip(gt, proposals)]


def _matcher_args(matcher):
    """(thresholds without the +-inf sentinels, labels, allow_low_quality_matches) of this package's Matcher or of
    the reference's (matcher.py:36-61 keeps `thresholds` with the sentinels inserted)."""
    thr = getattr(matcher, "_user_thresholds", None)
    if thr is None:
        thr = [t for t in matcher.thresholds if math.isfinite(t)]
    return [float(t) for t in thr], [int(l) for l in matcher.labels], bool(matcher.allow_low_quality_matches)


def finish_image(proposals_per_image, targets_per_image, sampled_idxs, gt_classes, matched_idxs):
    """roi_heads.py:289-307: the sampled proposals with their class and the gt_* fields of the matched targets."""
    out = proposals_per_image[sampled_idxs]
    out.gt_classes = gt_classes
    if len(targets_per_image) > 0:
        sampled_targets = matched_idxs[sampled_idxs]
        for name, value in targets_per_image.get_fields().items():
            if name.startswith("gt_") and not out.has(name):
                out.set(name, value[sampled_targets])
    return out


def sample_proposals(matched_idxs, matched_labels, gt_classes, *, num_classes: int, batch_size_per_image: int,
                     positive_fraction: float, only_sample_fg_proposals: bool = False):
    """roi_heads.py:196-235 for one image."""
    has_gt = gt_classes.numel() > 0
    if has_gt:
        gt_classes = gt_classes[matched_idxs]
        gt_classes[matched_labels == 0] = num_classes
        gt_classes[matched_labels == -1] = -1
    else:
        gt_classes = torch.zeros_like(matched_idxs) + num_classes
    if only_sample_fg_proposals:            # MODEL.CLIP.ONLY_SAMPLE_FG_PROPOSALS, roi_heads.py:216-228
        if has_gt:
            positive = ((gt_classes != -1) & (gt_classes != num_classes)).nonzero(as_tuple=True)[0]
            num_pos = min(positive.numel(), int(batch_size_per_image * positive_fraction))
            sampled_idxs = positive[torch.randperm(positive.numel(), device=positive.device)[:num_pos]]
        else:                               # no gt: one background proposal fills the slot
            sampled_idxs = torch.zeros_like(matched_idxs[0:1])
        return sampled_idxs, gt_classes[sampled_idxs]
    fg, bg = subsample_labels(gt_classes, batch_size_per_image, positive_fraction, num_classes)
    sampled_idxs = torch.cat([fg, bg], dim=0)
    return sampled_idxs, gt_classes[sampled_idxs]


def label_and_sample_batched(proposals: List, targets: List, *, matcher, num_classes: int, batch_size_per_image: int,
                             positive_fraction: float, only_sample_fg_proposals: bool = False) -> List:
    """All images of a batch at once on the device: fused IoU + matcher kernel over a padded [B, M] layout, masked
    sampling with random keys, ONE device->host read.  A free function of explicit arguments: it works on any
    `Instances` / `Boxes` / `Matcher` objects with the reference's interface."""
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
    thr, lab, low = _matcher_args(matcher)
    matches, mlabels, _ = _ops.match_boxes(gtb, gcounts, boxes, counts, thr, lab, low)
    # roi_heads.py:205-213: class of the matched gt, background for label 0, ignore for -1; no gt -> background
    cls = gtc.gather(1, matches)
    cls = torch.where(mlabels == 0, torch.full_like(cls, num_classes), cls)
    cls = torch.where(mlabels == -1, torch.full_like(cls, -1), cls)
    cls = torch.where((gcounts > 0)[:, None], cls, torch.full_like(cls, num_classes))
    valid = torch.arange(m_max, device=dev)[None, :] < counts[:, None]
    pos_idx, neg_idx, num_pos, num_neg = subsample_labels_batched(cls, valid, batch_size_per_image, positive_fraction,
                                                                  num_classes)
    if only_sample_fg_proposals:   # roi_heads.py:216-228: positives only; an image without gt keeps proposal 0
        num_neg = torch.where(gcounts > 0, torch.zeros_like(num_neg), torch.ones_like(num_neg))
        num_pos = torch.where(gcounts > 0, num_pos, torch.zeros_like(num_pos))
        neg_idx = torch.zeros_like(neg_idx[:, :1])
    host = torch.stack([num_pos, num_neg], dim=1).tolist()  # the one device->host read of the batch
    out = []
    for b, (p, t) in enumerate(zip(proposals, targets)):
        n_pos, n_neg = host[b]
        sampled = torch.cat([pos_idx[b, :n_pos], neg_idx[b, :n_neg]])
        out.append(finish_image(p, t, sampled, cls[b][sampled], matches[b]))
    if out:
        n_bg = [int((o.gt_classes == num_classes).sum()) for o in out] if only_sample_fg_proposals else \
            [h[1] for h in host]
        n_all = [len(o) for o in out]
        _fr._put_scalar("roi_head/num_fg_samples", sum(a - g for a, g in zip(n_all, n_bg)) / nb)
        _fr._put_scalar("roi_head/num_bg_samples", sum(n_bg) / nb)
    return out


@torch.no_grad()
def label_and_sample_proposals(self, proposals: List, targets: List) -> List:
    """`ROIHeads.label_and_sample_proposals` (roi_heads.py:236-319) as a function of `self`: it only reads the
    attributes the REFERENCE class has (`proposal_append_gt`, `proposal_matcher`, `num_classes`,
    `batch_size_per_image`, `positive_fraction`, `only_sample_fg_proposals`), so it can be assigned to the upstream
    class as well as to the mirror below."""
    if self.proposal_append_gt:
        proposals = add_ground_truth_to_proposals(targets, proposals)
    kw = dict(num_classes=self.num_classes, batch_size_per_image=self.batch_size_per_image,
              positive_fraction=self.positive_fraction,
              only_sample_fg_proposals=bool(getattr(self, "only_sample_fg_proposals", False)))
    on_cuda = len(proposals) > 0 and proposals[0].proposal_boxes.tensor.is_cuda
    if BATCHED_IMAGES and on_cuda:
        return label_and_sample_batched(proposals, targets, matcher=self.proposal_matcher, **kw)
    out, num_fg, num_bg = [], [], []
    for p, t in zip(proposals, targets):
        mqm = pairwise_iou(t.gt_boxes, p.proposal_boxes)
        matched_idxs, matched_labels = self.proposal_matcher(mqm)
        sampled_idxs, gt_classes = sample_proposals(matched_idxs, matched_labels, t.gt_classes, **kw)
        out.append(finish_image(p, t, sampled_idxs, gt_classes, matched_idxs))
        num_bg.append((gt_classes == self.num_classes).sum().item())
        num_fg.append(gt_classes.numel() - num_bg[-1])
    if out:
        _fr._put_scalar("roi_head/num_fg_samples", sum(num_fg) / len(num_fg))
        _fr._put_scalar("roi_head/num_bg_samples", sum(num_bg) / len(num_bg))
    return out


class ROIHeads:
    """The sampling half of `ROIHeads` (roi_heads.py:118-319); ctor arguments = the cfg keys of `from_config`
    (`ROI_HEADS.NUM_CLASSES`, `BATCH_SIZE_PER_IMAGE`, `POSITIVE_FRACTION`, `IOU_THRESHOLDS`, `IOU_LABELS`,
    `PROPOSAL_APPEND_GT`, `MODEL.CLIP.ONLY_SAMPLE_FG_PROPOSALS`)."""

    def __init__(self, *, num_classes: int, batch_size_per_image: int = 512, positive_fraction: float = 0.25,
                 proposal_matcher: Matcher = None, proposal_append_gt: bool = True,
                 only_sample_fg_proposals: bool = False):
        self.num_classes = num_classes
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.proposal_matcher = proposal_matcher or Matcher([0.5], [0, 1], allow_low_quality_matches=False)
        self.proposal_append_gt = proposal_append_gt
        self.only_sample_fg_proposals = only_sample_fg_proposals

    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes) -> Tuple[torch.Tensor, torch.Tensor]:
        """roi_heads.py:196-235."""
        return sample_proposals(matched_idxs, matched_labels, gt_classes, num_classes=self.num_classes,
                                batch_size_per_image=self.batch_size_per_image,
                                positive_fraction=self.positive_fraction,
                                only_sample_fg_proposals=self.only_sample_fg_proposals)

    label_and_sample_proposals = label_and_sample_proposals


class CLIPRes5ROIHeads(ROIHeads):
    """The ROI head both CDDMSL configs use (detectron2/modeling/roi_heads/clip_roi_heads.py:36-175, box branch):
    pooler -> the backbone's `layer4` (res5) -> attention pooling -> CLIP box predictor.  `res5` and `attnpool`
    belong to the backbone and are handed in per call exactly like upstream (rcnn.py:608-609); they stay PyTorch
    (cuDNN convolutions / MHA, out of scope).  Same constructor keywords as the reference minus the mask head."""

    def __init__(self, *, in_features, pooler, res5=None, box_predictor, **kwargs):
        super().__init__(**kwargs)
        self.in_features = in_features
        self.pooler = pooler
        self.res5 = res5            # None: this head uses the res5 of the backbone (clip_roi_heads.py:65)
        self.box_predictor = box_predictor
        self.mask_on = False
        self.training = True

    def train(self, mode: bool = True):
        self.training = mode
        for m in (self.pooler, self.box_predictor):
            if hasattr(m, "train"):
                m.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def _shared_roi_transform(self, features, boxes, backbone_res5):
        """clip_roi_heads.py:113-115"""
        x = self.pooler(features, boxes)
        return backbone_res5(x)

    def forward_get_features(self, features_src, features_trgt, proposals, targets=None, res5=None, attnpool=None):
        """clip_roi_heads.py:117-132 -- region embeddings of the source and the target image for the SAME proposal
        boxes (the region-level caption-consistency branch, rcnn.py:441-444).  The two ROIAligns of the reference
        run as one dual-map call (`ROIPooler.forward_pair`)."""
        if self.training:
            assert targets
        del targets
        proposal_boxes = [x.proposal_boxes for x in proposals]
        pooled_src, pooled_trgt = self.pooler.forward_pair([features_src[f] for f in self.in_features],
                                                           [features_trgt[f] for f in self.in_features],
                                                           proposal_boxes)
        box_features_src = res5(pooled_src)
        box_features_trgt = res5(pooled_trgt)
        if attnpool:
            att_feats_src = attnpool(box_features_src)
            att_feats_trgt = attnpool(box_features_trgt)
        return att_feats_src, att_feats_trgt   # (like upstream: undefined without attnpool)

    def forward(self, images, features, proposals, targets=None, res5=None, attnpool=None):
        """clip_roi_heads.py:134-175 (box branch).  Training: ([], losses); inference: (instances, {})."""
        del images
        if self.training:
            assert targets
            proposals = self.label_and_sample_proposals(proposals, targets)
        del targets
        proposal_boxes = [x.proposal_boxes for x in proposals]
        box_features = self._shared_roi_transform([features[f] for f in self.in_features], proposal_boxes, res5)
        if attnpool:
            predictions = self.box_predictor(attnpool(box_features))
        else:
            predictions = self.box_predictor(box_features.mean(dim=[2, 3]))
        if self.training:
            del features
            return [], self.box_predictor.losses(predictions, proposals)
        pred_instances, _ = self.box_predictor.inference(predictions, proposals)
        return pred_instances, {}

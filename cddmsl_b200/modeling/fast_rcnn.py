"""Box predictor with the reference's interface (`FastRCNNOutputLayers`,
detectron2/modeling/roi_heads/fast_rcnn.py:371-810) on top of the fused sm_100a head kernel.

What is kept identical: constructor keywords, module / parameter names (`cls_score`, `cls_bg_score`,
`test_cls_score`, `bbox_pred` — they are checkpoint keys), `forward(x) -> (scores, deltas)`,
`losses(predictions, proposals) -> {"loss_cls", "loss_box_reg"}`, `inference`, `predict_boxes`,
`predict_probs`, the EventStorage scalar names.

What is different underneath:
  * `forward` computes the CLIP logits with ONE kernel (normalise x, K+1 dot products, /T) instead of
    F.normalize x2 + 2 matmuls + cat + div (fast_rcnn.py:547-565);
  * `losses` recognises scores produced by its own `forward` and then evaluates logits -> softmax -> focal /
    CE loss -> statistics in ONE fused kernel straight from `x` (its backward is one more launch that writes
    dx), so the [R,K+1] logits never take part in autograd; hand-made scores fall back to the two-op path
    (our `clip_head_scores` autograd + PyTorch loss arithmetic);
  * the three `nonzero().numel()` host syncs of `_log_classification_stats` (:115-121) become one 16-byte
    read of counters the fused kernel already produced.
`bbox_pred` (a cuBLAS Linear) and `box_reg_loss` are not named by the north-star path and stay in PyTorch
(SURVEY.md §8f row 1).
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import nn
from torch.nn import functional as F

from .. import ops
from ..layers import batched_nms, cat, cross_entropy, nonzero_tuple
from ..structures import Boxes, Instances

_storage_hook = None  # callable(name, value) — e.g. detectron2's get_event_storage().put_scalar

# True: `box_reg_loss` selects the foreground rows with `nonzero` and asserts on degenerate boxes exactly like the
# reference (two device->host syncs per step); False (default): the same sum evaluated with a mask, no sync.
STRICT_BOX_REG_SYNC = False


def set_scalar_sink(fn) -> None:
    """Where `fast_rcnn/cls_accuracy` & co. go (the reference writes them to EventStorage, :123-127).
    With no sink registered the statistics are not read back at all (no device->host sync in `losses`)."""
    global _storage_hook
    _storage_hook = fn


def _put_scalar(name: str, value: float) -> None:
    if _storage_hook is not None:
        _storage_hook(name, value)


def smooth_l1_loss(input: torch.Tensor, target: torch.Tensor, beta: float, reduction: str = "none"):
    """fvcore.nn.smooth_l1_loss (used at fast_rcnn.py:671-673)."""
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        return loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
    return loss.sum() if reduction == "sum" else loss


def _log_classification_stats(pred_logits, gt_classes, prefix="fast_rcnn", counters: Optional[torch.Tensor] = None):
    """fast_rcnn.py:100-127.  `counters` = int32[4] from the fused kernel (accurate, fg, fg accurate,
    false negative); without it the reference arithmetic is evaluated in PyTorch."""
    num_instances = gt_classes.numel()
    if num_instances == 0 or _storage_hook is None:
        return
    if counters is not None:
        num_accurate, num_fg, fg_num_accurate, num_false_negative = counters.tolist()
    else:
        pred_classes = pred_logits.argmax(dim=1)
        bg_class_ind = pred_logits.shape[1] - 1
        fg_inds = (gt_classes >= 0) & (gt_classes < bg_class_ind)
        num_fg = int(fg_inds.sum())
        num_false_negative = int((pred_classes[fg_inds] == bg_class_ind).sum())
        num_accurate = int((pred_classes == gt_classes).sum())
        fg_num_accurate = int((pred_classes[fg_inds] == gt_classes[fg_inds]).sum())
    _put_scalar(f"{prefix}/cls_accuracy", num_accurate / num_instances)
    if num_fg > 0:
        _put_scalar(f"{prefix}/fg_cls_accuracy", fg_num_accurate / num_fg)
        _put_scalar(f"{prefix}/false_negative", num_false_negative / num_fg)


def fast_rcnn_inference(boxes, scores, image_shapes, score_thresh, nms_thresh, soft_nms_enabled=False,
                        soft_nms_method="gaussian", soft_nms_sigma=0.5, soft_nms_prune=0.001, topk_per_image=100,
                        scores_bf_multiply=None, vis=False):
    """fast_rcnn.py:42-98."""
    if scores_bf_multiply is None:
        scores_bf_multiply = scores
    result_per_image = [
        fast_rcnn_inference_single_image(b, s, shape, score_thresh, nms_thresh, soft_nms_enabled, soft_nms_method,
                                         soft_nms_sigma, soft_nms_prune, topk_per_image, sbf, vis)
        for s, b, shape, sbf in zip(scores, boxes, image_shapes, scores_bf_multiply)
    ]
    return [x[0] for x in result_per_image], [x[1] for x in result_per_image]


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, soft_nms_enabled=False,
                                     soft_nms_method="gaussian", soft_nms_sigma=0.5, soft_nms_prune=0.001,
                                     topk_per_image=100, scores_bf_multiply=None, vis=False):
    """fast_rcnn.py:130-209: drop the background column, clip, score threshold, class-aware NMS (:184),
    top-k.  Soft-NMS is an optional alternative upstream (default off, config/defaults.py:399) and is not
    part of this build."""
    if soft_nms_enabled:
        raise NotImplementedError("soft-NMS is outside the accelerated path (MODEL.ROI_HEADS.SOFT_NMS_ENABLED=False)")
    if scores_bf_multiply is None:
        scores_bf_multiply = scores
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes, scores, scores_bf_multiply = boxes[valid_mask], scores[valid_mask], scores_bf_multiply[valid_mask]
    scores = scores[:, :-1]
    scores_bf_multiply = scores_bf_multiply[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    bx = Boxes(boxes.reshape(-1, 4))
    bx.clip(image_shape)
    boxes = bx.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    scores_bf_multiply = scores_bf_multiply[filter_mask]
    # (only the first topk_per_image entries are used below: the kernels stop once they are known)
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh, max_keep=max(int(topk_per_image), 0))
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    scores_bf_multiply = scores_bf_multiply[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores_bf_multiply if vis else scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


class FastRCNNOutputLayers(nn.Module):
    """(1) proposal-to-detection box regression deltas, (2) classification scores — CLIP text embeddings as
    the classifier when `clip_cls_emb[0]` is set (fast_rcnn.py:440-475)."""

    def __init__(self, input_shape, *, box2box_transform, num_classes: int, test_score_thresh: float = 0.0,
                 test_nms_thresh: float = 0.5, soft_nms_enabled=False, soft_nms_method="gaussian",
                 soft_nms_sigma=0.5, soft_nms_prune=0.001, test_topk_per_image: int = 100,
                 cls_agnostic_bbox_reg: bool = False, smooth_l1_beta: float = 0.0,
                 box_reg_loss_type: str = "smooth_l1", loss_weight: Union[float, Dict[str, float]] = 1.0,
                 clip_cls_emb: tuple = (False, None), no_box_delta: bool = False, bg_cls_loss_weight=None,
                 multiply_rpn_score: tuple = (False, False), openset_test=None, strict_focal_nan: bool = False):
        super().__init__()
        self.box2box_transform = box2box_transform
        self.smooth_l1_beta = smooth_l1_beta
        self.test_score_thresh = test_score_thresh
        self.test_nms_thresh = test_nms_thresh
        self.soft_nms_enabled = soft_nms_enabled
        self.soft_nms_method = soft_nms_method
        self.soft_nms_sigma = soft_nms_sigma
        self.soft_nms_prune = soft_nms_prune
        self.test_topk_per_image = test_topk_per_image
        self.box_reg_loss_type = box_reg_loss_type
        if isinstance(loss_weight, float):
            loss_weight = {"loss_cls": loss_weight, "loss_box_reg": loss_weight}
        self.loss_weight = loss_weight
        self.strict_focal_nan = strict_focal_nan
        self.num_classes = num_classes
        if isinstance(input_shape, int):
            input_size = input_shape
        elif hasattr(input_shape, "channels"):
            input_size = input_shape.channels * (getattr(input_shape, "width", None) or 1) * \
                (getattr(input_shape, "height", None) or 1)
        else:
            input_size = int(input_shape)
        if openset_test is None:
            openset_test = (None, None, 0.01, None)

        self.use_clip_cls_emb = clip_cls_emb[0]
        if self.use_clip_cls_emb:
            if len(clip_cls_emb) >= 4 and clip_cls_emb[2] in ("CLIPRes5ROIHeads", "CLIPStandardROIHeads",
                                                             "CLIPRes5ROIHeadsPseudoLab"):
                input_size = clip_cls_emb[3]
            self.use_bias = False
            self.temperature = openset_test[2]
            self.cls_score = nn.Linear(input_size, num_classes, bias=self.use_bias)
            with torch.no_grad():
                if clip_cls_emb[1] is not None:  # path (as upstream) or an in-memory tensor
                    w = clip_cls_emb[1] if isinstance(clip_cls_emb[1], torch.Tensor) else torch.load(clip_cls_emb[1])
                    self.cls_score.weight.copy_(w)
                self.cls_score.weight.requires_grad = False  # frozen text embeddings
            self.cls_bg_score = nn.Linear(input_size, 1, bias=self.use_bias)
            with torch.no_grad():
                nn.init.constant_(self.cls_bg_score.weight, 0)  # zero background embedding
                self.cls_bg_score.weight.requires_grad = False
            self.test_cls_score = None
            if openset_test[1] is not None:
                w = openset_test[1] if isinstance(openset_test[1], torch.Tensor) else torch.load(openset_test[1])
                self.openset_test_num_cls = w.size(0)
                self.test_cls_score = nn.Linear(input_size, self.openset_test_num_cls, bias=self.use_bias)
                self.test_cls_score.weight.requires_grad = False
                with torch.no_grad():
                    self.test_cls_score.weight.copy_(w)
        else:
            self.cls_score = nn.Linear(input_size, num_classes + 1)
            nn.init.normal_(self.cls_score.weight, std=0.01)
            nn.init.constant_(self.cls_score.bias, 0)

        num_bbox_reg_classes = 1 if cls_agnostic_bbox_reg else num_classes
        box_dim = len(box2box_transform.weights)
        self.bbox_pred = nn.Linear(input_size, num_bbox_reg_classes * box_dim)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        nn.init.constant_(self.bbox_pred.bias, 0)

        self.cls_loss_weight = None
        if bg_cls_loss_weight is not None:
            self.cls_loss_weight = torch.ones(num_classes + 1)
            self.cls_loss_weight[-1] = bg_cls_loss_weight
        self._bg_cls_loss_weight = bg_cls_loss_weight
        self.focal_scaled_loss = openset_test[3]
        self.no_box_delta = no_box_delta
        self.multiply_rpn_score = multiply_rpn_score[0]
        self.vis = multiply_rpn_score[1]
        self._last = None  # (weakref to scores, x, weight) of the latest forward, for the fused loss

    @classmethod
    def from_config(cls, cfg, input_shape):
        """Same keys as fast_rcnn.py:499-527; `cfg` is the reference's CfgNode (attribute access)."""
        from .box_regression import Box2BoxTransform

        return cls(
            input_shape,
            box2box_transform=Box2BoxTransform(weights=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            cls_agnostic_bbox_reg=cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG,
            smooth_l1_beta=cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA,
            test_score_thresh=cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST,
            test_nms_thresh=cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST,
            soft_nms_enabled=cfg.MODEL.ROI_HEADS.SOFT_NMS_ENABLED,
            soft_nms_method=cfg.MODEL.ROI_HEADS.SOFT_NMS_METHOD,
            soft_nms_sigma=cfg.MODEL.ROI_HEADS.SOFT_NMS_SIGMA,
            soft_nms_prune=cfg.MODEL.ROI_HEADS.SOFT_NMS_PRUNE,
            test_topk_per_image=cfg.TEST.DETECTIONS_PER_IMAGE,
            box_reg_loss_type=cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE,
            loss_weight={"loss_box_reg": cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT},
            clip_cls_emb=(cfg.MODEL.CLIP.USE_TEXT_EMB_CLASSIFIER, cfg.MODEL.CLIP.TEXT_EMB_PATH,
                          cfg.MODEL.ROI_HEADS.NAME, cfg.MODEL.CLIP.TEXT_EMB_DIM),
            no_box_delta=cfg.MODEL.CLIP.NO_BOX_DELTA or cfg.MODEL.CLIP.CROP_REGION_TYPE == "GT",
            bg_cls_loss_weight=cfg.MODEL.CLIP.BG_CLS_LOSS_WEIGHT,
            multiply_rpn_score=(cfg.MODEL.CLIP.MULTIPLY_RPN_SCORE, cfg.MODEL.CLIP.VIS),
            openset_test=(cfg.MODEL.CLIP.OPENSET_TEST_NUM_CLASSES, cfg.MODEL.CLIP.OPENSET_TEST_TEXT_EMB_PATH,
                          cfg.MODEL.CLIP.CLSS_TEMP, cfg.MODEL.CLIP.FOCAL_SCALED_LOSS),
        )

    # ------------------------------------------------------------------ forward (fast_rcnn.py:529-572)
    def _active_weight(self) -> torch.Tensor:
        if not self.training and self.test_cls_score is not None:  # open-set inference (:549-552)
            return self.test_cls_score.weight
        return self.cls_score.weight

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        if self.use_clip_cls_emb:
            w = self._active_weight()
            scores = ops.clip_head_scores(x, w, self.cls_bg_score.weight, float(self.temperature))
            self._last = (weakref.ref(scores), x, w)
        else:
            scores = self.cls_score(x)
            self._last = None
        proposal_deltas = self.bbox_pred(x)
        return scores, proposal_deltas

    # ------------------------------------------------------------------- losses (fast_rcnn.py:574-622)
    def _loss_mode(self):
        if self.focal_scaled_loss is not None:
            bw = -1.0 if self.cls_loss_weight is None else float(self._bg_cls_loss_weight)
            return ops.LOSS_FOCAL, float(self.focal_scaled_loss), bw
        if self.cls_loss_weight is None:
            return ops.LOSS_CE, 0.0, -1.0
        return ops.LOSS_WEIGHTED_CE, 0.0, float(self._bg_cls_loss_weight)

    def losses(self, predictions, proposals):
        scores, proposal_deltas = predictions
        gt_classes = (cat([p.gt_classes for p in proposals], dim=0) if len(proposals)
                      else torch.empty(0, dtype=torch.int64, device=scores.device))  # upstream: float empty(0)
        if len(proposals):
            proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            assert not proposal_boxes.requires_grad, "Proposals should not require gradients!"
            gt_boxes = cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)

        fused = (self.use_clip_cls_emb and self._last is not None and self._last[0]() is scores
                 and gt_classes.numel() == scores.shape[0] and scores.is_cuda)
        if fused:
            _, x, w = self._last
            self._last = None   # do not keep x (and its graph) alive until the next forward
            mode, gamma, bgw = self._loss_mode()
            # when x will need its gradient it is produced by the same pass (the op keeps it for backward)
            loss_cls, _, _, counters = ops.clip_head_loss(x, w, self.cls_bg_score.weight,
                                                          gt_classes.to(scores.device), float(self.temperature),
                                                          mode, gamma, bgw, None, self.strict_focal_nan, False,
                                                          bool(x.requires_grad and torch.is_grad_enabled()))
            _log_classification_stats(scores, gt_classes, counters=counters)
        else:
            _log_classification_stats(scores, gt_classes)
            if self.cls_loss_weight is not None and self.cls_loss_weight.device != scores.device:
                self.cls_loss_weight = self.cls_loss_weight.to(scores.device)
            if self.focal_scaled_loss is not None:
                loss_cls = self.focal_loss(scores, gt_classes, gamma=self.focal_scaled_loss)
            elif self.cls_loss_weight is None:
                loss_cls = cross_entropy(scores, gt_classes, reduction="mean")
            else:
                loss_cls = cross_entropy(scores, gt_classes, reduction="mean", weight=self.cls_loss_weight)
        losses = {"loss_cls": loss_cls,
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def focal_loss(self, inputs, targets, gamma=0.5, reduction="mean"):
        """fast_rcnn.py:624-644 in PyTorch arithmetic (used only for scores that did not come from `forward`).
        Empty input returns the gradient-connected zero (the reference crashes there, :626-627)."""
        if targets.numel() == 0 and reduction == "mean":
            return inputs.sum() * 0.0
        ce_loss = F.cross_entropy(inputs, targets, reduction="none")
        p = F.softmax(inputs, dim=-1)
        p_t = p[torch.arange(p.size(0), device=p.device), targets]
        loss = ce_loss * ((1 - p_t) ** gamma)
        if self.cls_loss_weight is not None:
            loss_weight = torch.ones(loss.size(0), device=p.device)
            loss_weight[targets == self.num_classes] = float(self.cls_loss_weight[-1])
            loss = loss * loss_weight
        return loss.mean() if reduction == "mean" else loss

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        """fast_rcnn.py:646-689 (smooth-L1 on foreground rows, normalised by R)."""
        box_dim = proposal_boxes.shape[1]
        if self.box_reg_loss_type != "smooth_l1":
            raise ValueError(f"Invalid bbox reg loss type '{self.box_reg_loss_type}' (this build ships smooth_l1)")
        fg = (gt_classes >= 0) & (gt_classes < self.num_classes)
        if not STRICT_BOX_REG_SYNC and pred_deltas.is_cuda and (pred_deltas.shape[1] in (box_dim, box_dim * self.num_classes)) \
                and box_dim == 4:
            # one fused kernel (+ a fixed-order reduction): no nonzero(), no assert, no host sync
            loss, _ = ops.box_reg_loss(proposal_boxes, gt_boxes, pred_deltas, gt_classes, int(self.num_classes),
                                       [float(v) for v in self.box2box_transform.weights], float(self.smooth_l1_beta),
                                       bool(pred_deltas.requires_grad and torch.is_grad_enabled()))
            return loss
        if not STRICT_BOX_REG_SYNC:
            # same sum as below, evaluated on every row and masked: no nonzero(), no assert, no host sync
            if pred_deltas.shape[1] == box_dim:
                pred = pred_deltas
            else:
                cls = gt_classes.clamp(0, self.num_classes - 1)
                pred = pred_deltas.view(-1, self.num_classes, box_dim)[torch.arange(cls.numel(), device=cls.device), cls]
            tgt = self.box2box_transform.get_deltas(proposal_boxes, gt_boxes, valid=fg)
            per = smooth_l1_loss(pred, tgt, self.smooth_l1_beta, reduction="none")
            loss_box_reg = torch.where(fg[:, None], per, torch.zeros_like(per)).sum()
            return loss_box_reg / max(gt_classes.numel(), 1.0)
        fg_inds = nonzero_tuple(fg)[0]
        if pred_deltas.shape[1] == box_dim:
            fg_pred_deltas = pred_deltas[fg_inds]
        else:
            fg_pred_deltas = pred_deltas.view(-1, self.num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
        gt_pred_deltas = self.box2box_transform.get_deltas(proposal_boxes[fg_inds], gt_boxes[fg_inds])
        loss_box_reg = smooth_l1_loss(fg_pred_deltas, gt_pred_deltas, self.smooth_l1_beta, reduction="sum")
        return loss_box_reg / max(gt_classes.numel(), 1.0)

    # ---------------------------------------------------------------- inference (fast_rcnn.py:691-810)
    def inference(self, predictions, proposals):
        boxes = self.predict_boxes(predictions, proposals)
        scores = self.predict_probs(predictions, proposals)
        image_shapes = [x.image_size for x in proposals]
        scores_bf_multiply = scores
        if self.multiply_rpn_score and not self.training:
            rpn_scores = [p.get("objectness_logits") for p in proposals]
            scores = [(s * rpn_s[:, None]) ** 0.5 for s, rpn_s in zip(scores, rpn_scores)]
        return fast_rcnn_inference(boxes, scores, image_shapes, self.test_score_thresh, self.test_nms_thresh,
                                   self.soft_nms_enabled, self.soft_nms_method, self.soft_nms_sigma,
                                   self.soft_nms_prune, self.test_topk_per_image,
                                   scores_bf_multiply=scores_bf_multiply, vis=bool(self.vis))

    def predict_boxes(self, predictions, proposals):
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        num_prop_per_image = [len(p) for p in proposals]
        proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        if self.no_box_delta:
            predict_boxes = proposal_boxes
        else:
            predict_boxes = self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes)
        return predict_boxes.split(num_prop_per_image)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        num_inst_per_image = [len(p) for p in proposals]
        probs = F.softmax(scores, dim=-1)
        return probs.split(num_inst_per_image, dim=0)

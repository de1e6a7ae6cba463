"""PyTorch custom ops (`torch.ops.cddmsl_b200.*`) over the C ABI of include/cddmsl_b200.h.

Each op is a thin shim: validate, allocate outputs/workspace with the caching allocator, hand raw device
pointers + the current CUDA stream to the library.  Autograd is registered on the ops, so the reference's
call sites (`detectron2.layers.ROIAlign`, the box predictor, the consistency loss) differentiate through
them unchanged.  CUDA only — a CPU tensor raises.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

LOSS_FOCAL, LOSS_CE, LOSS_WEIGHTED_CE = 0, 1, 2


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _f32c(t: Tensor) -> Tensor:
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


# ------------------------------------------------------------------------------------------ ROIAlign
@torch.library.custom_op("cddmsl_b200::roi_align", mutates_args=(), device_types="cuda")
def roi_align(input: Tensor, rois: Tensor, spatial_scale: float, pooled_height: int, pooled_width: int,
              sampling_ratio: int, aligned: bool) -> Tensor:
    _lib.require_cuda(input, "input")
    _lib.require_cuda(rois, "rois")
    x, r = _f32c(input), _f32c(rois)
    n, c, h, w = x.shape
    nr = r.shape[0]
    out = torch.empty((nr, c, pooled_height, pooled_width), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    ws = _ws(L.cddmsl_roi_align_fwd_workspace_bytes(n, c, h, w, nr), x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.cddmsl_roi_align_fwd(_lib.ptr(x), _lib.ptr(r), _lib.ptr(out), n, c, h, w, nr, pooled_height,
                                          pooled_width, spatial_scale, sampling_ratio, int(aligned), _lib.ptr(ws),
                                          ws.numel(), _lib.stream_ptr(x.device)), "roi_align_fwd")
    return out if input.dtype == torch.float32 else out.to(input.dtype)


@roi_align.register_fake
def _(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned):
    return input.new_empty((rois.shape[0], input.shape[1], pooled_height, pooled_width))


@torch.library.custom_op("cddmsl_b200::roi_align_backward", mutates_args=(), device_types="cuda")
def roi_align_backward(grad: Tensor, rois: Tensor, spatial_scale: float, pooled_height: int, pooled_width: int,
                       batch_size: int, channels: int, height: int, width: int, sampling_ratio: int,
                       aligned: bool) -> Tensor:
    _lib.require_cuda(grad, "grad")
    g, r = _f32c(grad), _f32c(rois)
    nr = r.shape[0]
    gin = torch.empty((batch_size, channels, height, width), dtype=torch.float32, device=g.device)
    L = _lib.lib()
    ws = _ws(L.cddmsl_roi_align_bwd_workspace_bytes(batch_size, channels, height, width, nr), g.device)
    with torch.cuda.device(g.device):
        _lib.check(L.cddmsl_roi_align_bwd(_lib.ptr(g), _lib.ptr(r), _lib.ptr(gin), batch_size, channels, height,
                                          width, nr, pooled_height, pooled_width, spatial_scale, sampling_ratio,
                                          int(aligned), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(g.device)),
                   "roi_align_bwd")
    return gin if grad.dtype == torch.float32 else gin.to(grad.dtype)


@roi_align_backward.register_fake
def _(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
      sampling_ratio, aligned):
    return grad.new_empty((batch_size, channels, height, width))


def _roi_align_setup(ctx, inputs, output):
    input, rois, scale, ph, pw, sr, aligned = inputs
    ctx.save_for_backward(rois)
    ctx.meta = (scale, ph, pw, tuple(input.shape), sr, aligned)


def _roi_align_bwd(ctx, grad):
    (rois,) = ctx.saved_tensors
    scale, ph, pw, (n, c, h, w), sr, aligned = ctx.meta
    gin = None
    if ctx.needs_input_grad[0]:
        gin = roi_align_backward(grad, rois, scale, ph, pw, n, c, h, w, sr, aligned)
    return gin, None, None, None, None, None, None


roi_align.register_autograd(_roi_align_bwd, setup_context=_roi_align_setup)


# ---- the same RoIs on two maps (clip_roi_heads.py:117-132) -------------------------------------------------------
@torch.library.custom_op("cddmsl_b200::roi_align_pair", mutates_args=(), device_types="cuda")
def roi_align_pair(input_a: Tensor, input_b: Tensor, rois: Tensor, spatial_scale: float, pooled_height: int,
                   pooled_width: int, sampling_ratio: int, aligned: bool) -> Tuple[Tensor, Tensor]:
    _lib.require_cuda(input_a, "input_a")
    _lib.require_cuda(input_b, "input_b")
    _lib.require_cuda(rois, "rois")
    assert input_a.shape == input_b.shape, "roi_align_pair: the two feature maps must have the same shape"
    xa, xb, r = _f32c(input_a), _f32c(input_b), _f32c(rois)
    n, c, h, w = xa.shape
    nr = r.shape[0]
    oa = torch.empty((nr, c, pooled_height, pooled_width), dtype=torch.float32, device=xa.device)
    ob = torch.empty_like(oa)
    L = _lib.lib()
    ws = _ws(L.cddmsl_roi_align_fwd_workspace_bytes(n, c, h, w, nr), xa.device)
    with torch.cuda.device(xa.device):
        _lib.check(L.cddmsl_roi_align_fwd2(_lib.ptr(xa), _lib.ptr(xb), _lib.ptr(r), _lib.ptr(oa), _lib.ptr(ob), n, c,
                                           h, w, nr, pooled_height, pooled_width, spatial_scale, sampling_ratio,
                                           int(aligned), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(xa.device)),
                   "roi_align_fwd2")
    if input_a.dtype != torch.float32:
        oa, ob = oa.to(input_a.dtype), ob.to(input_a.dtype)
    return oa, ob


@roi_align_pair.register_fake
def _(input_a, input_b, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned):
    shape = (rois.shape[0], input_a.shape[1], pooled_height, pooled_width)
    return input_a.new_empty(shape), input_a.new_empty(shape)


@torch.library.custom_op("cddmsl_b200::roi_align_pair_backward", mutates_args=(), device_types="cuda")
def roi_align_pair_backward(grad_a: Tensor, grad_b: Tensor, rois: Tensor, spatial_scale: float, pooled_height: int,
                            pooled_width: int, batch_size: int, channels: int, height: int, width: int,
                            sampling_ratio: int, aligned: bool) -> Tuple[Tensor, Tensor]:
    ga, gb, r = _f32c(grad_a), _f32c(grad_b), _f32c(rois)
    nr = r.shape[0]
    ia = torch.empty((batch_size, channels, height, width), dtype=torch.float32, device=ga.device)
    ib = torch.empty_like(ia)
    L = _lib.lib()
    ws = _ws(L.cddmsl_roi_align_bwd_workspace_bytes(batch_size, channels, height, width, nr), ga.device)
    with torch.cuda.device(ga.device):
        _lib.check(L.cddmsl_roi_align_bwd2(_lib.ptr(ga), _lib.ptr(gb), _lib.ptr(r), _lib.ptr(ia), _lib.ptr(ib),
                                           batch_size, channels, height, width, nr, pooled_height, pooled_width,
                                           spatial_scale, sampling_ratio, int(aligned), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr(ga.device)), "roi_align_bwd2")
    return ia, ib


@roi_align_pair_backward.register_fake
def _(grad_a, grad_b, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
      sampling_ratio, aligned):
    shape = (batch_size, channels, height, width)
    return grad_a.new_empty(shape), grad_a.new_empty(shape)


def _roi_pair_setup(ctx, inputs, output):
    input_a, _, rois, scale, ph, pw, sr, aligned = inputs
    ctx.save_for_backward(rois)
    ctx.meta = (scale, ph, pw, tuple(input_a.shape), sr, aligned)


def _roi_pair_bwd(ctx, grad_a, grad_b):
    (rois,) = ctx.saved_tensors
    scale, ph, pw, (n, c, h, w), sr, aligned = ctx.meta
    ga = gb = None
    if ctx.needs_input_grad[0] and ctx.needs_input_grad[1]:
        ga, gb = roi_align_pair_backward(grad_a, grad_b, rois, scale, ph, pw, n, c, h, w, sr, aligned)
    elif ctx.needs_input_grad[0]:
        ga = roi_align_backward(grad_a, rois, scale, ph, pw, n, c, h, w, sr, aligned)
    elif ctx.needs_input_grad[1]:
        gb = roi_align_backward(grad_b, rois, scale, ph, pw, n, c, h, w, sr, aligned)
    return ga, gb, None, None, None, None, None, None


roi_align_pair.register_autograd(_roi_pair_bwd, setup_context=_roi_pair_setup)


# ------------------------------------------------------------------------------------------------ NMS
@torch.library.custom_op("cddmsl_b200::batched_nms", mutates_args=(), device_types="cuda")
def batched_nms(boxes: Tensor, scores: Tensor, idxs: Optional[Tensor], iou_threshold: float,
                coord_trick: bool, max_keep: int = 0) -> Tensor:
    """Kept original indices, score-descending (ties: lower index first).  One device->host read of the
    kept count sizes the result (the same sync PyTorch needs for any data-dependent shape).
    max_keep > 0: the first max_keep entries only (bit-identical prefix; the kernels stop early)."""
    _lib.require_cuda(boxes, "boxes")
    b = _f32c(boxes)
    s = _f32c(scores)
    m = b.shape[0]
    if m == 0:
        return torch.empty((0,), dtype=torch.int64, device=b.device)
    ids = None if idxs is None else idxs.to(torch.int64).contiguous()
    L = _lib.lib()
    keep = torch.empty((m,), dtype=torch.int64, device=b.device)
    nk = torch.empty((1,), dtype=torch.int32, device=b.device)
    ws = _ws(L.cddmsl_nms_workspace_bytes(m), b.device)
    with torch.cuda.device(b.device):
        _lib.check(L.cddmsl_nms_topk(_lib.ptr(b), _lib.ptr(s), _lib.ptr(ids), m, float(iou_threshold),
                                     int(coord_trick), int(max_keep), 0, _lib.ptr(keep), _lib.ptr(nk), _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr(b.device)), "nms_topk")
    return keep[: int(nk.item())]


@batched_nms.register_fake
def _(boxes, scores, idxs, iou_threshold, coord_trick, max_keep=0):
    n = torch.library.get_ctx().new_dynamic_size()
    return boxes.new_empty((n,), dtype=torch.int64)


@torch.library.custom_op("cddmsl_b200::nms_images", mutates_args=(), device_types="cuda")
def nms_images(boxes: Tensor, scores: Tensor, idxs: Optional[Tensor], counts: Tensor, iou_threshold: float,
               coord_trick: bool, max_keep: int = 0, presorted: bool = False) -> Tuple[Tensor, Tensor]:
    """NMS of B images in one launch sequence.  boxes [B,M,4], scores [B,M], idxs [B,M] or None, counts int32 [B]
    (device; image b uses its first counts[b] rows).  Returns (keep int64 [B,M], num_keep int32 [B]): keep[b, :num_keep[b]]
    are the kept row indices of image b in score order.  No host sync here -- the caller decides when to read.
    max_keep > 0: only the first max_keep kept boxes per image are produced (bit-identical prefix of the full list).
    presorted: the scores of every image are already non-increasing (the sort is skipped)."""
    _lib.require_cuda(boxes, "boxes")
    b = _f32c(boxes)
    s = _f32c(scores)
    nb, m = s.shape
    assert b.shape == (nb, m, 4) and counts.numel() == nb
    keep = torch.empty((nb, m), dtype=torch.int64, device=b.device)
    nk = torch.zeros((nb,), dtype=torch.int32, device=b.device)
    if nb == 0 or m == 0:
        return keep, nk
    ids = None if idxs is None else idxs.to(torch.int64).contiguous()
    cnt = counts.to(device=b.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    ws = _ws(L.cddmsl_nms_batched_workspace_bytes(nb, m), b.device)
    with torch.cuda.device(b.device):
        _lib.check(L.cddmsl_nms_batched_topk(_lib.ptr(b), _lib.ptr(s), _lib.ptr(ids), _lib.ptr(cnt), nb, m,
                                             float(iou_threshold), int(coord_trick), int(max_keep), int(presorted),
                                             _lib.ptr(keep), _lib.ptr(nk), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr(b.device)),
                   "nms_batched_topk")
    return keep, nk


@nms_images.register_fake
def _(boxes, scores, idxs, counts, iou_threshold, coord_trick, max_keep=0, presorted=False):
    return scores.new_empty(scores.shape, dtype=torch.int64), scores.new_empty((scores.shape[0],), dtype=torch.int32)


# ------------------------------------------------------------------------------------ proposal matching
@torch.library.custom_op("cddmsl_b200::match_boxes", mutates_args=(), device_types="cuda")
def match_boxes(gt_boxes: Tensor, gt_counts: Tensor, boxes: Tensor, counts: Optional[Tensor],
                thresholds: List[float], labels: List[int], allow_low_quality_matches: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """pairwise IoU + Matcher for B images at once (boxes.py:346-368, matcher.py:63-127), no IoU matrix.
    gt_boxes [B,G,4], gt_counts int32 [B], boxes [B,M,4], counts int32 [B] or None.
    Returns (matches int64 [B,M], match_labels int8 [B,M], matched_vals float [B,M])."""
    import ctypes

    _lib.require_cuda(boxes, "boxes")
    bx, gb = _f32c(boxes), _f32c(gt_boxes)
    nb, m = bx.shape[0], bx.shape[1]
    g = gb.shape[1]
    assert bx.shape == (nb, m, 4) and gb.shape == (nb, g, 4) and gt_counts.numel() == nb
    dev = bx.device
    matches = torch.zeros((nb, m), dtype=torch.int64, device=dev)
    mlabels = torch.zeros((nb, m), dtype=torch.int8, device=dev)
    mvals = torch.zeros((nb, m), dtype=torch.float32, device=dev)
    if nb == 0 or m == 0:
        return matches, mlabels, mvals
    gc = gt_counts.to(device=dev, dtype=torch.int32).contiguous()
    cnt = None if counts is None else counts.to(device=dev, dtype=torch.int32).contiguous()
    thr = (ctypes.c_float * len(thresholds))(*[float(t) for t in thresholds])
    lab = (ctypes.c_int32 * len(labels))(*[int(v) for v in labels])
    assert len(labels) == len(thresholds) + 1
    L = _lib.lib()
    ws = _ws(L.cddmsl_match_boxes_workspace_bytes(nb, g), dev)
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_match_boxes(_lib.ptr(gb) if g > 0 else None, _lib.ptr(gc), _lib.ptr(bx), _lib.ptr(cnt), nb,
                                        g, m, ctypes.cast(thr, ctypes.c_void_p), ctypes.cast(lab, ctypes.c_void_p),
                                        len(thresholds), int(allow_low_quality_matches), _lib.ptr(matches),
                                        _lib.ptr(mlabels), _lib.ptr(mvals), _lib.ptr(ws), ws.numel(),
                                        _lib.stream_ptr(dev)), "match_boxes")
    return matches, mlabels, mvals


@match_boxes.register_fake
def _(gt_boxes, gt_counts, boxes, counts, thresholds, labels, allow_low_quality_matches):
    nb, m = boxes.shape[0], boxes.shape[1]
    return (boxes.new_empty((nb, m), dtype=torch.int64), boxes.new_empty((nb, m), dtype=torch.int8),
            boxes.new_empty((nb, m), dtype=torch.float32))


# ------------------------------------------------------------------------------------------ CLIP head
def _head_ws(r, d, k, device):
    return _ws(_lib.lib().cddmsl_clip_head_workspace_bytes(r, d, k), device)


@torch.library.custom_op("cddmsl_b200::clip_head_scores", mutates_args=(), device_types="cuda")
def clip_head_scores(x: Tensor, weight: Tensor, bg_weight: Tensor, temperature: float) -> Tensor:
    _lib.require_cuda(x, "x")
    xx, w, wb = _f32c(x), _f32c(weight), _f32c(bg_weight).reshape(-1)
    r, d = xx.shape
    k = w.shape[0]
    scores = torch.empty((r, k + 1), dtype=torch.float32, device=xx.device)
    ws = _head_ws(r, d, k, xx.device)
    with torch.cuda.device(xx.device):
        _lib.check(_lib.lib().cddmsl_clip_head_scores(_lib.ptr(xx), _lib.ptr(w), _lib.ptr(wb), r, d, k, temperature,
                                                      _lib.ptr(scores), _lib.ptr(ws), ws.numel(),
                                                      _lib.stream_ptr(xx.device)), "clip_head_scores")
    return scores


@clip_head_scores.register_fake
def _(x, weight, bg_weight, temperature):
    return x.new_empty((x.shape[0], weight.shape[0] + 1))


@torch.library.custom_op("cddmsl_b200::clip_head_scores_backward", mutates_args=(), device_types="cuda")
def clip_head_scores_backward(x: Tensor, weight: Tensor, bg_weight: Tensor, dscores: Tensor,
                              temperature: float) -> Tensor:
    xx, w, wb, ds = _f32c(x), _f32c(weight), _f32c(bg_weight).reshape(-1), _f32c(dscores)
    r, d = xx.shape
    k = w.shape[0]
    dx = torch.empty_like(xx)
    ws = _head_ws(r, d, k, xx.device)
    with torch.cuda.device(xx.device):
        _lib.check(_lib.lib().cddmsl_clip_head_scores_bwd(_lib.ptr(xx), _lib.ptr(w), _lib.ptr(wb), _lib.ptr(ds), r, d,
                                                          k, temperature, _lib.ptr(dx), _lib.ptr(ws), ws.numel(),
                                                          _lib.stream_ptr(xx.device)), "clip_head_scores_bwd")
    return dx


@clip_head_scores_backward.register_fake
def _(x, weight, bg_weight, dscores, temperature):
    return torch.empty_like(x)


def _frozen_embeddings(weight, bg_weight):
    """The reference freezes the concept and background embeddings (fast_rcnn.py:453, :461); the fused backward only
    produces d/dx.  Unfrozen embeddings would silently train on zero gradients: refuse instead."""
    if weight.requires_grad or bg_weight.requires_grad:
        raise RuntimeError("cddmsl_b200 CLIP head: the concept / background embeddings must be frozen "
                           "(requires_grad=False, as in the reference); no gradient is computed for them")


def _scores_setup(ctx, inputs, output):
    x, weight, bg_weight, temperature = inputs
    _frozen_embeddings(weight, bg_weight)
    ctx.save_for_backward(x, weight, bg_weight)
    ctx.temperature = temperature


def _scores_bwd(ctx, ds):
    x, weight, bg_weight = ctx.saved_tensors
    dx = clip_head_scores_backward(x, weight, bg_weight, ds, ctx.temperature) if ctx.needs_input_grad[0] else None
    # concept / background embeddings are frozen in the reference (fast_rcnn.py:453,461)
    return dx, None, None, None


clip_head_scores.register_autograd(_scores_bwd, setup_context=_scores_setup)


@torch.library.custom_op("cddmsl_b200::clip_head_loss", mutates_args=(), device_types="cuda")
def clip_head_loss(x: Tensor, weight: Tensor, bg_weight: Tensor, gt: Tensor, temperature: float, loss_mode: int,
                   gamma: float, bg_cls_weight: float, grad_scale: Optional[Tensor], strict_nan: bool,
                   want_scores: bool, want_dx: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Fused logits -> loss (-> dx).  Returns (loss[], scores[R,K+1] or empty, dx[R,D] or empty, stats int32[4])."""
    _lib.require_cuda(x, "x")
    xx, w, wb = _f32c(x), _f32c(weight), _f32c(bg_weight).reshape(-1)
    g = gt.to(torch.int64).contiguous()
    r, d = xx.shape
    k = w.shape[0]
    dev = xx.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    scores = torch.empty((r, k + 1) if want_scores else (0,), dtype=torch.float32, device=dev)
    dx = torch.empty((r, d) if want_dx else (0,), dtype=torch.float32, device=dev)
    stats = torch.empty((4,), dtype=torch.int32, device=dev)
    gs = None if grad_scale is None else _f32c(grad_scale)
    ws = _head_ws(r, d, k, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cddmsl_clip_head_loss(
            _lib.ptr(xx), _lib.ptr(w), _lib.ptr(wb), _lib.ptr(g), r, d, k, temperature, loss_mode, gamma,
            bg_cls_weight, _lib.ptr(gs), int(strict_nan), _lib.ptr(scores) if want_scores else None, _lib.ptr(loss),
            _lib.ptr(dx) if want_dx else None, _lib.ptr(stats), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
            "clip_head_loss")
    return loss, scores, dx, stats


@clip_head_loss.register_fake
def _(x, weight, bg_weight, gt, temperature, loss_mode, gamma, bg_cls_weight, grad_scale, strict_nan, want_scores,
      want_dx):
    r, d = x.shape
    k = weight.shape[0]
    return (x.new_empty(()), x.new_empty((r, k + 1) if want_scores else (0,)),
            x.new_empty((r, d) if want_dx else (0,)), x.new_empty((4,), dtype=torch.int32))


def _loss_setup(ctx, inputs, output):
    x, weight, bg_weight, gt, temperature, loss_mode, gamma, bg_cls_weight, gs, strict_nan, _ws_, want_dx = inputs
    _frozen_embeddings(weight, bg_weight)
    # only `loss` is differentiable: the optional scores / dx / counters are side products of the same pass
    ctx.mark_non_differentiable(output[1], output[2], output[3])
    ctx.have_dx = bool(want_dx) and gs is None
    if ctx.have_dx:
        ctx.save_for_backward(output[2])  # d loss / d x for a unit upstream gradient, produced by the forward pass
    else:
        ctx.save_for_backward(x, weight, bg_weight, gt)
        ctx.meta = (temperature, loss_mode, gamma, bg_cls_weight, strict_nan)


def _loss_bwd(ctx, gloss, gscores, gdx, gstats):
    dx = None
    if ctx.needs_input_grad[0]:
        if ctx.have_dx:
            (dx1,) = ctx.saved_tensors
            dx = dx1 * gloss
        else:
            x, weight, bg_weight, gt = ctx.saved_tensors
            t, mode, gamma, bgw, strict = ctx.meta
            # recompute from x with the upstream scalar folded in: 1 read of x + 1 write of dx
            _, _, dx, _ = clip_head_loss(x, weight, bg_weight, gt, t, mode, gamma, bgw, gloss.reshape(1), strict,
                                         False, True)
    return (dx,) + (None,) * 11


clip_head_loss.register_autograd(_loss_bwd, setup_context=_loss_setup)


# ------------------------------------------------------------------------------------ box regression loss
@torch.library.custom_op("cddmsl_b200::box_reg_loss", mutates_args=(), device_types="cuda")
def box_reg_loss(proposal_boxes: Tensor, gt_boxes: Tensor, pred_deltas: Tensor, gt: Tensor, num_classes: int,
                 weights: List[float], beta: float, want_grad: bool) -> Tuple[Tensor, Tensor]:
    """Smooth-L1 box regression loss over foreground rows / R (fast_rcnn.py:646-689), no host sync.
    Returns (loss[], d loss / d pred_deltas or empty)."""
    _lib.require_cuda(pred_deltas, "pred_deltas")
    pb, gb, pd = _f32c(proposal_boxes), _f32c(gt_boxes), _f32c(pred_deltas)
    g = gt.to(torch.int64).contiguous()
    r = pd.shape[0]
    dev = pd.device
    agnostic = pd.shape[1] == 4
    assert agnostic or pd.shape[1] == 4 * num_classes, "pred_deltas must be [R,4] or [R,4K]"
    assert pb.shape == (r, 4) and gb.shape == (r, 4) and g.numel() == r
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dpred = torch.empty(pd.shape if want_grad else (0,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = torch.empty(int(L.cddmsl_box_reg_loss_workspace_bytes(r)), dtype=torch.uint8, device=dev)
    wx, wy, ww, wh = [float(v) for v in weights]
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_box_reg_loss(_lib.ptr(pb), _lib.ptr(gb), _lib.ptr(pd), _lib.ptr(g), r, num_classes,
                                         int(agnostic), wx, wy, ww, wh, beta, None, _lib.ptr(loss),
                                         _lib.ptr(dpred) if want_grad and r > 0 else None, _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr(dev)), "box_reg_loss")
    return loss, dpred


@box_reg_loss.register_fake
def _(proposal_boxes, gt_boxes, pred_deltas, gt, num_classes, weights, beta, want_grad):
    return pred_deltas.new_empty(()), pred_deltas.new_empty(pred_deltas.shape if want_grad else (0,))


def _boxreg_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.dtype = inputs[2].dtype


def _boxreg_bwd(ctx, gloss, _gd):
    (dpred,) = ctx.saved_tensors
    g = (dpred * gloss).to(ctx.dtype) if dpred.numel() else None
    return None, None, g, None, None, None, None, None


box_reg_loss.register_autograd(_boxreg_bwd, setup_context=_boxreg_setup)


# ------------------------------------------------------------------------------------ alignment loss
@torch.library.custom_op("cddmsl_b200::align_pack", mutates_args=(), device_types="cuda")
def align_pack(src: Tensor, tgt: Tensor) -> Tuple[Tensor, Tensor]:
    _lib.require_cuda(src, "src")
    a, b = _f32c(src), _f32c(tgt)
    n, d = a.shape
    packed = torch.empty((2, n, d), dtype=torch.float32, device=a.device)
    norms = torch.empty((2, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().cddmsl_align_pack_normalized(_lib.ptr(a), _lib.ptr(b), n, d, _lib.ptr(packed),
                                                           _lib.ptr(norms), _lib.stream_ptr(a.device)),
                   "align_pack_normalized")
    return packed, norms


@align_pack.register_fake
def _(src, tgt):
    n, d = src.shape
    return src.new_empty((2, n, d)), src.new_empty((2, n))


@torch.library.custom_op("cddmsl_b200::align_loss", mutates_args=(), device_types="cuda")
def align_loss(packed_all: Tensor, norms_local: Tensor, rank: int, grad_scale: Optional[Tensor],
               want_grads: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """packed_all [world,2,n_local,D] (normalised rows).  Returns (loss[], da, db [n_local,D] or empty)."""
    _lib.require_cuda(packed_all, "packed_all")
    p = _f32c(packed_all)
    world, two, n_local, d = p.shape
    assert two == 2
    dev = p.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    shape = (n_local, d) if want_grads else (0,)
    da = torch.empty(shape, dtype=torch.float32, device=dev)
    db = torch.empty(shape, dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _ws(L.cddmsl_align_loss_workspace_bytes(world, n_local, d), dev)
    gs = None if grad_scale is None else _f32c(grad_scale)
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_align_loss(_lib.ptr(p), _lib.ptr(_f32c(norms_local)), world, n_local, d, rank,
                                       _lib.ptr(gs), _lib.ptr(loss), _lib.ptr(da) if want_grads else None,
                                       _lib.ptr(db) if want_grads else None, _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(dev)), "align_loss")
    return loss, da, db


@align_loss.register_fake
def _(packed_all, norms_local, rank, grad_scale, want_grads):
    world, _, n_local, d = packed_all.shape
    shape = (n_local, d) if want_grads else (0,)
    return packed_all.new_empty(()), packed_all.new_empty(shape), packed_all.new_empty(shape)


# ------------------------------------------------------------------------------------------ RPN decode
@torch.library.custom_op("cddmsl_b200::rpn_decode_topk", mutates_args=(), device_types="cuda")
def rpn_decode_topk(anchors: Tensor, deltas: Tensor, topk_idx: Tensor, topk_scores: Tensor, image_hw: Tensor,
                    weights: List[float], scale_clamp: float,
                    min_box_size: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """rpn.py:514-533 + box_regression.py:77-117 + proposal_utils.py:95-114 for a batch (csrc/rpn_decode.cu).
    Returns (boxes [N,K,4], scores [N,K] -- survivors first, score order --, counts int32 [N], all_finite int32 [1])."""
    _lib.require_cuda(deltas, "deltas")
    an, dl = _f32c(anchors), _f32c(deltas)
    idx = topk_idx.to(torch.int64).contiguous()
    sc, hw = _f32c(topk_scores), _f32c(image_hw)
    n, a = dl.shape[0], an.shape[0]
    k = idx.shape[1]
    dev = dl.device
    boxes = torch.empty((n, k, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((n, k), dtype=torch.float32, device=dev)
    counts = torch.empty((n,), dtype=torch.int32, device=dev)
    fin = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cddmsl_rpn_decode_topk(_lib.ptr(an), _lib.ptr(dl), _lib.ptr(idx), _lib.ptr(sc), _lib.ptr(hw),
                                                     n, a, k, float(weights[0]), float(weights[1]), float(weights[2]),
                                                     float(weights[3]), float(scale_clamp), float(min_box_size),
                                                     _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(counts), _lib.ptr(fin),
                                                     _lib.stream_ptr(dev)), "rpn_decode_topk")
    return boxes, scores, counts, fin


@rpn_decode_topk.register_fake
def _(anchors, deltas, topk_idx, topk_scores, image_hw, weights, scale_clamp, min_box_size):
    n, k = topk_idx.shape
    return (deltas.new_empty((n, k, 4)), deltas.new_empty((n, k)), deltas.new_empty((n,), dtype=torch.int32),
            deltas.new_empty((1,), dtype=torch.int32))


# ------------------------------------------------------------------------------------------ KD regulariser
@torch.library.custom_op("cddmsl_b200::kd_l1", mutates_args=(), device_types="cuda")
def kd_l1(teacher: Tensor, student: Tensor, want_grad: bool) -> Tuple[Tensor, Tensor]:
    """rcnn.py:265-272: mean |teacher - student| and, from the same pass, d loss / d student (unit upstream scale)."""
    _lib.require_cuda(student, "student")
    _lib.require_cuda(teacher, "teacher")
    assert teacher.shape == student.shape, "L1Loss: teacher and student features must have the same shape"
    t, s = _f32c(teacher), _f32c(student)
    dev = s.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ds = torch.empty_like(s) if want_grad else torch.empty((0,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _ws(L.cddmsl_kd_l1_loss_workspace_bytes(), dev)
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_kd_l1_loss(_lib.ptr(t), _lib.ptr(s), s.numel(), None, _lib.ptr(loss),
                                       _lib.ptr(ds) if want_grad else None, _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(dev)), "kd_l1_loss")
    return loss, ds


@kd_l1.register_fake
def _(teacher, student, want_grad):
    return student.new_empty(()), (torch.empty_like(student) if want_grad else student.new_empty((0,)))


def _kd_setup(ctx, inputs, output):
    ctx.have = inputs[2]
    if inputs[2]:
        ctx.save_for_backward(output[1])


def _kd_bwd(ctx, gloss, _gds):
    if not ctx.have:
        raise RuntimeError("kd_l1: backward without a recorded gradient")
    (ds,) = ctx.saved_tensors
    return None, ds * gloss, None   # the teacher is detached (rcnn.py:268)


kd_l1.register_autograd(_kd_bwd, setup_context=_kd_setup)


# ------------------------------------------------------------------------------ RegionCLIP pretraining losses
@torch.library.custom_op("cddmsl_b200::contrastive_loss", mutates_args=(), device_types="cuda")
def contrastive_loss(packed_all: Tensor, norms_local: Tensor, rank: int, logit_scale: float, grad_mult: float,
                     want_grads: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """`align_loss` with a temperature (logits = logit_scale * A^ B^T) and a gradient multiplier (clip_rcnn.py:608-640
    with comm.py:268-322).  Returns (loss[], da, db [n_local,D] or empty)."""
    _lib.require_cuda(packed_all, "packed_all")
    p = _f32c(packed_all)
    world, two, n_local, d = p.shape
    assert two == 2
    dev = p.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    shape = (n_local, d) if want_grads else (0,)
    da = torch.empty(shape, dtype=torch.float32, device=dev)
    db = torch.empty(shape, dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _ws(L.cddmsl_align_loss_workspace_bytes(world, n_local, d), dev)
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_contrastive_loss(_lib.ptr(p), _lib.ptr(_f32c(norms_local)), world, n_local, d, rank,
                                             float(logit_scale), float(grad_mult), None, _lib.ptr(loss),
                                             _lib.ptr(da) if want_grads else None, _lib.ptr(db) if want_grads else None,
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "contrastive_loss")
    return loss, da, db


@contrastive_loss.register_fake
def _(packed_all, norms_local, rank, logit_scale, grad_mult, want_grads):
    world, _, n_local, d = packed_all.shape
    shape = (n_local, d) if want_grads else (0,)
    return packed_all.new_empty(()), packed_all.new_empty(shape), packed_all.new_empty(shape)


TARGET_KL, TARGET_MIL = 0, 1


@torch.library.custom_op("cddmsl_b200::softmax_target_loss", mutates_args=(), device_types="cuda")
def softmax_target_loss(logits: Tensor, target: Tensor, mode: int, want_grad: bool) -> Tuple[Tensor, Tensor]:
    """Row-softmax loss against a dense target (clip_rcnn.py:597-606): logits [R, ld >= K], target [R, K]; the first K
    logit columns count.  Returns (loss[], d loss / d logits [R, ld] or empty)."""
    _lib.require_cuda(logits, "logits")
    lg, tg = _f32c(logits), _f32c(target)
    r, ld = lg.shape
    k = tg.shape[1]
    assert tg.shape[0] == r and ld >= k
    dev = lg.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dl = torch.empty_like(lg) if want_grad else torch.empty((0,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _ws(L.cddmsl_softmax_target_loss_workspace_bytes(r), dev)
    with torch.cuda.device(dev):
        _lib.check(L.cddmsl_softmax_target_loss(_lib.ptr(lg), ld, _lib.ptr(tg), r, k, int(mode), None, _lib.ptr(loss),
                                                _lib.ptr(dl) if want_grad else None, _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(dev)), "softmax_target_loss")
    return loss, dl


@softmax_target_loss.register_fake
def _(logits, target, mode, want_grad):
    return logits.new_empty(()), (torch.empty_like(logits) if want_grad else logits.new_empty((0,)))


def _tgt_setup(ctx, inputs, output):
    ctx.have = inputs[3]
    if inputs[3]:
        ctx.save_for_backward(output[1])


def _tgt_bwd(ctx, gloss, _gdl):
    if not ctx.have:
        raise RuntimeError("softmax_target_loss: backward without a recorded gradient")
    (dl,) = ctx.saved_tensors
    return dl * gloss, None, None, None   # the target is the teacher's (detached) distribution / a label matrix


softmax_target_loss.register_autograd(_tgt_bwd, setup_context=_tgt_setup)

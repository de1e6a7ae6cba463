"""Caption-consistency alignment loss (detectron2/modeling/meta_arch/rcnn.py:305-317 image level,
:455-468 region level; the reference has no function boundary for it — the code is inline in
`GeneralizedRCNN`).  One call replaces: 2x GatherLayer, 2x row normalisation, the n x n matmul, two
cross-entropies and their autograd graph.

Data-parallel semantics: each rank owns `n_local` rows of src and tgt; rows are normalised
and packed locally, ONE all-gather (NCCL over NVLink) exchanges them, every rank evaluates the full
symmetric InfoNCE and back-propagates only into its own rows — exactly what GatherLayer.backward
(gather.py:16-20) yields.  No collective runs in backward.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import ops


def shard_bounds(n_total: int, world: int, rank: int):
    """Images (and all their RoIs / embedding rows) are split contiguously and evenly over ranks."""
    assert n_total % world == 0, "the reference's sampler gives every rank the same number of images"
    per = n_total // world
    return rank * per, (rank + 1) * per


class _CaptionConsistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, group):
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        packed, norms = ops.align_pack(a, b)
        if world > 1:
            packed_all = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(packed_all, packed, group=group)
        else:
            packed_all = packed.unsqueeze(0)
        # The gradient of the local rows shares the logits / log-sum-exps with the loss: when a gradient will be
        # asked for it is produced in the same pass (unit upstream scale) instead of re-evaluating S in backward.
        want = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        loss, da, db = ops.align_loss(packed_all, norms, rank, None, want)
        if want:
            ctx.save_for_backward(da, db)
        ctx.have = want
        return loss

    @staticmethod
    def backward(ctx, gloss):
        if not ctx.have:  # forward ran under no_grad-like conditions; nothing was kept
            raise RuntimeError("caption_consistency_loss: backward without a recorded forward")
        da, db = ctx.saved_tensors
        return da * gloss, db * gloss, None


def caption_consistency_loss(src: torch.Tensor, tgt: torch.Tensor,
                             group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """rcnn.py:455-468: `S = norm(gather(src)) @ norm(gather(tgt)).T`, `(CE(S, I) + CE(S.T, I)) / 2`.
    src, tgt: this rank's [n_local, D] projector outputs."""
    assert src.dim() == 2 and src.shape == tgt.shape
    return _CaptionConsistency.apply(src, tgt, group)


def image_caption_consistency_loss(trgt: torch.Tensor, src: torch.Tensor,
                                   group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """rcnn.py:305-317 (`v2l_contrastive` tail): same loss with the operands in the reference's order,
    `joint = trgt @ src.T` (the loss is symmetric in the pair, gradients follow the operands)."""
    return caption_consistency_loss(trgt, src, group)


class _CaptionConsistencyPair(torch.autograd.Function):
    """Both branches of a training step (image level rcnn.py:305-317 and region level :455-468) with ONE all-gather:
    the two packed, normalised buffers travel in one NCCL message (4 all-gathers in the reference, 2 with one call per
    branch, 1 here -- the collective is latency-bound at 4-64 KB per rank)."""

    @staticmethod
    def forward(ctx, a1, b1, a2, b2, group):
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        p1, n1 = ops.align_pack(a1, b1)
        p2, n2 = ops.align_pack(a2, b2)
        if world > 1:
            flat = torch.cat([p1.reshape(-1), p2.reshape(-1)])
            allf = torch.empty((world, flat.numel()), dtype=flat.dtype, device=flat.device)
            dist.all_gather_into_tensor(allf, flat, group=group)
            all1 = allf[:, : p1.numel()].reshape((world,) + tuple(p1.shape))
            all2 = allf[:, p1.numel():].reshape((world,) + tuple(p2.shape))
        else:
            all1, all2 = p1.unsqueeze(0), p2.unsqueeze(0)
        want = any(ctx.needs_input_grad[:4])
        l1, da1, db1 = ops.align_loss(all1, n1, rank, None, want)
        l2, da2, db2 = ops.align_loss(all2, n2, rank, None, want)
        if want:
            ctx.save_for_backward(da1, db1, da2, db2)
        ctx.have = want
        return l1, l2

    @staticmethod
    def backward(ctx, g1, g2):
        if not ctx.have:
            raise RuntimeError("caption_consistency_losses: backward without a recorded forward")
        da1, db1, da2, db2 = ctx.saved_tensors
        return da1 * g1, db1 * g1, da2 * g2, db2 * g2, None


def caption_consistency_losses(trgt_img: torch.Tensor, src_img: torch.Tensor, src_reg: torch.Tensor,
                               tgt_reg: torch.Tensor, group: Optional["dist.ProcessGroup"] = None):
    """(image-level loss, region-level loss) = (`image_caption_consistency_loss(trgt_img, src_img)`,
    `caption_consistency_loss(src_reg, tgt_reg)`) with a single all-gather for both."""
    assert trgt_img.shape == src_img.shape and src_reg.shape == tgt_reg.shape
    return _CaptionConsistencyPair.apply(trgt_img, src_img, src_reg, tgt_reg, group)


def kd_l1_loss(teacher: torch.Tensor, student: torch.Tensor) -> torch.Tensor:
    """KD regulariser of the image-level branch, rcnn.py:265-272: `L1Loss()(v2l(offline_backbone(src)).detach(),
    v2l(backbone(src)))` on the [B, 768] V2L features.  One kernel yields the loss and the student's gradient;
    the teacher never receives one."""
    loss, _ = ops.kd_l1(teacher.detach(), student, bool(student.requires_grad and torch.is_grad_enabled()))
    return loss

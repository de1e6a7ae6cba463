"""RegionCLIP pretraining losses: `PretrainFastRCNN.region_concept_matching` and
`image_text_matching` (detectron2/modeling/meta_arch/clip_rcnn.py:583-640), `MILCrossEntropy` and `gather_tensors`
(detectron2/utils/comm.py:268-355).  They share the cosine-logit pattern of the box predictor, so they reuse its
kernels: the logits come from `ops.clip_head_scores` (tcgen05 3xTF32 GEMM for large vocabularies, warp-reduction kernel
otherwise; its autograd gives d/d features), the dense-target softmax losses are one fused row kernel
(`cddmsl_softmax_target_loss`), and the image-text loss is the gathered contrastive kernel with a temperature.
The teacher model that produces the pseudo labels (`get_psuedo_concept_labels`) is out of scope."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import ops


def _cosine_logits(feats: torch.Tensor, embs: torch.Tensor, temperature: float) -> torch.Tensor:
    """normalise(feats) @ normalise(embs).T / temperature, plus the (unused) background column of the head op."""
    bg = torch.zeros((1, embs.shape[1]), dtype=torch.float32, device=embs.device)
    return ops.clip_head_scores(feats, embs.detach(), bg, float(temperature))


def region_concept_distill_loss(keep_region_feats: torch.Tensor, concept_emb: torch.Tensor,
                                concept_scores: torch.Tensor, matching_temp: float) -> torch.Tensor:
    """clip_rcnn.py:592-600 (`loss_region_distill`): KL(teacher distribution || softmax(cos / T)), batchmean.
    keep_region_feats [R, D] (un-normalised), concept_emb [K, D], concept_scores [R, K] teacher probabilities."""
    scores = _cosine_logits(keep_region_feats, concept_emb, matching_temp)
    loss, _ = ops.softmax_target_loss(scores, concept_scores.detach(), ops.TARGET_KL,
                                      bool(scores.requires_grad and torch.is_grad_enabled()))
    return loss


def concept_contrastive_loss(keep_region_feats: torch.Tensor, target_embs: torch.Tensor, label_mtx: torch.Tensor,
                             matching_temp: float) -> torch.Tensor:
    """clip_rcnn.py:602-606 (`loss_concept_contrastive`): MILCrossEntropy(cos(region, target concept) / T, label_mtx)
    with avg_positives=False (comm.py:332-355).  target_embs [R, D], label_mtx [R, R] (1 = same concept)."""
    scores = _cosine_logits(keep_region_feats, target_embs, matching_temp)
    loss, _ = ops.softmax_target_loss(scores, label_mtx.detach().to(torch.float32), ops.TARGET_MIL,
                                      bool(scores.requires_grad and torch.is_grad_enabled()))
    return loss


class _ImageText(torch.autograd.Function):
    @staticmethod
    def forward(ctx, region_feats, text_embs, inv_t, gather, group):
        world = dist.get_world_size(group) if (gather and dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        packed, norms = ops.align_pack(region_feats, text_embs)     # x / |x| per row (clip_rcnn.py:625-626)
        if world > 1:
            packed_all = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(packed_all, packed, group=group)
        else:
            packed_all = packed.unsqueeze(0)
        want = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        # gather_tensors (diffdist all_gather) sums every rank's (identical) gradient into the local rows: x world
        loss, da, db = ops.contrastive_loss(packed_all, norms, rank, inv_t, float(world), want)
        if want:
            ctx.save_for_backward(da, db)
        ctx.have = want
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.have:
            raise RuntimeError("image_text_matching_loss: backward without a recorded forward")
        da, db = ctx.saved_tensors
        return da * g, db * g, None, None, None


def image_text_matching_loss(region_feats: torch.Tensor, text_embs: torch.Tensor, matching_temp: float,
                             gather_gpus: bool = True, group: Optional["dist.ProcessGroup"] = None) -> torch.Tensor:
    """clip_rcnn.py:608-640 (`loss_img_txt_level`): image-level region features against one caption embedding per
    image, normalised, gathered over the ranks (`gather_tensors`, comm.py:268-322: ranks hold equally many images),
    symmetric cross-entropy of `feats @ text.T / matching_temp` against the diagonal."""
    assert region_feats.dim() == 2 and region_feats.shape == text_embs.shape
    return _ImageText.apply(region_feats, text_embs, 1.0 / float(matching_temp), bool(gather_gpus), group)

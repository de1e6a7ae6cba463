"""`batched_nms` / `nms` with the reference's signatures (detectron2/layers/nms.py:19-39 and the
re-exported `torchvision.ops.nms`), backed by csrc/nms.cu.

Mode selection mirrors what the reference's call chain does ON A GPU (where it trains):
`len(boxes) < 40000` -> torchvision `batched_nms`, which applies the coordinate-offset trick when
`boxes.numel() <= COORD_TRICK_NUMEL_LIMIT` and the per-class loop otherwise (torchvision/ops/boxes.py:80-83:
100000 on CUDA, 4000 on CPU -- the oracle tests set the CPU rule explicitly to compare with the CPU oracle);
`>= 40000` -> detectron2's own per-class loop.  Both branches that loop over
classes are the same function of the inputs and map to the class-aware kernel (`coord_trick=False`).
Kept indices come back ordered by score (descending), ties by ascending index — the order of `nms` itself;
the reference's per-class branches re-sort with an unstable sort, so among *exactly equal* scores their
order is unspecified.
"""
from __future__ import annotations

import torch

from .. import ops

# torchvision's CUDA rule (what the reference's training runs); 4000 is its CPU rule (set by the oracle tests).
COORD_TRICK_NUMEL_LIMIT = 100_000


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return ops.batched_nms(boxes, scores, None, float(iou_threshold), False)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                max_keep: int = 0):
    """Same as torchvision.ops.boxes.batched_nms, but safer (fp32 boxes).  `max_keep` > 0 (an extension): the caller
    will slice the result to its first `max_keep` entries anyway (`keep[:topk_per_image]`, fast_rcnn.py:186-187), so
    only that prefix is produced -- bit-identical, the kernels look at the best-scoring candidates first."""
    assert boxes.shape[-1] == 4
    trick = len(boxes) < 40000 and boxes.numel() <= COORD_TRICK_NUMEL_LIMIT
    return ops.batched_nms(boxes.float(), scores, idxs, float(iou_threshold), bool(trick), int(max_keep))


def batched_nms_images(boxes: torch.Tensor, scores: torch.Tensor, idxs, counts: torch.Tensor, iou_threshold: float,
                       max_keep: int = 0, presorted: bool = False):
    """`batched_nms` for the B images of a batch at once (the per-image loop of proposal_utils.py:42-66 in one launch
    sequence, no host sync).  boxes [B,M,4], scores [B,M], idxs [B,M] (or [M], shared by all images) or None,
    counts [B] on the device: image b uses its first counts[b] rows.  Returns (keep [B,M], num_keep [B]); per image
    the result is bit-identical to `batched_nms(boxes[b, :counts[b]], ...)`.
    The class-handling mode follows the same rule as `batched_nms`, evaluated on the padded size M (equal to the
    per-image rule whenever nothing was filtered out; with a single class both modes give identical results).
    `max_keep` > 0: only the first `max_keep` kept boxes of every image are wanted (the `[:post_nms_topk]` of
    proposal_utils.py:116-118): the kernels then look at the best-scoring candidates first and stop early; the
    returned prefix is bit-identical to the full result's.  `presorted`: the caller guarantees non-increasing scores
    per image (the sorted top-k of the RPN path): the internal stable sort is the identity and is skipped."""
    assert boxes.shape[-1] == 4 and boxes.dim() == 3
    nb, m = scores.shape
    if idxs is not None and idxs.dim() == 1:
        idxs = idxs.unsqueeze(0).expand(nb, m)
    trick = m < 40000 and m * 4 <= COORD_TRICK_NUMEL_LIMIT
    return ops.nms_images(boxes.float(), scores, idxs, counts, float(iou_threshold), bool(trick), int(max_keep),
                          bool(presorted))

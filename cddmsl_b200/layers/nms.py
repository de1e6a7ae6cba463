"""`batched_nms` / `nms` with the reference's signatures (detectron2/layers/nms.py:19-39 and the
re-exported `torchvision.ops.nms`), backed by csrc/nms.cu.

Mode selection mirrors what the reference's call chain does on the oracle's device (CPU):
`len(boxes) < 40000` -> torchvision `batched_nms`, which applies the coordinate-offset trick when
`boxes.numel() <= COORD_TRICK_NUMEL_LIMIT` and the per-class loop otherwise (torchvision/ops/boxes.py:80-83:
4000 on CPU, 100000 on CUDA); `>= 40000` -> detectron2's own per-class loop.  Both branches that loop over
classes are the same function of the inputs and map to the class-aware kernel (`coord_trick=False`).
Kept indices come back ordered by score (descending), ties by ascending index — the order of `nms` itself;
the reference's per-class branches re-sort with an unstable sort, so among *exactly equal* scores their
order is unspecified.
"""
from __future__ import annotations

import torch

from .. import ops

# torchvision's CPU rule (the oracle's).  Set to 100_000 to mirror torchvision's CUDA rule instead.
COORD_TRICK_NUMEL_LIMIT = 4000


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return ops.batched_nms(boxes, scores, None, float(iou_threshold), False)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float):
    """Same as torchvision.ops.boxes.batched_nms, but safer (fp32 boxes)."""
    assert boxes.shape[-1] == 4
    trick = len(boxes) < 40000 and boxes.numel() <= COORD_TRICK_NUMEL_LIMIT
    return ops.batched_nms(boxes.float(), scores, idxs, float(iou_threshold), bool(trick))

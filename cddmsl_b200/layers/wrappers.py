"""Small helpers the box predictor uses (detectron2/layers/wrappers.py:16-33, :100-110)."""
from typing import List

import torch
from torch.nn import functional as F


def cat(tensors: List[torch.Tensor], dim: int = 0):
    """torch.cat that skips the copy for a single-element list."""
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def cross_entropy(input, target, *, reduction="mean", **kwargs):
    """F.cross_entropy, but a gradient-connected 0 (not NaN) for empty inputs."""
    if target.numel() == 0 and reduction == "mean":
        return input.sum() * 0.0
    return F.cross_entropy(input, target, reduction=reduction, **kwargs)


def nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)

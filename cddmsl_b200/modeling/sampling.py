"""`subsample_labels` (detectron2/modeling/sampling.py:9-55, this fork draws the negatives' permutation on the CPU)
and its batched, sync-free counterpart used by `label_and_sample_proposals`."""
from __future__ import annotations

from typing import Tuple

import torch

from ..layers import nonzero_tuple


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int):
    """Reference signature and random-number calls (two `randperm`s; the second one on the CPU like upstream)."""
    positive = nonzero_tuple((labels != -1) & (labels != bg_label))[0]
    negative = nonzero_tuple(labels == bg_label)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = torch.randperm(negative.numel())[:num_neg].to(negative.device)
    return positive[perm1], negative[perm2]


def subsample_labels_batched(labels: torch.Tensor, valid: torch.Tensor, num_samples: int, positive_fraction: float,
                             bg_label: int, generator: torch.Generator = None) -> Tuple[torch.Tensor, ...]:
    """The same sampling law for B images at once without data-dependent shapes: every candidate draws a uniform key,
    the `num_pos` smallest keys among the positives and the `num_neg` smallest among the negatives are kept (a uniform
    random subset, exactly what `randperm(n)[:k]` selects).  labels [B,M] (bg_label / -1 / class), valid [B,M] bool.

    Returns (pos_idx int64 [B,P], neg_idx int64 [B,num_samples], num_pos int64 [B], num_neg int64 [B]) with
    P = int(num_samples * positive_fraction): image b keeps pos_idx[b, :num_pos[b]] and neg_idx[b, :num_neg[b]]
    (num_pos = min(#pos, P), num_neg = min(#neg, num_samples - num_pos), sampling.py:44-47).  Everything stays on the
    device; the caller reads the two count vectors once per batch."""
    nb, m = labels.shape
    dev = labels.device
    max_pos = int(num_samples * positive_fraction)
    pos = valid & (labels != -1) & (labels != bg_label)
    neg = valid & (labels == bg_label)
    keys = torch.rand((nb, m), device=dev, generator=generator)
    inf = torch.full_like(keys, float("inf"))
    kp = min(max_pos, m)
    kn = min(num_samples, m)
    pos_idx = torch.topk(torch.where(pos, keys, inf), kp, dim=1, largest=False).indices if kp > 0 else \
        torch.zeros((nb, 0), dtype=torch.int64, device=dev)
    neg_idx = torch.topk(torch.where(neg, keys, inf), kn, dim=1, largest=False).indices if kn > 0 else \
        torch.zeros((nb, 0), dtype=torch.int64, device=dev)
    num_pos = pos.sum(dim=1).clamp(max=max_pos)
    num_neg = torch.minimum(neg.sum(dim=1), num_samples - num_pos)
    return pos_idx, neg_idx, num_pos, num_neg

"""ROIPooler (detectron2/modeling/poolers.py:98-250) over the sm_100a ROIAlign.

The CDDMSL configs use a single level (res4, scale 1/16, "ROIAlignV2" = aligned, 14x14, adaptive sampling:
config/defaults.py:369,423-426); the FPN level assignment (poolers.py:24-58) is kept for interface parity.
Box lists may be the reference's `Boxes` (anything with `.tensor`) or plain Nx4 tensors.
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn

from ..layers import ROIAlign, cat, nonzero_tuple


def _t(b) -> torch.Tensor:
    return b.tensor if hasattr(b, "tensor") else b


def _fmt_box_list(box_tensor: torch.Tensor, batch_index: int) -> torch.Tensor:
    repeated_index = torch.full_like(box_tensor[:, :1], batch_index, dtype=box_tensor.dtype, device=box_tensor.device)
    return cat((repeated_index, box_tensor), dim=1)


def convert_boxes_to_pooler_format(box_lists) -> torch.Tensor:
    """N per-image box lists -> [M,5] = (batch index, x0, y0, x1, y1) (poolers.py:68-95)."""
    return cat([_fmt_box_list(_t(b), i) for i, b in enumerate(box_lists)], dim=0)


def assign_boxes_to_levels(box_lists, min_level: int, max_level: int, canonical_box_size: int,
                           canonical_level: int) -> torch.Tensor:
    """FPN paper eqn. 1 (poolers.py:24-58)."""
    b = cat([_t(x) for x in box_lists])
    box_sizes = torch.sqrt((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))
    lv = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    lv = torch.clamp(lv, min=min_level, max=max_level)
    return lv.to(torch.int64) - min_level


class ROIPooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = output_size
        if pooler_type == "ROIAlign":
            aligned = False
        elif pooler_type == "ROIAlignV2":
            aligned = True
        else:
            raise ValueError(f"Unknown pooler type: {pooler_type} (this build ships ROIAlign / ROIAlignV2)")
        self.level_poolers = nn.ModuleList(
            ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio, aligned=aligned) for s in scales)
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1, "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert 0 <= self.min_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size

    def forward_pair(self, x_a: List[torch.Tensor], x_b: List[torch.Tensor], box_lists):
        """`forward` on two feature lists with the SAME boxes (clip_roi_heads.py:123-128 pools features_src and
        features_trgt with identical proposal boxes): a single level shares one planning pass between the maps."""
        if len(self.level_poolers) != 1 or len(box_lists) == 0 or x_a[0].shape != x_b[0].shape:
            return self.forward(x_a, box_lists), self.forward(x_b, box_lists)
        assert isinstance(x_a, list) and isinstance(x_b, list) and isinstance(box_lists, list)
        assert len(x_a) == 1 and len(x_b) == 1 and len(box_lists) == x_a[0].size(0)
        rois = convert_boxes_to_pooler_format(box_lists)
        return self.level_poolers[0].forward_pair(x_a[0], x_b[0], rois)

    def forward(self, x: List[torch.Tensor], box_lists) -> torch.Tensor:
        num_level_assignments = len(self.level_poolers)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == num_level_assignments, \
            f"unequal value, num_level_assignments={num_level_assignments}, but x is list of {len(x)} Tensors"
        assert len(box_lists) == x[0].size(0), \
            f"unequal value, x[0] batch dim 0 is {x[0].size(0)}, but box_list has length {len(box_lists)}"
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        if num_level_assignments == 1:
            return self.level_poolers[0](x[0], pooler_fmt_boxes)
        level_assignments = assign_boxes_to_levels(box_lists, self.min_level, self.max_level,
                                                   self.canonical_box_size, self.canonical_level)
        output = torch.zeros((pooler_fmt_boxes.size(0), x[0].shape[1]) + self.output_size, dtype=x[0].dtype,
                             device=x[0].device)
        for level, pooler in enumerate(self.level_poolers):
            inds = nonzero_tuple(level_assignments == level)[0]
            output.index_put_((inds,), pooler(x[level], pooler_fmt_boxes[inds]))
        return output

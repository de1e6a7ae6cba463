"""`ROIAlign` with the reference's signature (detectron2/layers/roi_align.py:7-74), backed by the sm_100a
kernels of csrc/roi_align.cu instead of `torchvision.ops.roi_align`."""
from __future__ import annotations

from typing import List, Union

import torch
from torch import nn

from .. import ops


def _boxes_to_rois(boxes: Union[torch.Tensor, List[torch.Tensor]]) -> torch.Tensor:
    """torchvision accepts Tensor[K,5] or a per-image list of Tensor[L,4] (ops/_utils.py convert_boxes_to_roi_format)."""
    if isinstance(boxes, torch.Tensor):
        return boxes
    parts = [torch.cat((torch.full_like(b[:, :1], i), b), dim=1) for i, b in enumerate(boxes)]
    return torch.cat(parts, dim=0)


def roi_align(input: torch.Tensor, boxes, output_size, spatial_scale: float = 1.0, sampling_ratio: int = -1,
              aligned: bool = False) -> torch.Tensor:
    """Same contract as `torchvision.ops.roi_align` (what roi_align.py:58-65 upstream calls)."""
    rois = _boxes_to_rois(boxes)
    assert rois.dim() == 2 and rois.size(1) == 5
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    return ops.roi_align(input, rois.to(dtype=input.dtype), float(spatial_scale), int(output_size[0]),
                         int(output_size[1]), int(sampling_ratio), bool(aligned))


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        """
        Args:
            output_size (tuple): h, w
            spatial_scale (float): scale the input boxes by this number
            sampling_ratio (int): samples per bin and axis; 0 = adaptive ceil(roi_size / output_size)
            aligned (bool): shift box coordinates by -0.5 (pixel-centre model) — the detectron2 default;
                False is the legacy Detectron behaviour (with the 1x1 minimum RoI size).
        """
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        """
        Args:
            input: NCHW images
            rois: Bx5 boxes. First column is the index into N. The other 4 columns are xyxy.
        """
        assert rois.dim() == 2 and rois.size(1) == 5
        if input.is_quantized:
            input = input.dequantize()
        return roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale,
                         self.sampling_ratio, self.aligned)

    def forward_pair(self, input_a, input_b, rois):
        """The same RoIs on two feature maps of the same shape (source / target of the region-level consistency
        branch, clip_roi_heads.py:117-132): one planning pass serves both; equals two `forward` calls bit for bit."""
        assert rois.dim() == 2 and rois.size(1) == 5
        ph, pw = (self.output_size, self.output_size) if isinstance(self.output_size, int) else self.output_size
        return ops.roi_align_pair(input_a, input_b, rois.to(dtype=input_a.dtype), float(self.spatial_scale), int(ph),
                                  int(pw), int(self.sampling_ratio), bool(self.aligned))

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")

"""Drop-in equivalents of the `detectron2.layers` symbols on the hot path (layers/__init__.py:5-6 upstream)."""
from .nms import batched_nms, batched_nms_images, nms
from .roi_align import ROIAlign, roi_align
from .wrappers import cat, cross_entropy, nonzero_tuple

__all__ = ["ROIAlign", "roi_align", "batched_nms", "batched_nms_images", "nms", "cat", "cross_entropy", "nonzero_tuple"]

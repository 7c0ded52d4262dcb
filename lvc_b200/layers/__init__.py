"""Operator API mirroring detectron2/layers/__init__.py:1-13 for the ops on the mining hot path."""
from .nms import batched_nms, nms
from .roi_align import ROIAlign, roi_align
from .wrappers import Conv2d, FrozenBatchNorm2d, Linear, ShapeSpec, cat, get_norm, nonzero_tuple

__all__ = ["batched_nms", "nms", "ROIAlign", "roi_align", "Conv2d", "FrozenBatchNorm2d", "Linear", "ShapeSpec", "cat", "get_norm", "nonzero_tuple"]

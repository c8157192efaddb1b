"""jdet_b200.ops — drop-in mirrors of jdet.ops.* for the oriented-box geometry hot path.

Same re-exports as the reference's python/jdet/ops/__init__.py:1-2; the other ops are reached as
submodules (``from jdet_b200.ops import roi_align_rotated_v1`` -> module), as in the reference.
"""
from .box_iou_rotated import box_iou_rotated
from .box_iou_rotated_v1 import box_iou_rotated_v1
from . import nms_rotated, roi_align_rotated, roi_align_rotated_v1, fr, dcn_v1, orn, bbox_transforms  # noqa: F401

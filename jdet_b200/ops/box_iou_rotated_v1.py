"""jdet.ops.box_iou_rotated_v1 mirror (reference: python/jdet/ops/box_iou_rotated_v1.py:507-525).

Mirrored vertex convention (:69-72) plus the post-pass that zeroes rows/columns of boxes whose
min(w,h) < 1e-3 (:516-523) — fused into the kernel here.
"""
from .box_iou_rotated import _iou


def box_iou_rotated_v1(boxes1, boxes2):
    return _iou(boxes1, boxes2, 1)

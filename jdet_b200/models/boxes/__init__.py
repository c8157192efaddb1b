from .iou_calculator import BboxOverlaps2D_rotated, BboxOverlaps2D_rotated_v1, bbox_overlaps_rotated  # noqa: F401

from .iou_calculator import BboxOverlaps2D_rotated, BboxOverlaps2D_rotated_v1, bbox_overlaps_rotated  # noqa: F401
from .box_ops import delta2bbox_rotated, norm_angle, rotated_box_to_poly  # noqa: F401
from .anchor_generator import AnchorGeneratorRotatedS2ANet  # noqa: F401
from .anchor_generator import AnchorGenerator  # noqa: F401
from .coder import (midpoint_offset_decode, oriented_delta_xywht_decode, obb2hbb, obb2poly, rectpoly2obb,  # noqa: F401
                    regular_obb, regular_theta)

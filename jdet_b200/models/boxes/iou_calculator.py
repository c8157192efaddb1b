"""Rotated IoU calculators (reference: python/jdet/models/boxes/iou_calculator.py:121-233)."""
from ...ops import box_iou_rotated, box_iou_rotated_v1


def bbox_overlaps_rotated(rboxes1, rboxes2, version=0):
    if version == 0:
        return box_iou_rotated(rboxes1.float(), rboxes2.float())
    return box_iou_rotated_v1(rboxes1.float(), rboxes2.float())


class BboxOverlaps2D_rotated:
    """bboxes (m,5|6) x (n,5|6) -> (m,n) IoU; a 6th (score) column is dropped."""
    _version = 0

    def __call__(self, bboxes1, bboxes2, mode='iou', is_aligned=False):
        assert bboxes1.size(-1) in [0, 5, 6]
        assert bboxes2.size(-1) in [0, 5, 6]
        if bboxes2.size(-1) == 6:
            bboxes2 = bboxes2[..., :5]
        if bboxes1.size(-1) == 6:
            bboxes1 = bboxes1[..., :5]
        assert mode == "iou" and is_aligned is False
        return bbox_overlaps_rotated(bboxes1, bboxes2, version=self._version)

    def __repr__(self):
        return self.__class__.__name__ + '()'


class BboxOverlaps2D_rotated_v1(BboxOverlaps2D_rotated):
    _version = 1

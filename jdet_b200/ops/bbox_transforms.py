"""jdet.ops.bbox_transforms — the box-format helpers on either side of the geometry path (SURVEY 8f rank 4:
detections leave the heads as oriented boxes and are written out / merged as polygons or horizontal boxes).

Mirrors python/jdet/ops/bbox_transforms.py:499-704 by name and behaviour, batched torch, any device:
regular_theta / regular_obb (:499-517), get_bbox_type / get_bbox_dim (:519-545), rectpoly2obb (:575-597),
poly2hbb (:600-607), obb2poly (:610-636), obb2hbb (:639-645), hbb2poly (:648-650), hbb2obb (:653-666),
bbox2type (:676-685), get_bbox_areas (:687-704).  `poly2obb` (:547-573) is cv2.minAreaRect in the reference — a
third-party arithmetic that is not under /root/reference; rectangles (what the detectors emit) go through
rectpoly2obb, general polygons raise.
"""
import math

import torch

from ..models.boxes.coder import obb2hbb, obb2poly, rectpoly2obb, regular_obb, regular_theta  # noqa: F401


def get_bbox_type(bboxes, with_score=False):
    dim = bboxes.shape[-1] - (1 if with_score else 0)
    return {4: "hbb", 5: "obb", 8: "poly"}.get(dim, "notype")


def get_bbox_dim(bbox_type, with_score=False):
    dims = {"hbb": 4, "obb": 5, "poly": 8}
    if bbox_type not in dims:
        raise ValueError(f"don't know {bbox_type} bbox dim")
    return dims[bbox_type] + (1 if with_score else 0)


def poly2hbb(polys):
    pts = polys.reshape(*polys.shape[:-1], polys.shape[-1] // 2, 2)
    return torch.cat([pts.min(-2)[0], pts.max(-2)[0]], -1)


def hbb2poly(hbboxes):
    l, t, r, b = hbboxes.unbind(-1)
    return torch.stack([l, t, r, t, r, b, l, b], -1)


def hbb2obb(hbboxes):
    """long side first: a box taller than wide becomes (h, w, -pi/2)."""
    l, t, r, b = hbboxes.unbind(-1)
    x, y, w, h = (l + r) * 0.5, (t + b) * 0.5, r - l, b - t
    wide = w >= h
    zero = torch.zeros_like(x)
    return torch.stack([x, y, torch.where(wide, w, h), torch.where(wide, h, w),
                        torch.where(wide, zero, zero - math.pi / 2)], -1)


def poly2obb(polys):
    raise NotImplementedError("poly2obb is cv2.minAreaRect in the reference (ops/bbox_transforms.py:547-573); "
                              "use rectpoly2obb for rectangles")


_CONVERT = {("poly", "obb"): poly2obb, ("poly", "hbb"): poly2hbb, ("obb", "poly"): obb2poly,
            ("obb", "hbb"): obb2hbb, ("hbb", "poly"): hbb2poly, ("hbb", "obb"): hbb2obb}


def bbox2type(bboxes, to_type):
    assert to_type in ["hbb", "obb", "poly"]
    src = get_bbox_type(bboxes)
    if src == "notype":
        raise ValueError("Not a bbox type")
    return bboxes if src == to_type else _CONVERT[(src, to_type)](bboxes)


def get_bbox_areas(bboxes):
    kind = get_bbox_type(bboxes)
    if kind == "hbb":
        return (bboxes[..., 2] - bboxes[..., 0]) * (bboxes[..., 3] - bboxes[..., 1])
    if kind == "obb":
        return bboxes[..., 2] * bboxes[..., 3]
    if kind == "poly":   # shoelace over the 4 corners
        pts = bboxes.reshape(*bboxes.shape[:-1], 4, 2)
        prev = torch.roll(pts, 1, dims=-2)
        return 0.5 * (pts[..., 0] * prev[..., 1] - prev[..., 0] * pts[..., 1]).sum(-1).abs()
    raise ValueError("The type of bboxes is notype")

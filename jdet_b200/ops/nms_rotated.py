"""jdet.ops.nms_rotated mirror (reference: python/jdet/ops/nms_rotated.py:495-596).

All entry points keep the reference names, argument order and return conventions:
  nms_rotated(dets, scores, iou_threshold)                     -> kept indices, ascending   (:527-538)
  ml_nms_rotated(dets, scores, labels, iou_threshold)          -> kept indices, ascending   (:515-525)
  multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1,
                         score_factors=None)                   -> ((k,6) dets, (k,) labels) (:540-596)
  nms_rotated_cuda(dets, order_t, iou_threshold, box_length=6) -> (n,) bool keep mask       (:506-513)
Indices are int64 (torch idiom; the reference returns int32 — same values).
"""
import torch

from ._common import check, f32c, lib, require_cuda, scratch, stream_ptr


def argsort_desc(scores):
    """Stable descending argsort on the device (ties keep the lower index first), int32."""
    require_cuda(scores)
    s = f32c(scores).reshape(-1)
    n = s.numel()
    order = torch.empty((n,), dtype=torch.int32, device=s.device)
    if n == 0:
        return order
    L = lib()
    with torch.cuda.device(s.device):
        ws = scratch(L.jdet_argsort_desc_workspace_bytes(n), s.device)
        check(L.jdet_argsort_desc(s.data_ptr(), n, order.data_ptr(), ws.data_ptr(), ws.numel(),
                                  stream_ptr(s.device)), "argsort_desc")
    return order


def _nms_keep(dets, order_t, iou_threshold, box_length, convention):
    require_cuda(dets, order_t)
    d = f32c(dets)
    assert d.dim() == 2 and d.shape[1] == box_length and box_length in (5, 6)
    n = d.shape[0]
    keep = torch.empty((n,), dtype=torch.bool, device=d.device)
    if n == 0:
        return keep
    order = order_t.to(torch.int32).contiguous()
    assert order.numel() == n
    L = lib()
    with torch.cuda.device(d.device):
        ws = scratch(L.jdet_nms_rotated_workspace_bytes(n, box_length), d.device)
        check(L.jdet_nms_rotated_ex(d.data_ptr(), n, box_length, order.data_ptr(), float(iou_threshold), convention,
                                    keep.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(d.device)), "nms_rotated")
    return keep


def nms_rotated_cuda(dets, order_t, iou_threshold, box_length=6):
    """keep mask (n,) bool.  dets (n, box_length): [x,y,w,h,theta(,label)]; order_t: indices by
    descending score.  Strict `IoU > iou_threshold`, as the reference CUDA kernel (:403-404)."""
    return _nms_keep(dets, order_t, iou_threshold, box_length, 1)


def nms_rotated_cpu(dets, order_t, iou_threshold, box_length=6):
    """The reference's CPU path (ops/nms_rotated.py:495-504, loop :414-449) — suppress on `IoU >= thr`, IoU with the
    CPU build's std::sort hull — evaluated on the GPU, bit for bit (jdet_nms_rotated_ex, convention 0).  The two
    reference builds can disagree by far more than rounding on the same pair (the CPU build reads stale distances after
    its sort, box_iou_rotated.py:219-224), so the arithmetic is selectable, not just the comparison.  No host path."""
    return _nms_keep(dets, order_t, iou_threshold, box_length, 0)


def py_cpu_nms_obb(dets, thresh):
    """Tile -> image merge NMS (data/devkits/result_merge.py:132-145): dets (n,9) = 4 corner points + score ->
    indices kept, ascending.  The reference converts polygons with cv2.minAreaRect; detections here are rectangles
    produced by obb2poly / rotated_box_to_poly, so the closed-form rectpoly2obb is used instead."""
    from ..models.boxes.coder import rectpoly2obb
    if dets.numel() == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dets.device)
    boxes = rectpoly2obb(dets[:, :8].float())
    order = argsort_desc(dets[:, 8].float().contiguous())
    return torch.where(nms_rotated_cpu(boxes, order, thresh, box_length=5))[0]


def _nms_poly(dets, thresh, fast):
    require_cuda(dets)
    if dets.numel() == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dets.device)
    d = f32c(dets)
    assert d.dim() == 2 and d.shape[1] == 9, "dets must be (n, 9): 4 corner points + score"
    n = d.shape[0]
    order = argsort_desc(d[:, 8].contiguous())
    keep = torch.empty((n,), dtype=torch.bool, device=d.device)
    L = lib()
    with torch.cuda.device(d.device):
        ws = scratch(L.jdet_nms_poly_workspace_bytes(n), d.device)
        check(L.jdet_nms_poly(d.data_ptr(), n, order.data_ptr(), float(thresh), int(fast), keep.data_ptr(), ws.data_ptr(),
                              ws.numel(), stream_ptr(d.device)), "nms_poly")
    kept = order.long()[keep[order.long()]]            # kept indices in descending score order, like the reference's `keep` list
    return kept


def py_cpu_nms_poly_fast(dets, thresh):
    """Tile -> image merge NMS over quadrilaterals (data/devkits/result_merge.py:69-131): dets (n,9) = 4 corner points + score
    -> kept indices in descending score order (the order the reference appends them).  Bounding-box pre-filter, then
    iou_poly (ops/nms_poly.py:247-252) on the survivors, suppress on iou > thresh — on the GPU, in binary64."""
    return _nms_poly(dets, thresh, True)


def py_cpu_nms_poly(dets, thresh):
    """result_merge.py:33-66: the same without the bounding-box pre-filter."""
    return _nms_poly(dets, thresh, False)


def ml_nms_rotated(dets, scores, labels, iou_threshold):
    assert dets.numel() > 0 and dets.dim() == 2              # nms_rotated.py:516
    assert dets.dtype == scores.dtype                         # :517
    require_cuda(dets, scores, labels)
    d = torch.cat([f32c(dets), labels.to(torch.float32).unsqueeze(1)], dim=1)
    order_t = argsort_desc(scores)
    keep = nms_rotated_cuda(d, order_t, iou_threshold, box_length=6)
    return torch.where(keep)[0]


def ml_nms_rotated_record(dets, scores, labels, iou_threshold, max_per_img=2000, out=None):
    """ml_nms_rotated + the tail of multiclass_nms_rotated (re-sort by score, [:max_per_img], :584-596) as a fixed-size
    record (max_per_img + 1, 7): rows [x,y,w,h,theta,score,label] of the kept boxes in descending score order, zero
    padded, last row = count.  Three library calls (argsort, NMS, pack), no host sync, no index kernels; `out` may be a
    slice of a persistent send buffer (jdet_b200.dist.all_gather_records)."""
    require_cuda(dets, scores, labels)
    if out is None:
        out = torch.empty((max_per_img + 1, 7), dtype=torch.float32, device=dets.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (max_per_img + 1, 7)
    n = dets.shape[0]
    L = lib()
    if n == 0:
        out.zero_()
        return out
    assert dets.dim() == 2 and dets.dtype == scores.dtype
    d = torch.cat([f32c(dets), labels.to(torch.float32).unsqueeze(1)], dim=1)
    s = f32c(scores).reshape(-1)
    order_t = argsort_desc(s)
    keep = nms_rotated_cuda(d, order_t, iou_threshold, box_length=6)
    with torch.cuda.device(d.device):
        check(L.jdet_pack_detections(d.data_ptr(), n, 6, s.data_ptr(), order_t.data_ptr(), keep.data_ptr(), max_per_img,
                                     out.data_ptr(), stream_ptr(d.device)), "pack_detections")
    return out


def nms_rotated(dets, scores, iou_threshold):
    if dets.numel() == 0:                                     # :528-529
        return torch.zeros((0,), dtype=torch.int64, device=dets.device)
    assert dets.numel() > 0 and dets.dim() == 2
    assert dets.dtype == scores.dtype
    require_cuda(dets, scores)
    order_t = argsort_desc(scores)
    keep = nms_rotated_cuda(f32c(dets), order_t, iou_threshold, box_length=5)
    return torch.where(keep)[0]


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None):
    """NMS for multi-class rotated boxes; column 0 of multi_scores is background and ignored.
    Returns ((k,6) [box, score], (k,) 0-based labels).  `max_num=-1` drops the lowest-scoring
    detection, exactly like the reference (:590-591)."""
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 5:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 5)[:, 1:]
    else:
        bboxes = multi_bboxes[:, None].expand(multi_bboxes.shape[0], num_classes, 5)
    scores = multi_scores[:, 1:]
    valid_mask = scores > score_thr
    bboxes = bboxes[valid_mask]
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    scores = scores[valid_mask]
    labels = valid_mask.nonzero()[:, 1]
    if bboxes.numel() == 0:
        return (torch.zeros((0, 6), device=multi_bboxes.device),
                torch.zeros((0,), dtype=torch.int32, device=multi_bboxes.device))
    nms_cfg_ = nms_cfg.copy()
    nms_cfg_.pop('type', 'nms')
    iou_thr = nms_cfg_.pop('iou_thr', 0.1)
    keep = ml_nms_rotated(bboxes, scores, labels, iou_thr)
    bboxes, scores, labels = bboxes[keep], scores[keep], labels[keep]
    inds = torch.argsort(scores, descending=True, stable=True)
    if keep.size(0) > max_num:
        inds = inds[:max_num]
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    return torch.cat([bboxes, scores[:, None]], 1), labels

"""jdet.ops.box_iou_rotated mirror (reference: python/jdet/ops/box_iou_rotated.py:502-509)."""
import torch

from ._common import check, f32c, lib, require_cuda, scratch, stream_ptr


def _iou(boxes1, boxes2, version, arithmetic=1):
    assert boxes1.dtype == boxes2.dtype                       # box_iou_rotated.py:503
    require_cuda(boxes1, boxes2)
    assert boxes1.dim() == 2 and boxes2.dim() == 2 and boxes1.shape[1] == 5 and boxes2.shape[1] == 5, \
        "boxes must be (N,5) [x_ctr, y_ctr, w, h, theta(rad)]"
    b1, b2 = f32c(boxes1), f32c(boxes2)
    n1, n2 = b1.shape[0], b2.shape[0]
    out = torch.empty((n1, n2), dtype=torch.float32, device=b1.device)
    if n1 == 0 or n2 == 0:
        return out
    L = lib()
    with torch.cuda.device(b1.device):
        nbytes = L.jdet_box_iou_rotated_workspace_bytes(n1, n2)
        ws = scratch(nbytes, b1.device)
        check(L.jdet_box_iou_rotated_ex(b1.data_ptr(), n1, b2.data_ptr(), n2, out.data_ptr(), version, arithmetic,
                                        ws.data_ptr(), ws.numel(), stream_ptr(b1.device)), "box_iou_rotated")
    return out


def box_iou_rotated(boxes1, boxes2, cpu_arithmetic=False):
    """Pairwise IoU of rotated boxes: (N,5),(M,5) -> (N,M) fp32.  cpu_arithmetic=True reproduces the reference's CPU
    build (the cpu_src branch of box_iou_rotated.py:504-509, std::sort hull) instead of its CUDA build — on the GPU."""
    return _iou(boxes1, boxes2, 0, 0 if cpu_arithmetic else 1)

"""jdet.data.devkits.result_merge mirror (reference: python/jdet/data/devkits/result_merge.py): the tile -> image merge.

A DOTA result file holds one detection per line, `<tile name> <score> <x1 y1 ... x4 y4>`, the tile name carrying the source
image, the scale and the tile's offset: `<image>__<rate>__<x>___<y>` (:216-229).  Merging moves every detection back into
image coordinates (`poly2origpoly`, :193-200), groups by image and runs one NMS per image (`nmsbynamedict`, :174-190); the
survivors are written as `<image> <score> <8 coordinates>` (:250-257).

Same function names and file formats; what differs from the reference:
  * the NMS functions run on the GPU (`jdet_nms_poly`, `jdet_nms_rotated_ex`) instead of shapely / a Python loop, one call per
    image; `mergebase_parallel` needs no process pool and is `mergebase`;
  * the per-class threshold switch the reference reads from its global config (`cfg.merge_nms_threshold_type`, :209-214) is
    the keyword `nms_threshold_type`;
  * coordinates go through float32 on their way to the GPU (the reference keeps Python floats): a pair whose IoU sits
    within ~1e-6 of the threshold can be decided differently.
"""
import os
import re

import numpy as np
import torch

from . import dota_utils as util

# the thresholds of the reference (:25-31): one for all classes, or one per class
nms_threshold_0 = 0.1
nms_threshold_1 = {'roundabout': 0.1, 'tennis-court': 0.3, 'swimming-pool': 0.1, 'storage-tank': 0.2,
                   'soccer-ball-field': 0.3, 'small-vehicle': 0.2, 'ship': 0.2, 'plane': 0.3,
                   'large-vehicle': 0.1, 'helicopter': 0.2, 'harbor': 0.0001, 'ground-track-field': 0.3,
                   'bridge': 0.0001, 'basketball-court': 0.3, 'baseball-diamond': 0.3,
                   'container-crane': 0.05, 'airport': 0.1, 'helipad': 0.1}

_OFFSET = re.compile(r'__(\d+)___(\d+)')
_RATE = re.compile(r'__([\d+\.]+)__\d+___')


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("jdet_b200 result_merge: the merge NMS runs on the GPU; no CUDA device is available (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_gpu(dets):
    """numpy (what nmsbynamedict hands over) or torch in; (device tensor, came-from-numpy)."""
    if isinstance(dets, torch.Tensor):
        return dets, False
    arr = np.asarray(dets, dtype=np.float32)
    if arr.ndim != 2:
        arr = arr.reshape(0, 9)
    return torch.as_tensor(np.ascontiguousarray(arr), device=_device()), True


def _back(keep, was_numpy):
    return keep.cpu().numpy() if was_numpy else keep


def py_cpu_nms_poly_fast(dets, thresh):
    """(n, 9) polygons + score -> kept indices in descending score order (:69-131)."""
    from ...ops.nms_rotated import py_cpu_nms_poly_fast as f
    d, was = _to_gpu(dets)
    return _back(f(d, thresh), was)


def py_cpu_nms_poly(dets, thresh):
    """The same without the bounding-box pre-filter (:33-66)."""
    from ...ops.nms_rotated import py_cpu_nms_poly as f
    d, was = _to_gpu(dets)
    return _back(f(d, thresh), was)


def py_cpu_nms_obb(dets, thresh):
    """(n, 9) rectangles as polygons + score -> kept indices, ascending (:132-145)."""
    from ...ops.nms_rotated import py_cpu_nms_obb as f
    d, was = _to_gpu(dets)
    return _back(f(d, thresh), was)


def py_cpu_nms(dets, thresh):
    """(n, 5) [x1, y1, x2, y2, score] -> kept indices in descending score order (:147-178).  The reference's areas and
    overlaps count pixels (`x2 - x1 + 1`): that is the ordinary IoU of the boxes grown by one pixel to the right and bottom,
    which is what goes to the library's NMS (theta = 0)."""
    from ...ops.nms_rotated import argsort_desc, nms_rotated_cuda
    d, was = _to_gpu(dets)
    if d.numel() == 0:
        return _back(torch.zeros((0,), dtype=torch.int64, device=d.device), was)
    d = d.float()
    w, h = d[:, 2] - d[:, 0] + 1, d[:, 3] - d[:, 1] + 1
    boxes = torch.stack([d[:, 0] + 0.5 * w, d[:, 1] + 0.5 * h, w, h, torch.zeros_like(w)], 1).contiguous()
    order = argsort_desc(d[:, 4].contiguous())
    keep = nms_rotated_cuda(boxes, order, thresh, box_length=5)
    o = order.long()
    return _back(o[keep[o]], was)


def nmsbynamedict(nameboxdict, nms, thresh):
    """image -> detections  =>  image -> the detections `nms` keeps, in the order it returns them (:174-190)."""
    out = {}
    for imgname, boxes in nameboxdict.items():
        keep = nms(np.array(boxes), thresh)
        out[imgname] = [boxes[int(i)] for i in keep]
    return out


def poly2origpoly(poly, x, y, rate):
    """tile coordinates -> image coordinates: (p + tile offset) / scale (:193-200)."""
    rate = float(rate)
    off = (x, y)
    return [float(v + off[i & 1]) / rate for i, v in enumerate(poly)]


def parse_result_file(fullname):
    """A result file -> {image: [[x1, y1, ..., x4, y4, score], ...]} in image coordinates (:216-245)."""
    nameboxdict = {}
    with open(fullname, 'r') as f:
        for line in f:
            parts = line.strip().split(' ')
            if len(parts) < 3:
                continue
            subname = parts[0]
            oriname = subname.split('__')[0]
            m = _OFFSET.search(subname)
            x, y = int(m.group(1)), int(m.group(2))
            rate = _RATE.search(subname).group(1)
            det = poly2origpoly([float(v) for v in parts[2:]], x, y, rate)
            det.append(float(parts[1]))
            nameboxdict.setdefault(oriname, []).append(det)
    return nameboxdict


def mergesingle(dstpath, nms, fullname, nms_threshold_type=0):
    """One class file: parse, merge, NMS per image, write `<dstpath>/<same name>.txt` (:203-257)."""
    name = util.custombasename(fullname)
    dstname = os.path.join(dstpath, name + '.txt')
    nameboxdict = parse_result_file(fullname)
    thresh = nms_threshold_0 if nms_threshold_type == 0 else nms_threshold_1[name]
    merged = nmsbynamedict(nameboxdict, nms, thresh)
    with open(dstname, 'w') as f:
        for imgname, dets in merged.items():
            for det in dets:
                f.write(imgname + ' ' + str(det[-1]) + ' ' + ' '.join(map(str, det[:-1])) + '\n')


def mergebase(srcpath, dstpath, nms, nms_threshold_type=0):
    for filename in util.GetFileFromThisRootDir(srcpath):
        mergesingle(dstpath, nms, filename, nms_threshold_type)


mergebase_parallel = mergebase        # (:260-267 uses a 16-process pool around the Python NMS; the GPU NMS needs none)


def mergebyrec(srcpath, dstpath, nms_threshold_type=0):
    """Horizontal boxes (:281-292)."""
    mergebase(srcpath, dstpath, py_cpu_nms, nms_threshold_type)


def mergebypoly(srcpath, dstpath, nms_threshold_type=0):
    """Quadrilaterals (:295-308)."""
    mergebase_parallel(srcpath, dstpath, py_cpu_nms_poly_fast, nms_threshold_type)


def mergebyobb(srcpath, dstpath, nms_threshold_type=0):
    """Oriented rectangles (:310-322)."""
    mergebase(srcpath, dstpath, py_cpu_nms_obb, nms_threshold_type)

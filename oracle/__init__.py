"""oracle — CPU restatement of the reference hot path + loaders for the compiled reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  ``jdet_b200`` never imports this package.

* ``oracle.lib()``      -> ctypes handle of liboracle.so (oracle/oracle.cpp, my restatement)
* ``oracle.ref_cpu()``  -> oracle/_ref/libref_cpu.so  (reference cpu_src compiled by g++) or None
* ``oracle.ref_cuda()`` -> oracle/_ref/libref_cuda.so (reference CUDA kernels, nvcc sm_100a) or None
* numpy-level wrappers: box_iou_rotated, nms_rotated, ml_nms_rotated, multiclass_nms_rotated,
  roi_align_rotated, feature_refine, align_conv_offset, deform_conv, align_conv

Parity status: PINNED (see header of oracle.cpp and tests/test_oracle_pin.py).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.cpp")
SO = os.path.join(HERE, "_build", "liboracle.so")
VARIANT_CPU, VARIANT_CUDA = 0, 1

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u8p = ctypes.POINTER(ctypes.c_ubyte)


def build(force=False):
    """Compile oracle.cpp -> oracle/_build/liboracle.so (g++, OpenMP, no FMA contraction)."""
    if not force and os.path.exists(SO) and os.path.getmtime(SO) > os.path.getmtime(SRC):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           SRC, "-o", SO]
    subprocess.check_call(cmd)
    return SO


_lib = None
_ref_cpu = None
_ref_cuda = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        L.orc_single_iou.restype = ctypes.c_float
        L.orc_single_iou.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _load(name):
    p = os.path.join(HERE, "_ref", name)
    return ctypes.CDLL(p) if os.path.exists(p) else None


def ref_cpu():
    global _ref_cpu
    if _ref_cpu is None:
        _ref_cpu = _load("libref_cpu.so")
        if _ref_cpu is not None:
            for n in ("ref_single_iou_v0_cpu", "ref_single_iou_v1_cpu", "ref_single_iou_nms5_cpu",
                      "ref_single_iou_nms6_cpu"):
                getattr(_ref_cpu, n).restype = ctypes.c_float
    return _ref_cpu


def ref_cuda():
    """Needs libcudart at load time; returns None where it cannot be loaded."""
    global _ref_cuda
    if _ref_cuda is None:
        try:
            _ref_cuda = _load("libref_cuda.so")
        except OSError:
            _ref_cuda = None
        if _ref_cuda is not None:
            for n in ("ref_single_iou_v0_cudavariant_host", "ref_single_iou_v1_cudavariant_host",
                      "ref_single_iou_nms5_cudavariant_host", "ref_single_iou_nms6_cudavariant_host"):
                getattr(_ref_cuda, n).restype = ctypes.c_float
    return _ref_cuda


def max_threads():
    return int(lib().orc_max_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_f32p)


# ---------------------------------------------------------------------------------------------
def box_iou_rotated(boxes1, boxes2, version=0, variant=VARIANT_CUDA, threads=0):
    """jdet.ops.box_iou_rotated / box_iou_rotated_v1 (ops/box_iou_rotated.py:502-509, _v1.py:507-525)."""
    b1, b2 = _f32(boxes1).reshape(-1, 5), _f32(boxes2).reshape(-1, 5)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    if out.size:
        lib().orc_box_iou_rotated(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out), version, variant,
                                  threads)
        if version == 1:
            lib().orc_box_iou_rotated_v1_postzero(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out))
    return out


def single_iou(a, b, version=0, variant=VARIANT_CUDA, box_length=5):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_single_iou(_p(a), _p(b), version, variant, box_length))


def nms_rotated_keep(dets, order, thr, variant=VARIANT_CUDA):
    """keep mask (n,) bool — nms_rotated_cuda / nms_rotated_cpu (ops/nms_rotated.py:495-513)."""
    d = _f32(dets)
    n, bl = d.shape
    order = np.ascontiguousarray(order, dtype=np.int32)
    keep = np.zeros(n, np.uint8)
    if n:
        lib().orc_nms_rotated(_p(d), n, bl, order.ctypes.data_as(_i32p), ctypes.c_float(np.float32(thr)),
                              variant, keep.ctypes.data_as(_u8p))
    return keep.astype(bool)


def argsort_desc(scores):
    """Descending, ties by ascending index (== torch.argsort(descending=True, stable=True))."""
    s = np.asarray(scores, np.float32)
    return np.argsort(-s.astype(np.float64), kind="stable").astype(np.int32)


def ml_nms_rotated(dets, scores, labels, thr, variant=VARIANT_CUDA):
    """ops/nms_rotated.py:515-525 -> kept indices, ascending."""
    d = np.concatenate([_f32(dets), np.asarray(labels, np.float32)[:, None]], 1)
    keep = nms_rotated_keep(d, argsort_desc(scores), thr, variant)
    return np.nonzero(keep)[0]


def nms_rotated(dets, scores, thr, variant=VARIANT_CUDA):
    """ops/nms_rotated.py:527-538"""
    d = _f32(dets)
    if d.size == 0:
        return np.zeros((0,), np.int64)
    keep = nms_rotated_keep(d, argsort_desc(scores), thr, variant)
    return np.nonzero(keep)[0]


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1, score_factors=None,
                           variant=VARIANT_CUDA):
    """ops/nms_rotated.py:540-596 (including the max_num=-1 quirk that drops the last detection)."""
    mb, ms = _f32(multi_bboxes), _f32(multi_scores)
    ncls = ms.shape[1] - 1
    if mb.shape[1] > 5:
        bboxes = mb.reshape(ms.shape[0], -1, 5)[:, 1:]
    else:
        bboxes = np.broadcast_to(mb[:, None], (mb.shape[0], ncls, 5))
    scores = ms[:, 1:]
    valid = scores > np.float32(score_thr)
    bboxes = bboxes[valid]
    if score_factors is not None:
        scores = scores * _f32(score_factors)[:, None]
    scores = scores[valid]
    labels = np.nonzero(valid)[1]
    if bboxes.size == 0:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int32)
    keep = ml_nms_rotated(bboxes, scores, labels, nms_cfg.get("iou_thr", 0.1), variant)
    bboxes, scores, labels = bboxes[keep], scores[keep], labels[keep]
    inds = argsort_desc(scores)
    if keep.shape[0] > max_num:
        inds = inds[:max_num]
    return np.concatenate([bboxes[inds], scores[inds][:, None]], 1), labels[inds].astype(np.int32)


def roi_align_rotated(input, rois, output_size, spatial_scale, sampling_ratio=0, version=1, threads=0):
    x, r = _f32(input), _f32(rois).reshape(-1, 6)
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    B, C, H, W = x.shape
    out = np.zeros((r.shape[0], C, ph, pw), np.float32)
    if out.size:
        lib().orc_roi_align_rotated(version, _p(x), _p(r), r.shape[0], C, H, W, ph, pw,
                                    ctypes.c_float(np.float32(spatial_scale)), int(sampling_ratio),
                                    _p(out), threads)
    return out


def roi_align_rotated_backward(grad_out, rois, input_shape, spatial_scale, sampling_ratio=0, version=1):
    g, r = _f32(grad_out), _f32(rois).reshape(-1, 6)
    B, C, H, W = input_shape
    R, _, ph, pw = g.shape
    out = np.zeros((B, C, H, W), np.float32)
    lib().orc_roi_align_rotated_backward(version, _p(g), _p(r), R, B, C, H, W, ph, pw,
                                         ctypes.c_float(np.float32(spatial_scale)), int(sampling_ratio), _p(out))
    return out


def feature_refine_backward(grad_out, best_rbboxes, spatial_scale, points=1):
    g = _f32(grad_out)
    N, C, H, W = g.shape
    b = _f32(best_rbboxes).reshape(N, H, W, 5)
    out = np.zeros_like(g)
    lib().orc_feature_refine_backward(_p(g), _p(b), N, C, H, W, points, ctypes.c_float(np.float32(spatial_scale)), _p(out))
    return out


def feature_refine(features, best_rbboxes, spatial_scale, points=1, threads=0):
    x = _f32(features)
    N, C, H, W = x.shape
    b = _f32(best_rbboxes).reshape(N, H, W, 5)
    out = np.zeros_like(x)
    if out.size:
        lib().orc_feature_refine(_p(x), _p(b), N, C, H, W, points, ctypes.c_float(np.float32(spatial_scale)),
                                 _p(out), threads)
    return out


def align_conv_offset(anchors, stride, kernel_size=3):
    """anchors (N,H,W,5) -> offsets (N, 2*k*k, H, W)   (s2anet_head.py:677-721)"""
    a = _f32(anchors)
    N, H, W, _ = a.shape
    off = np.zeros((N, 2 * kernel_size * kernel_size, H, W), np.float32)
    for i in range(N):
        lib().orc_align_conv_offset(_p(a[i]), H, W, ctypes.c_float(np.float32(stride)), kernel_size, _p(off[i]))
    return off


def deform_conv(x, offset, weight, stride=1, padding=0, dilation=1, deformable_groups=1, relu=False,
                threads=0):
    """DeformConv v1 forward, groups=1 (ops/dcn_v1.py:412-454)."""
    x, offset, weight = _f32(x), _f32(offset), _f32(weight)
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    pr = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    (sh, sw), (ph, pw), (dh, dw) = pr(stride), pr(padding), pr(dilation)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = np.zeros((B, Co, Ho, Wo), np.float32)
    lib().orc_deform_conv(_p(x), _p(offset), _p(weight), B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw,
                          deformable_groups, int(relu), _p(out), threads)
    return out


def deform_conv_backward(x, offset, weight, grad_out, stride=1, padding=0, dilation=1, deformable_groups=1):
    """-> (grad_x, grad_offset, grad_weight), groups = 1 (ops/dcn_v1.py:457-556)."""
    x, offset, weight, grad_out = _f32(x), _f32(offset), _f32(weight), _f32(grad_out)
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    gx, go, gw = np.zeros_like(x), np.zeros_like(offset), np.zeros_like(weight)
    lib().orc_deform_conv_backward(_p(x), _p(offset), _p(weight), _p(grad_out), B, C, H, W, Co, kh, kw, stride, stride,
                                   padding, padding, dilation, dilation, deformable_groups, _p(gx), _p(go), _p(gw))
    return gx, go, gw


def deform_im2col(x, offset, kh, kw, stride=1, padding=0, dilation=1, deformable_groups=1):
    x, offset = _f32(x), _f32(offset)
    B, C, H, W = x.shape
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    col = np.zeros((C * kh * kw, B, Ho, Wo), np.float32)
    lib().orc_deform_im2col(_p(x), _p(offset), B, C, H, W, kh, kw, stride, stride, padding, padding,
                            dilation, dilation, deformable_groups, _p(col))
    return col


def align_conv(x, anchors, stride, weight, threads=0):
    """AlignConv.execute (s2anet_head.py:715-723): offsets -> DeformConv(3x3,pad 1) -> ReLU."""
    k = weight.shape[-1]
    off = align_conv_offset(anchors, stride, k)
    return deform_conv(x, off, weight, 1, (k - 1) // 2, 1, 1, relu=True, threads=threads)

"""torch-side wrappers over oracle/_ref/libref_cuda.so (the reference's CUDA kernels compiled for
sm_100a).  Test infrastructure; returns None-able handles so tests can skip when it is absent."""
import ctypes

import numpy as np
import torch

import oracle


def available():
    return torch.cuda.is_available() and oracle.ref_cuda() is not None


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def box_iou_rotated(b1, b2, version=0):
    R = oracle.ref_cuda()
    out = torch.zeros((b1.shape[0], b2.shape[0]), device="cuda")
    fn = R.ref_box_iou_rotated_cuda if version == 0 else R.ref_box_iou_rotated_v1_cuda
    assert fn(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def nms_rotated_keep(dets, order, thr):
    """dets (n,L) cuda, order (n,) int32 cuda -> keep (n,) bool numpy, kernel ms"""
    R = oracle.ref_cuda()
    n, L = dets.shape
    ds = dets[order.long()].contiguous()
    oh = np.ascontiguousarray(order.cpu().numpy().astype(np.int32))
    keep = np.zeros(n, np.bool_)
    ms = ctypes.c_float(0)
    torch.cuda.synchronize()
    rc = R.ref_nms_rotated_cuda(_p(ds), n, L, oh.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_float(np.float32(thr)),
                                keep.ctypes.data_as(ctypes.POINTER(ctypes.c_bool)), ctypes.byref(ms))
    assert rc == 0, rc
    return keep, ms.value


def roi_align_rotated(x, rois, out_hw, scale, sr, version):
    R = oracle.ref_cuda()
    B, C, H, W = x.shape
    ph, pw = out_hw
    out = torch.zeros((rois.shape[0], C, ph, pw), device="cuda")
    assert R.ref_roi_align_rotated_cuda(version, _p(x), _p(rois), rois.shape[0], C, H, W, ph, pw,
                                        ctypes.c_float(np.float32(scale)), ctypes.c_float(float(sr)), _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def roi_align_rotated_backward(grad, rois, in_shape, scale, sr, version):
    R = oracle.ref_cuda()
    B, C, H, W = in_shape
    ph, pw = grad.shape[2:]
    out = torch.empty(in_shape, device="cuda")
    assert R.ref_roi_align_rotated_backward_cuda(version, _p(grad), _p(rois), rois.shape[0], B, C, H, W, ph, pw,
                                                 ctypes.c_float(np.float32(scale)), ctypes.c_float(float(sr)), _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def feature_refine_backward(grad, boxes, scale, points):
    R = oracle.ref_cuda()
    N, C, H, W = grad.shape
    out = torch.empty_like(grad)
    assert R.ref_feature_refine_backward_cuda(_p(grad), _p(boxes), N, C, H, W, points, ctypes.c_float(np.float32(scale)),
                                              _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def feature_refine(x, boxes, scale, points):
    R = oracle.ref_cuda()
    N, C, H, W = x.shape
    out = torch.zeros_like(x)
    assert R.ref_feature_refine_cuda(_p(x), _p(boxes), N, C, H, W, points, ctypes.c_float(np.float32(scale)), _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def deform_conv(x, offset, weight, stride=1, pad=0, dil=1, dg=1):
    """reference im2col kernel + fp32 matmul (what deform_conv_forward_cuda does, dcn_v1.py:412-454), groups=1."""
    R = oracle.ref_cuda()
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    col = torch.zeros((C * kh * kw, B * Ho * Wo), device="cuda")
    assert R.ref_deformable_im2col_cuda(_p(x), _p(offset), C, H, W, kh, kw, pad, pad, stride, stride, dil, dil, B, dg,
                                        _p(col), _st()) == 0
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    out = weight.reshape(Co, -1) @ col
    torch.backends.cuda.matmul.allow_tf32 = prev
    torch.cuda.synchronize()
    return out.reshape(Co, B, Ho, Wo).permute(1, 0, 2, 3).contiguous()


def deform_col2im(colg, offset, B, C, H, W, k, stride, pad, dil, dg):
    R = oracle.ref_cuda()
    out = torch.empty((B, C, H, W), device="cuda")
    assert R.ref_deformable_col2im_cuda(_p(colg), _p(offset), C, H, W, k, k, pad, pad, stride, stride, dil, dil, B, dg, _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out


def deform_col2im_coord(colg, x, offset, k, stride, pad, dil, dg):
    R = oracle.ref_cuda()
    B, C, H, W = x.shape
    out = torch.empty_like(offset)
    assert R.ref_deformable_col2im_coord_cuda(_p(colg), _p(x), _p(offset), C, H, W, k, k, pad, pad, stride, stride, dil, dil, B, dg,
                                              _p(out), _st()) == 0
    torch.cuda.synchronize()
    return out

"""jdet.ops.roi_align_rotated_v1 mirror (reference: python/jdet/ops/roi_align_rotated_v1.py:300-365).

v1 convention (used by OrientedSingleRoIExtractor / Oriented R-CNN): centre shifted by -0.5,
x = xx*cos + yy*sin, clamp on `< 0`, count = max(grid, 1).  Forward and backward (grad w.r.t. input).
"""
import torch
from torch import nn

from ._common import check, f32c, lib, require_cuda, scratch, stream_ptr


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _roi_align_impl(version, input, rois, output_size, spatial_scale, sampling_ratio):
    assert rois.shape[1] == 6                                  # roi_align_rotated_v1.py:306
    require_cuda(input, rois)
    assert input.dim() == 4
    ph, pw = _pair(output_size)
    B, C, H, W = input.shape
    r = f32c(rois)
    R = r.shape[0]
    out = torch.empty((R, C, ph, pw), dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    sr = int(sampling_ratio)        # the reference passes a float constant to an int parameter (truncation)
    L = lib()
    # a map that is already channel-last in memory (torch.channels_last) is gathered in place: no re-layout pass
    if (input.dtype == torch.float32 and C > 1 and not input.is_contiguous()
            and input.is_contiguous(memory_format=torch.channels_last)):
        with torch.cuda.device(input.device):
            ws = scratch(L.jdet_roi_align_rotated_nhwc_workspace_bytes(R, ph, pw, sr), input.device)
            rc = L.jdet_roi_align_rotated_nhwc(version, input.data_ptr(), B, C, H, W, r.data_ptr(), R, ph, pw,
                                               float(spatial_scale), sr, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                               stream_ptr(input.device))
        if rc == 0:
            return out
        if rc != -3:                # JDET_ERR_UNSUPPORTED: shape outside the channel-last kernel -> NCHW entry point
            check(rc, "roi_align_rotated_nhwc")
    x = f32c(input)
    with torch.cuda.device(x.device):
        ws = scratch(L.jdet_roi_align_rotated_workspace_bytes(B, C, H, W, R, ph, pw, sr), x.device)
        check(L.jdet_roi_align_rotated(version, x.data_ptr(), B, C, H, W, r.data_ptr(), R, ph, pw,
                                       float(spatial_scale), sr, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       stream_ptr(x.device)), "roi_align_rotated")
    return out


def _roi_align_backward_impl(version, grad_output, rois, input_shape, output_size, spatial_scale, sampling_ratio):
    g, r = f32c(grad_output), f32c(rois)
    ph, pw = _pair(output_size)
    B, C, H, W = input_shape
    R = r.shape[0]
    grad_in = torch.empty((B, C, H, W), dtype=torch.float32, device=g.device)
    if grad_in.numel() == 0:
        return grad_in
    sr = int(sampling_ratio)
    L = lib()
    with torch.cuda.device(g.device):
        ws = scratch(L.jdet_roi_align_rotated_backward_workspace_bytes(B, C, H, W, R, ph, pw, sr), g.device)
        check(L.jdet_roi_align_rotated_backward(version, g.data_ptr(), r.data_ptr(), R, B, C, H, W, ph, pw,
                                                float(spatial_scale), sr, grad_in.data_ptr(), ws.data_ptr(), ws.numel(),
                                                stream_ptr(g.device)), "roi_align_rotated_backward")
    return grad_in


class _RotatedROIAlignFn(torch.autograd.Function):
    """jt.Function mirror (execute/grad -> forward/backward); rois get no gradient, as in the reference."""

    @staticmethod
    def forward(ctx, version, input, rois, output_size, spatial_scale, sampling_ratio):
        ctx.version, ctx.output_size, ctx.spatial_scale, ctx.sampling_ratio = version, output_size, spatial_scale, sampling_ratio
        ctx.input_shape = tuple(input.shape)
        ctx.save_for_backward(rois)
        return _roi_align_impl(version, input, rois, output_size, spatial_scale, sampling_ratio)

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        gi = _roi_align_backward_impl(ctx.version, grad_output, rois, ctx.input_shape, ctx.output_size,
                                      ctx.spatial_scale, ctx.sampling_ratio)
        return None, gi, None, None, None, None


def _roi_align_dispatch(version, input, rois, output_size, spatial_scale, sampling_ratio):
    if torch.is_grad_enabled() and isinstance(input, torch.Tensor) and input.requires_grad:
        return _RotatedROIAlignFn.apply(version, input, rois, output_size, spatial_scale, sampling_ratio)
    return _roi_align_impl(version, input, rois, output_size, spatial_scale, sampling_ratio)


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio):
    """_RotatedROIAlign_v1.apply: input (B,C,H,W), rois (R,6)=[batch,cx,cy,w,h,theta] -> (R,C,PH,PW)."""
    return _roi_align_dispatch(1, input, rois, output_size, spatial_scale, sampling_ratio)


class ROIAlignRotated_v1(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio=0):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    execute = forward   # Jittor's name for forward

    def __repr__(self):
        return (self.__class__.__name__ + "(output_size=" + str(self.output_size) + ", spatial_scale=" +
                str(self.spatial_scale) + ", sampling_ratio=" + str(self.sampling_ratio) + ")")

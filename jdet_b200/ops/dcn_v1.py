"""jdet.ops.dcn_v1 mirror — DeformConv v1 forward (reference: python/jdet/ops/dcn_v1.py:559-712).

deform_conv(x, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step)
Forward: no columns tensor is materialised (the bilinear sampler feeds the GEMM directly).  Backward (groups == 1):
hand-written im2col / col2im / col2im_coord kernels around two plain fp32 library GEMMs, in bounded batch chunks.
"""
import math

import torch
from torch import nn

from ._common import check, f32c, lib, require_cuda, stream_ptr


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _output_size(input, weight, padding, dilation, stride):
    channels = weight.size(0)
    output_size = (input.size(0), channels)
    for d in range(input.dim() - 2):
        kernel = dilation[d] * (weight.size(d + 2) - 1) + 1
        output_size += ((input.size(d + 2) + 2 * padding[d] - kernel) // stride[d] + 1,)
    if not all(map(lambda s: s > 0, output_size)):
        raise ValueError("convolution input is too small (output would be {})".format(
            'x'.join(map(str, output_size))))
    return output_size


def _deform_conv_fwd(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step, relu):
    if input is not None and input.dim() != 4:                # dcn_v1.py:571-574
        raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
    require_cuda(input, offset, weight)                       # :588-589
    out_shape = _output_size(input, weight, padding, dilation, stride)
    cur_im2col_step = min(im2col_step, input.shape[0])
    assert (input.shape[0] % cur_im2col_step) == 0, 'im2col step must divide batchsize'   # :591-593
    assert offset.size(0) == input.size(0), "invalid batch size of offset"                  # :422
    x, off, w = f32c(input), f32c(offset), f32c(weight)
    B, C, H, W = x.shape
    Co, _, kh, kw = w.shape
    assert off.shape[1] == deformable_groups * 2 * kh * kw and tuple(off.shape[2:]) == tuple(out_shape[2:])
    out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().jdet_deform_conv_forward(x.data_ptr(), off.data_ptr(), w.data_ptr(), B, C, H, W, Co, kh, kw,
                                             stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
                                             groups, deformable_groups, int(relu), out.data_ptr(),
                                             stream_ptr(x.device)), "deform_conv")
    return out


def _deform_conv_bwd(input, offset, weight, grad_output, stride, padding, dilation, deformable_groups,
                     need_input=True, need_offset=True, need_weight=True, max_col_bytes=512 << 20):
    """deform_conv_backward_input_cuda + deform_conv_backward_parameters_cuda (dcn_v1.py:457-556), groups = 1.
    Hand-written im2col / col2im / col2im_coord kernels around two plain fp32 library GEMMs, in batch chunks so
    the (C*kh*kw, nb*Ho*Wo) columns buffer stays bounded."""
    x, off, w, go = f32c(input), f32c(offset), f32c(weight), f32c(grad_output)
    B, C, H, W = x.shape
    Co, _, kh, kw = w.shape
    Ho, Wo = go.shape[2:]
    K, P = C * kh * kw, Ho * Wo
    L = lib()
    gx = torch.empty_like(x) if need_input else None
    goff = torch.empty_like(off) if need_offset else None
    gw = torch.zeros((Co, K), dtype=torch.float32, device=x.device) if need_weight else None
    nb = max(1, min(B, max_col_bytes // max(1, K * P * 4)))
    geo = (kh, kw, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1], deformable_groups)
    w2d = w.reshape(Co, K)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False            # the reference GEMMs are fp32 SGEMM
    try:
        with torch.cuda.device(x.device):
            st = stream_ptr(x.device)
            for b0 in range(0, B, nb):
                b1 = min(B, b0 + nb)
                n = b1 - b0
                go2d = go[b0:b1].permute(1, 0, 2, 3).reshape(Co, n * P)
                xs, offs = x[b0:b1], off[b0:b1]
                if need_input or need_offset:
                    colg = torch.matmul(w2d.t(), go2d).contiguous()                  # (K, n*P)
                    if need_input:
                        check(L.jdet_deform_col2im(colg.data_ptr(), offs.data_ptr(), n, C, H, W, *geo, gx[b0:b1].data_ptr(), st),
                              "deform_col2im")
                    if need_offset:
                        check(L.jdet_deform_col2im_coord(colg.data_ptr(), xs.data_ptr(), offs.data_ptr(), n, C, H, W, *geo,
                                                         goff[b0:b1].data_ptr(), st), "deform_col2im_coord")
                if need_weight:
                    col = torch.empty((K, n * P), dtype=torch.float32, device=x.device)
                    check(L.jdet_deform_im2col(xs.data_ptr(), offs.data_ptr(), n, C, H, W, *geo, col.data_ptr(), st), "deform_im2col")
                    gw += torch.matmul(go2d, col.t())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return gx, goff, (gw.reshape(w.shape) if need_weight else None)


class DeformConvFunction(torch.autograd.Function):
    """jt.Function mirror of dcn_v1.py:559-650 (execute/grad -> forward/backward).  Backward: groups == 1."""

    @staticmethod
    def forward(ctx, input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step):
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        ctx.save_for_backward(input, offset, weight)
        return _deform_conv_fwd(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step, False)

    @staticmethod
    def backward(ctx, grad_output):
        input, offset, weight = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.cfg
        if groups != 1:
            raise NotImplementedError("DeformConv backward: groups > 1 is not built yet")
        gx, goff, gw = _deform_conv_bwd(input, offset, weight, grad_output, stride, padding, dilation, dg,
                                        ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, goff, gw, None, None, None, None, None, None


def deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64, _relu=False):
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    needs_grad = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (input, offset, weight))
    if needs_grad:
        out = DeformConvFunction.apply(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step)
        return torch.relu(out) if _relu else out
    return _deform_conv_fwd(input, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step, _relu)


class DeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
        assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels // groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        nn.init.uniform_(self.weight, -stdv, stdv)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)

    execute = forward

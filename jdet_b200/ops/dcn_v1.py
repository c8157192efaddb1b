"""jdet.ops.dcn_v1 mirror — DeformConv v1 forward (reference: python/jdet/ops/dcn_v1.py:559-712).

deform_conv(x, offset, weight, stride, padding, dilation, groups, deformable_groups, im2col_step)
No columns tensor is materialised: the bilinear sampler feeds the GEMM directly.  Forward only.
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


def deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64, _relu=False):
    if input is not None and input.dim() != 4:                # dcn_v1.py:571-574
        raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
    require_cuda(input, offset, weight)                       # :588-589
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
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
                                             groups, deformable_groups, int(_relu), out.data_ptr(),
                                             stream_ptr(x.device)), "deform_conv")
    return out


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
        return deform_conv(x, offset, self.weight.detach(), self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)

    execute = forward

"""AlignConv mirror (reference: python/jdet/models/roi_heads/s2anet_head.py:657-723).

Only AlignConv is mirrored from s2anet_head.py — it is the hot-path module of S2ANet's head; the
rest of the head is out of scope this round.
"""
import torch
from torch import nn

from ...ops._common import check, f32c, lib, require_cuda, scratch, stream_ptr
from ...ops.dcn_v1 import DeformConv, _deform_conv_bwd, deform_conv
from ..boxes.box_ops import delta2bbox_rotated


def bbox_decode(bbox_preds, anchors, means=[0, 0, 0, 0, 0], stds=[1, 1, 1, 1, 1]):
    """FAM deltas (N,5,H,W) + anchors (H*W,5) -> refined anchors (N,H,W,5)  (s2anet_head.py:631-654).
    One batched call instead of the reference's per-image Python loop; same arithmetic."""
    num_imgs, _, H, W = bbox_preds.shape
    deltas = bbox_preds.permute(0, 2, 3, 1).reshape(-1, 5)
    rois = anchors.unsqueeze(0).expand(num_imgs, -1, -1).reshape(-1, 5)
    return delta2bbox_rotated(rois, deltas, means, stds, wh_ratio_clip=1e-6).reshape(num_imgs, H, W, 5)


class _AlignConvFn(torch.autograd.Function):
    """Training path of the fused AlignConv: forward is the same tcgen05 kernel; backward goes through ReLU'
    (out > 0) and the DeformConv backward with the offsets recomputed from the anchors (they carry no gradient)."""

    @staticmethod
    def forward(ctx, module, x, anchors, weight, stride):
        out = module._fused(x, anchors, weight.detach(), stride)
        ctx.module, ctx.stride = module, stride
        ctx.save_for_backward(x, anchors, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        x, anchors, weight, out = ctx.saved_tensors
        offset = ctx.module.get_offset_batched(anchors, ctx.stride)
        g = grad_output * (out > 0).to(grad_output.dtype)
        gx, _, gw = _deform_conv_bwd(x, offset, weight, g, (1, 1), (1, 1), (1, 1), 1,
                                     ctx.needs_input_grad[1], False, ctx.needs_input_grad[3])
        return None, gx, None, gw, None


class AlignConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, deformable_groups=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.deform_conv = DeformConv(in_channels, out_channels, kernel_size=kernel_size,
                                      padding=(kernel_size - 1) // 2, deformable_groups=deformable_groups)
        self.relu = nn.ReLU()

    def init_weights(self):
        nn.init.normal_(self.deform_conv.weight, mean=0.0, std=0.01)

    @torch.no_grad()
    def get_offset(self, anchors, featmap_size, stride):
        """anchors (H*W,5) of ONE image -> (2*k*k, H, W) offset field (s2anet_head.py:677-713)."""
        feat_h, feat_w = featmap_size
        return self.get_offset_batched(anchors.reshape(1, feat_h, feat_w, 5), stride)[0]

    @torch.no_grad()
    def get_offset_batched(self, anchors, stride):
        require_cuda(anchors)
        a = f32c(anchors)
        N, H, W, _ = a.shape
        k = self.kernel_size
        off = torch.empty((N, 2 * k * k, H, W), dtype=torch.float32, device=a.device)
        if off.numel():
            with torch.cuda.device(a.device):
                check(lib().jdet_align_conv_offset(a.data_ptr(), N, H, W, float(stride), k, off.data_ptr(),
                                                   stream_ptr(a.device)), "align_conv_offset")
        return off

    def _fused(self, x, anchors, weight, stride):
        xx, aa, w = f32c(x), f32c(anchors), f32c(weight)
        N, H, W = aa.shape[:3]
        C, Co = xx.shape[1], w.shape[0]
        out = torch.empty((N, Co, H, W), dtype=torch.float32, device=xx.device)
        if out.numel() == 0:
            return out
        L = lib()
        with torch.cuda.device(xx.device):
            ws = scratch(L.jdet_align_conv_forward_workspace_bytes(N, C, H, W, Co), xx.device)
            check(L.jdet_align_conv_forward(xx.data_ptr(), aa.data_ptr(), w.data_ptr(), N, C, H, W, Co,
                                            float(stride), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                            stream_ptr(xx.device)), "align_conv")
        return out

    def forward(self, x, anchors, stride):
        """x (N,C,H,W), anchors (N,H,W,5) image space -> relu(deform_conv(x, offset(anchors)))."""
        require_cuda(x, anchors)
        dc = self.deform_conv
        fused = self.kernel_size == 3 and dc.deformable_groups == 1 and dc.groups == 1
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or dc.weight.requires_grad)
        if fused and not needs_grad:
            return self._fused(x, anchors, dc.weight.detach(), stride)
        if fused:
            return _AlignConvFn.apply(self, x, anchors, dc.weight, stride)
        offset = self.get_offset_batched(anchors, stride)      # no gradient, as in the reference (@jt.no_grad, :676)
        return deform_conv(x, offset, dc.weight, dc.stride, dc.padding, dc.dilation, dc.groups,
                           dc.deformable_groups, _relu=True)

    execute = forward

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

    @torch.no_grad()
    def forward_multi(self, xs, anchors_list, strides):
        """The same AlignConv on every FPN level in ONE library call (inference): one persistent tcgen05 launch over all
        levels' tiles, one weight split (jdet_align_conv_forward_multi).  xs[l] (N,C,H_l,W_l), anchors_list[l] (N,H_l,W_l,5).
        Falls back to the per-level calls outside the fused shape class."""
        import ctypes
        dc = self.deform_conv
        n = len(xs)
        fused = self.kernel_size == 3 and dc.deformable_groups == 1 and dc.groups == 1
        C, Co = xs[0].shape[1], dc.weight.shape[0]
        if not fused or n > 8 or C % 16 != 0 or Co % 32 != 0 or Co > 256 or xs[0].shape[0] == 0:
            return [self.forward(x, a, s) for x, a, s in zip(xs, anchors_list, strides)]
        require_cuda(*xs, *anchors_list)
        # torch.channels_last maps (what cuDNN convolutions produce under that memory format) are sampled in place
        cl = all(x.dtype == torch.float32 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last) for x in xs)
        xx = list(xs) if cl else [f32c(x) for x in xs]
        aa = [f32c(a) for a in anchors_list]
        w = f32c(dc.weight.detach())
        N = xx[0].shape[0]
        outs = [torch.empty((N, Co, x.shape[2], x.shape[3]), dtype=torch.float32, device=x.device) for x in xx]
        arr = lambda vals, ct: (ct * n)(*vals)
        Hs, Ws = arr([x.shape[2] for x in xx], ctypes.c_int), arr([x.shape[3] for x in xx], ctypes.c_int)
        L = lib()
        with torch.cuda.device(xx[0].device):
            ws = scratch(L.jdet_align_conv_forward_multi_workspace_bytes(n, N, C, Hs, Ws, Co), xx[0].device)
            check(L.jdet_align_conv_forward_multi(arr([x.data_ptr() for x in xx], ctypes.c_void_p), arr([a.data_ptr() for a in aa], ctypes.c_void_p),
                                                  w.data_ptr(), n, N, C, Hs, Ws, Co, arr([float(s_) for s_ in strides], ctypes.c_float),
                                                  arr([o.data_ptr() for o in outs], ctypes.c_void_p), int(cl), ws.data_ptr(), ws.numel(),
                                                  stream_ptr(xx[0].device)), "align_conv_multi")
        return outs

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


class _ConvReLU(nn.Module):
    """ConvModule(conv 3x3 + ReLU) as S2ANet builds it (models/utils/modules.py:93-190, no norm): the same attribute
    names (`conv`, `activate`), so a reference checkpoint's keys (`fam_reg_convs.0.conv.weight`, ...) load as they are."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=1, padding=1)
        self.activate = nn.ReLU()

    def forward(self, x):
        return self.activate(self.conv(x))


class S2ANetHead(nn.Module):
    """Forward-only (inference) restatement of S2ANetHead (s2anet_head.py:20-252, 510-601): the real caller of
    bbox_decode -> AlignConv -> ORConv2d/RotationInvariantPooling -> multiclass_nms_rotated.  Batched where the
    reference loops over images in Python; losses / target assignment are out of scope."""

    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=2, with_orconv=True, anchor_scales=[4],
                 anchor_ratios=[1.0], anchor_strides=[8, 16, 32, 64, 128], anchor_base_sizes=None,
                 target_means=(.0, .0, .0, .0, .0), target_stds=(1.0, 1.0, 1.0, 1.0, 1.0),
                 test_cfg=dict(nms_pre=2000, min_bbox_size=0, score_thr=0.05, nms=dict(type='nms_rotated', iou_thr=0.1),
                               max_per_img=2000)):
        super().__init__()
        from ...ops.orn import ORConv2d, RotationInvariantPooling
        from ..boxes.anchor_generator import AnchorGeneratorRotatedS2ANet
        self.num_classes, self.in_channels, self.feat_channels = num_classes, in_channels, feat_channels
        self.stacked_convs, self.with_orconv = stacked_convs, with_orconv
        self.anchor_strides = list(anchor_strides)
        self.anchor_base_sizes = list(anchor_strides) if anchor_base_sizes is None else anchor_base_sizes
        self.target_means, self.target_stds = target_means, target_stds
        self.cls_out_channels = num_classes - 1            # sigmoid classification (FocalLoss config)
        self.test_cfg = test_cfg
        self.anchor_generators = [AnchorGeneratorRotatedS2ANet(b, anchor_scales, anchor_ratios) for b in self.anchor_base_sizes]
        self.base_anchors = dict()
        self.fam_reg_convs = nn.ModuleList(_ConvReLU(in_channels if i == 0 else feat_channels, feat_channels)
                                           for i in range(stacked_convs))
        self.fam_reg = nn.Conv2d(feat_channels, 5, 1)
        self.align_conv = AlignConv(feat_channels, feat_channels, kernel_size=3)
        if with_orconv:
            self.or_conv = ORConv2d(feat_channels, int(feat_channels / 8), kernel_size=3, padding=1, arf_config=(1, 8))
        else:
            self.or_conv = nn.Conv2d(feat_channels, feat_channels, 3, padding=1)
        self.or_pool = RotationInvariantPooling(256, 8)
        self.odm_reg_convs = nn.ModuleList(_ConvReLU(feat_channels, feat_channels) for _ in range(stacked_convs))
        self.odm_cls_convs = nn.ModuleList(
            _ConvReLU(int(feat_channels / 8) if i == 0 and with_orconv else feat_channels, feat_channels)
            for i in range(stacked_convs))
        self.odm_cls = nn.Conv2d(feat_channels, self.cls_out_channels, 3, padding=1)
        self.odm_reg = nn.Conv2d(feat_channels, 5, 3, padding=1)
        self.init_weights()

    def init_weights(self):
        import math
        bias_cls = float(-math.log((1 - 0.01) / 0.01))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.odm_cls.bias, bias_cls)
        self.align_conv.init_weights()

    def _fam(self, x, stride):
        f = x
        for conv in self.fam_reg_convs:
            f = conv(f)
        fam_bbox_pred = self.fam_reg(f)
        lvl = self.anchor_strides.index(stride)
        size = tuple(fam_bbox_pred.shape[-2:])
        key = (lvl, size, x.device)
        if key not in self.base_anchors:
            self.base_anchors[key] = self.anchor_generators[lvl].grid_anchors(size, stride, device=x.device)
        return fam_bbox_pred, bbox_decode(fam_bbox_pred, self.base_anchors[key], self.target_means, self.target_stds)

    def _odm(self, align_feat):
        or_feat = self.or_conv(align_feat)
        reg_feat, cls_feat = or_feat, (self.or_pool(or_feat) if self.with_orconv else or_feat)
        for conv in self.odm_reg_convs:
            reg_feat = conv(reg_feat)
        for conv in self.odm_cls_convs:
            cls_feat = conv(cls_feat)
        return self.odm_cls(cls_feat), self.odm_reg(reg_feat)

    @torch.no_grad()
    def forward_single(self, x, stride):
        """one level, as the reference's forward_single (s2anet_head.py:207-252)"""
        fam_bbox_pred, refine_anchor = self._fam(x, stride)
        cls, reg = self._odm(self.align_conv(x, refine_anchor, stride))
        return fam_bbox_pred, refine_anchor, cls, reg

    @torch.no_grad()
    def forward_levels(self, feats):
        """all levels: FAM per level, then ONE AlignConv call over every level (AlignConv.forward_multi), then ODM per level;
        level by level the same values as forward_single"""
        fam = [self._fam(x, s) for x, s in zip(feats, self.anchor_strides)]
        aligned = self.align_conv.forward_multi(list(feats), [f[1] for f in fam], self.anchor_strides[:len(feats)])
        return [(f[0], f[1]) + self._odm(a) for f, a in zip(fam, aligned)]

    @torch.no_grad()
    def get_bboxes_single(self, cls_score_list, bbox_pred_list, mlvl_anchors, cfg=None, scale_factor=1.0, rescale=True):
        """s2anet_head.py:543-601; `rescale` divides the boxes' x, y, w, h by scale_factor before NMS (:585-586; get_bboxes
        defaults to rescale=True, :517), so detections come back in original-image coordinates."""
        from ...ops.nms_rotated import multiclass_nms_rotated
        from ..boxes.box_ops import delta2bbox_rotated, rotated_box_to_poly
        cfg = self.test_cfg if cfg is None else cfg
        mlvl_bboxes, mlvl_scores = [], []
        for cls_score, bbox_pred, anchors in zip(cls_score_list, bbox_pred_list, mlvl_anchors):
            scores = cls_score.permute(1, 2, 0).reshape(-1, self.cls_out_channels).sigmoid()
            bbox_pred = bbox_pred.permute(1, 2, 0).reshape(-1, 5)
            anchors = anchors.reshape(-1, 5)
            nms_pre = cfg.get('nms_pre', -1)
            if nms_pre > 0 and scores.shape[0] > nms_pre:
                _, topk_inds = scores.max(dim=1)[0].topk(nms_pre)
                anchors, bbox_pred, scores = anchors[topk_inds], bbox_pred[topk_inds], scores[topk_inds]
            mlvl_bboxes.append(delta2bbox_rotated(anchors, bbox_pred, self.target_means, self.target_stds))
            mlvl_scores.append(scores)
        mlvl_bboxes, mlvl_scores = torch.cat(mlvl_bboxes), torch.cat(mlvl_scores)
        if rescale:
            mlvl_bboxes = mlvl_bboxes.clone()
            mlvl_bboxes[..., :4] /= scale_factor
        mlvl_scores = torch.cat([mlvl_scores.new_zeros((mlvl_scores.shape[0], 1)), mlvl_scores], dim=1)
        det_bboxes, det_labels = multiclass_nms_rotated(mlvl_bboxes, mlvl_scores, cfg['score_thr'], cfg['nms'], cfg['max_per_img'])
        return rotated_box_to_poly(det_bboxes[:, :5]), det_bboxes[:, 5], det_labels

    @torch.no_grad()
    def forward(self, feats, img_metas=None, rescale=True):
        """feats: list of (N,C,H_l,W_l) FPN maps -> list over images of (polys (k,8), scores (k,), labels (k,)).
        img_metas (optional): per image a dict with 'scale_factor' (default 1.0), as the reference's get_bboxes reads it."""
        outs = self.forward_levels(feats)
        num_imgs = feats[0].shape[0]
        results = []
        for i in range(num_imgs):
            sf = 1.0 if img_metas is None else img_metas[i].get('scale_factor', 1.0)
            results.append(self.get_bboxes_single([o[2][i] for o in outs], [o[3][i] for o in outs], [o[1][i] for o in outs],
                                                  scale_factor=sf, rescale=rescale))
        return results

    execute = forward

"""Forward-only OrientedRPNHead (Oriented R-CNN): reference python/jdet/models/roi_heads/oriented_rpn_head.py
:104-226 (layers, forward_single, _get_bboxes_single, get_bboxes).  Losses / target assignment are out of
scope.  The class-agnostic proposal NMS is HORIZONTAL in the reference (`jt.nms` on the proposals' bounding
rectangles, levels kept apart by a coordinate offset, :198-206); here that third-party step is
`torchvision.ops.nms` (parity of that library call is not pinned by any reference test)."""
import torch
from torch import nn

from ..boxes.anchor_generator import AnchorGenerator
from ..boxes.coder import midpoint_offset_decode, obb2hbb


class OrientedRPNHead(nn.Module):
    def __init__(self, in_channels, num_classes=1, min_bbox_size=0, nms_thresh=0.8, nms_pre=2000, nms_post=2000,
                 feat_channels=256, use_sigmoid_cls=True,
                 anchor_generator=dict(scales=[8], ratios=[0.5, 1.0, 2.0], strides=[4, 8, 16, 32, 64]),
                 target_means=(0.,) * 6, target_stds=(1., 1., 1., 1., .5, .5)):
        super().__init__()
        self.min_bbox_size, self.nms_thresh, self.nms_pre, self.nms_post = min_bbox_size, nms_thresh, nms_pre, nms_post
        self.use_sigmoid_cls = use_sigmoid_cls
        self.cls_out_channels = num_classes if use_sigmoid_cls else num_classes + 1
        self.means, self.stds = tuple(target_means), tuple(target_stds)
        cfg = {k: v for k, v in anchor_generator.items() if k != "type"}
        self.anchor_generator = AnchorGenerator(**cfg)
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, self.num_anchors * self.cls_out_channels, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, self.num_anchors * 6, 1)
        for m in (self.rpn_conv, self.rpn_cls, self.rpn_reg):
            nn.init.normal_(m.weight, 0, 0.01)
            nn.init.constant_(m.bias, 0)

    def forward_single(self, x):
        x = torch.relu(self.rpn_conv(x))
        return self.rpn_cls(x), self.rpn_reg(x)

    @torch.no_grad()
    def proposals_single(self, cls_scores, bbox_preds, mlvl_anchors):
        """one image: per level top-nms_pre by score -> decode -> size filter -> per-level horizontal NMS -> top nms_post.
        Returns (k,6) [x,y,w,h,theta,score]."""
        from torchvision.ops import nms
        scores, deltas, anchors, level = [], [], [], []
        for lvl, (cs, bp, an) in enumerate(zip(cls_scores, bbox_preds, mlvl_anchors)):
            cs = cs.permute(1, 2, 0)
            sc = cs.reshape(-1).sigmoid() if self.use_sigmoid_cls else cs.reshape(-1, 2).softmax(1)[:, 1]
            bp = bp.permute(1, 2, 0).reshape(-1, 6)
            if 0 < self.nms_pre < sc.shape[0]:
                sc, idx = sc.sort(descending=True)
                sc, idx = sc[:self.nms_pre], idx[:self.nms_pre]
                bp, an = bp[idx], an[idx]
            scores.append(sc); deltas.append(bp); anchors.append(an)
            level.append(torch.full_like(sc, lvl))
        scores, level = torch.cat(scores), torch.cat(level)
        props = midpoint_offset_decode(torch.cat(anchors), torch.cat(deltas), self.means, self.stds)
        if self.min_bbox_size >= 0:
            ok = (props[:, 2] > self.min_bbox_size) & (props[:, 3] > self.min_bbox_size)
            props, scores, level = props[ok], scores[ok], level[ok]
        if props.shape[0] == 0:
            return props.new_zeros((0, 6))
        hbb = obb2hbb(props)
        hbb = hbb + (level * (hbb.max() - hbb.min() + 1))[:, None]       # levels never overlap
        keep = nms(hbb, scores, self.nms_thresh)[:self.nms_post]
        return torch.cat([props[keep], scores[keep, None]], 1)

    @torch.no_grad()
    def forward(self, feats):
        """feats: list of (N,C,H_l,W_l) -> list over images of (k,6) proposals."""
        outs = [self.forward_single(x) for x in feats]
        anchors = self.anchor_generator.grid_anchors([o[0].shape[-2:] for o in outs], device=feats[0].device)
        return [self.proposals_single([o[0][i] for o in outs], [o[1][i] for o in outs], anchors)
                for i in range(feats[0].shape[0])]

    execute = forward

"""Forward-only OrientedRPNHead (Oriented R-CNN): reference python/jdet/models/roi_heads/oriented_rpn_head.py
:104-226 (layers, forward_single, _get_bboxes_single, get_bboxes).  Losses / target assignment are out of
scope.  The class-agnostic proposal NMS is HORIZONTAL in the reference (`jt.nms` on the proposals' bounding
rectangles, levels kept apart by a coordinate offset, :198-206); here that third-party step runs on
the library's own NMS kernels (horizontal_nms below; parity of jt.nms is not pinned by any reference test)."""
import torch
from torch import nn

from ..boxes.anchor_generator import AnchorGenerator
from ..boxes.coder import midpoint_offset_decode, obb2hbb


def horizontal_nms(hbb, scores, group, thr):
    """Class-agnostic proposal NMS on bounding rectangles with groups kept apart (reference: jt.nms on coordinates offset
    per level, oriented_rpn_head.py:198-206) -> kept indices by descending score.  It is the library's own rotated NMS on
    axis-aligned boxes (theta = 0, group id in the label column): suppress on IoU > thr, by descending score.  (torchvision.ops.nms
    spends 18 ms in its single-thread gather_keep_from_mask on the 20k boxes of two tiles; this takes ~0.3 ms.)  The IoU of two
    axis-aligned boxes may differ from jt.nms's closed form in the last bit; no reference test pins that third-party call."""
    from ...ops.nms_rotated import argsort_desc, nms_rotated_cuda
    d6 = torch.cat([(hbb[:, :2] + hbb[:, 2:]) * 0.5, hbb[:, 2:] - hbb[:, :2], torch.zeros_like(scores)[:, None],
                    group.to(torch.float32)[:, None]], 1).contiguous()
    order = argsort_desc(scores.contiguous())
    kept = nms_rotated_cuda(d6, order, thr, box_length=6)
    order = order.long()
    return order[kept[order]]


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
        keep = horizontal_nms(obb2hbb(props), scores, level, self.nms_thresh)[:self.nms_post]
        return torch.cat([props[keep], scores[keep, None]], 1)

    @torch.no_grad()
    def proposals_batched(self, outs, mlvl_anchors):
        """All images and levels at once, no per-image Python loop: per (image, level) top-nms_pre -> ONE decode -> ONE
        horizontal NMS call over (image, level) groups -> per image the nms_post best.
        Returns (props (N, nms_post, 6) [x,y,w,h,theta,score] zero padded, counts (N,) int64).  Same proposals per image as
        proposals_single."""
        N = outs[0][0].shape[0]
        scores, deltas, anchors, level = [], [], [], []
        for lvl, ((cs, bp), an) in enumerate(zip(outs, mlvl_anchors)):
            cs = cs.permute(0, 2, 3, 1)
            sc = cs.reshape(N, -1).sigmoid() if self.use_sigmoid_cls else cs.reshape(N, -1, 2).softmax(2)[..., 1]
            bp = bp.permute(0, 2, 3, 1).reshape(N, -1, 6)
            an = an[None].expand(N, -1, 4)
            if 0 < self.nms_pre < sc.shape[1]:
                sc, idx = sc.topk(self.nms_pre, dim=1)                    # sorted descending, like sort()[:nms_pre]
                bp = bp.gather(1, idx[..., None].expand(-1, -1, 6))
                an = an.gather(1, idx[..., None].expand(-1, -1, 4))
            scores.append(sc); deltas.append(bp); anchors.append(an)
            level.append(torch.full((sc.shape[1],), lvl, device=sc.device, dtype=torch.float32))
        scores, deltas, anchors = torch.cat(scores, 1), torch.cat(deltas, 1), torch.cat(anchors, 1)
        K = scores.shape[1]
        group = torch.cat(level)[None] + len(outs) * torch.arange(N, device=scores.device, dtype=torch.float32)[:, None]   # (N, K)
        props = midpoint_offset_decode(anchors.reshape(-1, 4), deltas.reshape(-1, 6), self.means, self.stds)
        scores, group = scores.reshape(-1), group.reshape(-1)
        hbb = obb2hbb(props)
        if self.min_bbox_size >= 0:                                       # too-small boxes: zero area (IoU 0 with everything), score -1, dropped below
            ok = (props[:, 2] > self.min_bbox_size) & (props[:, 3] > self.min_bbox_size)
            scores = torch.where(ok, scores, scores.new_full((), -1.0))
            hbb = torch.where(ok[:, None], hbb, torch.zeros_like(hbb))
        keep = horizontal_nms(hbb, scores, group, self.nms_thresh)                # kept indices by descending score
        keep = keep[scores[keep] >= 0]
        img = keep // K
        cnt_all = torch.bincount(img, minlength=N)
        img_s, perm = torch.sort(img, stable=True)                                # image-major, score order inside an image
        rank = torch.arange(img_s.shape[0], device=img_s.device) - (cnt_all.cumsum(0) - cnt_all)[img_s]
        counts = cnt_all.clamp(max=self.nms_post)
        out = props.new_zeros((N * self.nms_post + 1, 6))                          # last row: dump slot of the overflow
        dst = torch.where(rank < self.nms_post, img_s * self.nms_post + rank, torch.full_like(rank, N * self.nms_post))
        src = keep[perm]
        out.index_copy_(0, dst, torch.cat([props[src], scores[src, None]], 1))
        return out[:-1].reshape(N, self.nms_post, 6), counts

    @torch.no_grad()
    def forward_batched(self, feats):
        outs = [self.forward_single(x) for x in feats]
        anchors = self.anchor_generator.grid_anchors([o[0].shape[-2:] for o in outs], device=feats[0].device)
        return self.proposals_batched(outs, anchors)

    @torch.no_grad()
    def forward(self, feats):
        """feats: list of (N,C,H_l,W_l) -> list over images of (k,6) proposals."""
        outs = [self.forward_single(x) for x in feats]
        anchors = self.anchor_generator.grid_anchors([o[0].shape[-2:] for o in outs], device=feats[0].device)
        return [self.proposals_single([o[0][i] for o in outs], [o[1][i] for o in outs], anchors)
                for i in range(feats[0].shape[0])]

    execute = forward

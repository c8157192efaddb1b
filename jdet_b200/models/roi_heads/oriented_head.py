"""Forward-only OrientedHead (Oriented R-CNN RoI head): reference python/jdet/models/roi_heads/oriented_head.py
:117-130 (arb2roi), :226-272 (forward_single), :412-443 (get_bboxes), :132-161 (get_results), :519-536 (test path).
The RoI features come from OrientedSingleRoIExtractor -> jdet_roi_align_rotated (version 1): this head is the
real caller of the rotated RoIAlign kernels at 2000 proposals per image.  The reference head applies the score
threshold and converts to polygons but runs no NMS (that happens at tile merge); same here."""
import torch
from torch import nn

from ..boxes.coder import obb2poly, oriented_delta_xywht_decode
from ..roi_extractors import OrientedSingleRoIExtractor


class OrientedHead(nn.Module):
    def __init__(self, num_classes=15, in_channels=256, fc_out_channels=1024, num_shared_fcs=2, score_thresh=0.05,
                 roi_feat_size=7, featmap_strides=(4, 8, 16, 32), sample_num=2, extend_factor=(1.4, 1.2),
                 target_means=(0.,) * 5, target_stds=(.1, .1, .2, .2, .1), reg_class_agnostic=True):
        super().__init__()
        self.num_classes, self.score_thresh, self.reg_class_agnostic = num_classes, score_thresh, reg_class_agnostic
        self.means, self.stds = tuple(target_means), tuple(target_stds)
        self.bbox_roi_extractor = OrientedSingleRoIExtractor(
            roi_layer=dict(type="ROIAlignRotated_v1", output_size=roi_feat_size, sampling_ratio=sample_num),
            out_channels=in_channels, featmap_strides=list(featmap_strides), extend_factor=extend_factor)
        dims = [in_channels * roi_feat_size * roi_feat_size] + [fc_out_channels] * num_shared_fcs
        self.shared_fcs = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.fc_cls = nn.Linear(dims[-1], num_classes + 1)
        self.fc_reg = nn.Linear(dims[-1], 5 if reg_class_agnostic else 5 * num_classes)
        for m in self.shared_fcs:
            nn.init.xavier_uniform_(m.weight); nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.fc_cls.weight, 0, 0.01); nn.init.constant_(self.fc_cls.bias, 0)
        nn.init.normal_(self.fc_reg.weight, 0, 0.001); nn.init.constant_(self.fc_reg.bias, 0)

    @staticmethod
    def arb2roi(bbox_list):
        """list over images of (k_i, >=5) oriented boxes -> (sum k_i, 6) [image index, x, y, w, h, theta]."""
        rois = [torch.cat([b.new_full((b.shape[0], 1), float(i)), b[:, :5]], 1) for i, b in enumerate(bbox_list)]
        return torch.cat(rois, 0) if rois else torch.zeros((0, 6))

    def forward_single(self, feats, rois):
        x = self.bbox_roi_extractor(feats[:self.bbox_roi_extractor.num_inputs], rois).flatten(1)
        for fc in self.shared_fcs:
            x = torch.relu(fc(x))
        return self.fc_cls(x), self.fc_reg(x)

    def get_bboxes(self, rois, cls_score, bbox_pred, scale_factor=1.0):
        """-> (polys+score (k,9), labels (k,)): every (box, class) pair above score_thresh, row-major (box, class)."""
        scores = cls_score.softmax(1)
        boxes = oriented_delta_xywht_decode(rois[:, 1:], bbox_pred, self.means, self.stds)
        boxes = boxes.reshape(boxes.shape[0], -1, 5).clone()
        boxes[..., :4] = boxes[..., :4] / scale_factor
        if boxes.shape[1] == 1:
            boxes = boxes.expand(-1, self.num_classes, 5)
        fg = scores[:, :-1]
        valid = fg > self.score_thresh
        if not bool(valid.any()):
            return boxes.new_zeros((0, 9)), torch.zeros((0,), dtype=torch.int64, device=boxes.device)
        return torch.cat([obb2poly(boxes[valid]), fg[valid][:, None]], 1), valid.nonzero()[:, 1]

    @torch.no_grad()
    def detect_records(self, feats, props, counts, iou_thr=0.1, max_per_img=2000, out=None):
        """The inference tail in one pass over the batch: padded proposals (N, P, >=5) + counts (N,) -> fused 4-level rotated
        RoIAlign -> shared FCs -> decode -> score threshold -> ONE per-class rotated NMS call for all images (labels offset
        per image) -> per image a detection record [x,y,w,h,theta,score,label] x max_per_img + count row, written into
        `out` (N, max_per_img + 1, 7) — e.g. the persistent all-gather send buffer.  One host sync (the nonzero of the
        score threshold)."""
        N, P = props.shape[:2]
        dev = props.device
        img = torch.arange(N, device=dev, dtype=props.dtype)[:, None, None].expand(N, P, 1)
        rois = torch.cat([img, props[..., :5]], 2).reshape(N * P, 6)
        cls_score, bbox_pred = self.forward_single(feats, rois)
        return self.records_from_scores(rois, cls_score, bbox_pred, counts, N, P, iou_thr, max_per_img, out)

    @torch.no_grad()
    def records_from_scores(self, rois, cls_score, bbox_pred, counts, N, P, iou_thr=0.1, max_per_img=2000, out=None):
        """decode -> score threshold -> one rotated NMS call -> per-image records (the tail of detect_records)"""
        from ...ops._common import check, lib, stream_ptr
        from ...ops.nms_rotated import argsort_desc, nms_rotated_cuda
        dev = rois.device
        if out is None:
            out = torch.empty((N, max_per_img + 1, 7), dtype=torch.float32, device=dev)
        scores = cls_score.softmax(1)[:, :-1]
        boxes = oriented_delta_xywht_decode(rois[:, 1:], bbox_pred, self.means, self.stds).reshape(N * P, -1, 5)
        live = (torch.arange(P, device=dev)[None] < counts[:, None]).reshape(N * P, 1)
        valid = (scores > self.score_thresh) & live
        ri, ci = valid.nonzero(as_tuple=True)                                   # row-major (box, class), like the reference's mask
        if ri.numel() == 0:
            out.zero_()
            return out
        b = boxes[ri, ci if boxes.shape[1] > 1 else torch.zeros_like(ci)]
        sc = scores[ri, ci].contiguous()
        lab = (ri // P) * self.num_classes + ci                                  # image-major label: classes of different images never meet
        d6 = torch.cat([b, lab.to(torch.float32)[:, None]], 1).contiguous()
        order = argsort_desc(sc)
        keep = nms_rotated_cuda(d6, order, iou_thr, box_length=6)
        L = lib()
        with torch.cuda.device(dev):
            for i in range(N):
                check(L.jdet_pack_detections_range(d6.data_ptr(), d6.shape[0], sc.data_ptr(), order.data_ptr(), keep.data_ptr(),
                                                   float(i * self.num_classes), float((i + 1) * self.num_classes), max_per_img,
                                                   out[i].data_ptr(), stream_ptr(dev)), "pack_detections_range")
        return out

    @torch.no_grad()
    def forward(self, feats, proposal_list, scale_factors=None):
        """feats: FPN maps (N,C,H_l,W_l); proposal_list: per image (k,>=5).  One batched pass over all images' RoIs
        (the reference loops over images, :519-535); returns per image (polys (k,8), scores (k,), labels (k,))."""
        rois = self.arb2roi(proposal_list)
        cls_score, bbox_pred = self.forward_single(feats, rois)
        out, lo = [], 0
        for i, p in enumerate(proposal_list):
            hi = lo + p.shape[0]
            sf = 1.0 if scale_factors is None else scale_factors[i]
            det, lab = self.get_bboxes(rois[lo:hi], cls_score[lo:hi], bbox_pred[lo:hi], sf)
            out.append((det[:, :8], det[:, 8], lab))
            lo = hi
        return out

    execute = forward

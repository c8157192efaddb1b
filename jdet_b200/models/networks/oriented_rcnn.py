"""Oriented R-CNN, inference only, tiles in -> detection records out (BASELINE configs[4]).

Reference: python/jdet/models/networks/rcnn.py:43-50 (backbone -> neck -> rpn -> roi head) with the Oriented R-CNN config
(Resnet50 + FPN(num_outs=5) + OrientedRPNHead + OrientedHead).  The backbone and FPN are OUT of the hot-path scope
(cuDNN convolutions in the reference as here): a torchvision ResNet-50 + FeaturePyramidNetwork with random weights stands
in for them so that the heads — the callers of the rotated RoIAlign / NMS kernels — run on
real feature-map shapes and the end-to-end number has the caller-true boundary: uint8 tiles up, 2001 x 7 records down.
"""
import torch
from torch import nn

from ..roi_heads import OrientedHead, OrientedRPNHead


class _R50FPN(nn.Module):
    def __init__(self, out_channels=256):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.fpn = torchvision.ops.FeaturePyramidNetwork([256, 512, 1024, 2048], out_channels)

    def forward(self, x):
        x = self.stem(x)
        feats = {}
        for i, l in enumerate(self.layers):
            x = l(x)
            feats[str(i)] = x
        p = list(self.fpn(feats).values())
        p.append(torch.nn.functional.max_pool2d(p[-1], 1, stride=2))          # P6: FPN(num_outs=5) subsamples P5
        return p


class OrientedRCNN(nn.Module):
    """images (N,3,H,W) uint8 / float in [0,255] -> (N, max_per_img + 1, 7) records (jdet_b200.dist.unpack_detections)."""

    def __init__(self, num_classes=15, max_per_img=2000, nms_iou_thr=0.1, mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375)):
        super().__init__()
        self.backbone = _R50FPN(256)
        self.rpn = OrientedRPNHead(256)
        self.roi_head = OrientedHead(num_classes=num_classes)
        self.max_per_img, self.nms_iou_thr = max_per_img, nms_iou_thr
        self.register_buffer("mean", torch.tensor(mean).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor(std).view(1, 3, 1, 1))

    @torch.no_grad()
    def forward(self, images, out=None):
        x = (images.to(torch.float32) - self.mean) / self.std
        # NCHW throughout: with TF32 off, cuDNN's fp32 convolutions of this backbone take 29.5 ms for two 1024^2 tiles in NCHW
        # and 56 ms in channels_last on a B200 (tools/backbone_time.py; cudnn.benchmark makes no difference)
        feats = [f.contiguous() for f in self.backbone(x.contiguous())]       # the RoI / RPN ops take NCHW maps
        props, counts = self.rpn.forward_batched(feats)
        return self.roi_head.detect_records(feats, props, counts, self.nms_iou_thr, self.max_per_img, out=out)

    execute = forward

#!/usr/bin/env python
"""Where the cfg5 network's time goes (2 x 1024^2 tiles, strict fp32): preprocessing, backbone + FPN, RPN, RoI head."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jdet_b200.models.networks import OrientedRCNN  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(11)
net = OrientedRCNN().to(dev).eval().requires_grad_(False)
torch.nn.init.normal_(net.rpn.rpn_cls.weight, 0, 0.05); torch.nn.init.normal_(net.rpn.rpn_reg.weight, 0, 0.02)
torch.nn.init.normal_(net.roi_head.fc_cls.weight, 0, 0.05)
x = torch.randint(0, 256, (2, 3, 1024, 1024), dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc = [0.0] * 4
with torch.no_grad():
    for it in range(8):
        ev[0].record()
        xin = ((x.to(torch.float32) - net.mean) / net.std).contiguous()
        ev[1].record()
        feats = [f.contiguous() for f in net.backbone(xin)]
        ev[2].record()
        props, counts = net.rpn.forward_batched(feats)
        ev[3].record()
        out = net.roi_head.detect_records(feats, props, counts, net.nms_iou_thr, net.max_per_img)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i in range(4):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) / 5
print("preprocess %.2f ms | backbone+FPN %.2f ms | RPN (convs, decode, top-k, proposal NMS) %.2f ms | RoI head (extractor, FCs, decode, NMS, records) %.2f ms" % tuple(acc))
print("proposals per image:", [int(c) for c in counts], " detections:", [int(r[-1, 0].item()) for r in out])

from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    feats = [f.contiguous() for f in net.backbone(xin)]
    props, counts = net.rpn.forward_batched(feats)
    out = net.roi_head.detect_records(feats, props, counts, net.nms_iou_thr, net.max_per_img)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:22]
for e in rows:
    print("%9.1f us  x%-4d %s" % (e.device_time_total, e.count, e.key[:110]))

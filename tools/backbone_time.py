#!/usr/bin/env python
"""cfg5 stand-in backbone (torchvision R50-FPN, random weights, strict fp32), 2 x 1024^2 tiles: cuDNN settings A/B."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jdet_b200.models.networks import OrientedRCNN  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
x = torch.randint(0, 256, (2, 3, 1024, 1024), dtype=torch.uint8, device=dev)


def t(fn, k=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k


for bench in (False, True):
    for cl in (False, True):
        torch.backends.cudnn.benchmark = bench
        torch.manual_seed(0)
        net = OrientedRCNN().to(dev).eval().requires_grad_(False)
        if cl:
            net.backbone = net.backbone.to(memory_format=torch.channels_last)
        xin = ((x.float() - net.mean) / net.std)
        xin = xin.contiguous(memory_format=torch.channels_last) if cl else xin.contiguous()
        with torch.no_grad():
            ms_b = t(lambda: net.backbone(xin))
            ms_n = t(lambda: net(x))
        print("cudnn.benchmark=%s weights channels_last=%s: backbone %.1f ms, whole network %.1f ms" % (bench, cl, ms_b, ms_n), flush=True)
        del net

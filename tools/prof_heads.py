#!/usr/bin/env python
"""Kernel-time breakdown of the cfg5 heads (OrientedRPNHead -> OrientedHead -> rotated NMS -> records) on one GPU."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jdet_b200.models.roi_heads import OrientedHead, OrientedRPNHead  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
rpn = OrientedRPNHead(256).to(dev).eval().requires_grad_(False)
torch.nn.init.normal_(rpn.rpn_cls.weight, 0, 0.05); torch.nn.init.normal_(rpn.rpn_reg.weight, 0, 0.02)
head = OrientedHead(num_classes=15).to(dev).eval().requires_grad_(False)
torch.nn.init.normal_(head.fc_cls.weight, 0, 0.05)
fpn5 = [torch.randn((2, 256, 1024 // s, 1024 // s), device=dev) for s in (4, 8, 16, 32, 64)]


def step():
    props, counts = rpn.forward_batched(fpn5)
    return head.detect_records(fpn5, props, counts, 0.1, 2000)


for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    step()
b.record(); torch.cuda.synchronize()
print("ms per step (2 tiles):", a.elapsed_time(b) / 5)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

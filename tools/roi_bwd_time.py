#!/usr/bin/env python
"""Time ROIAlignRotated_v1 backward at cfg2 and hash the gradient of a deterministic small case."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jdet_b200.ops.roi_align_rotated_v1 import _roi_align_backward_impl  # noqa: E402
from _inputs import dota_boxes  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
rois = torch.as_tensor(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048, 1024.0)], 1)).to(dev)
go = torch.randn((2048, 256, 7, 7), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: _roi_align_backward_impl(1, go, rois, (1, 256, 256, 256), (7, 7), 0.25, 2)
for _ in range(3):
    fn()
tot = 0.0
for _ in range(20):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    tot += a.elapsed_time(b)
# one RoI per pixel neighbourhood would collide; a single RoI has a deterministic (atomic-order-free) gradient per bin-pixel pair only
# when no two of its bins touch the same pixel channel — use the sum as a coarse check instead
g = fn()
print("backward %.1f us   sum %.6e  abs-sum %.6e" % (tot / 20 * 1e3, g.double().sum().item(), g.double().abs().sum().item()))

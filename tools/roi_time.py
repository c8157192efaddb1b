#!/usr/bin/env python
"""Time roi_align_rotated_v1 at cfg2 (NCHW input: relayout + gather; channels_last input: gather only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jdet_b200.ops as ops  # noqa: E402
from _inputs import dota_boxes  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
feat = torch.randn((1, 256, 256, 256), device=dev)
fcl = feat.contiguous(memory_format=torch.channels_last)
rois = torch.as_tensor(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048, 1024.0)], 1)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, k=30):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(k):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / k * 1e3


for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print("nchw %.1f us" % t(lambda: ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2)),
          "channels_last %.1f us" % t(lambda: ops.roi_align_rotated_v1.roi_align(fcl, rois, (7, 7), 0.25, 2)), flush=True)

#!/usr/bin/env python
"""Run every hot-path op at its BASELINE config a few times — the command ncu wraps.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_ops.py
  ncu --set full --clock-control none --import-source on -k regex:'jdet' -o gpurun_out/prof python tools/prof_ops.py --reps 1
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jdet_b200.ops as ops  # noqa: E402
from jdet_b200.models.roi_heads.s2anet_head import AlignConv  # noqa: E402
from _inputs import clustered_boxes, dota_boxes, s2anet_anchors, tie_free_scores  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--only", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cu = lambda x, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(x), dtype=dt).to(dev)
want = lambda k: (not a.only) or k in a.only.split(",")

if want("roi"):
    feat = torch.randn((1, 256, 256, 256), device=dev)
    rois = cu(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048)], 1))
    for _ in range(a.reps):
        ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2)
if want("roicl"):
    feat = torch.randn((1, 256, 256, 256), device=dev).contiguous(memory_format=torch.channels_last)
    rois = cu(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048)], 1))
    for _ in range(a.reps):
        ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2)
if want("nms"):
    n = 100000
    d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
    td, ts, tl = cu(d), cu(tie_free_scores(rng, n)), cu(rng.integers(0, 15, n), torch.int64)
    for _ in range(a.reps):
        ops.nms_rotated.ml_nms_rotated(td, ts, tl, 0.1)
if want("iou"):
    b1, b2 = cu(dota_boxes(rng, 16384)), cu(dota_boxes(rng, 16384))
    for _ in range(a.reps):
        ops.box_iou_rotated(b1, b2)
if want("fr") or want("ac"):
    x = torch.randn((8, 256, 128, 128), device=dev)
    an = cu(s2anet_anchors(rng, 8, 128, 128, 8))
    bx = an[..., [1, 0, 2, 3, 4]].contiguous()
    if want("fr"):
        for _ in range(a.reps):
            ops.fr.feature_refine(x, bx, 1 / 8., 1)
            ops.fr.feature_refine(x, bx, 1 / 8., 5)
    if want("ac"):
        m = AlignConv(256, 256, 3).to(dev).requires_grad_(False)
        for _ in range(a.reps):
            m(x, an, 8)
torch.cuda.synchronize()
print("done")

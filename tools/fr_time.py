#!/usr/bin/env python
"""Time feature_refine at the cfg4 levels (bs 8, 256 ch): points 1 and 5."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jdet_b200.ops as ops  # noqa: E402
from _inputs import s2anet_anchors  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
levels = [(128, 8), (64, 16), (32, 32), (16, 64), (8, 128)]
xs = [torch.randn((8, 256, h, h), device=dev) for h, _ in levels]
bs = [torch.as_tensor(s2anet_anchors(rng, 8, h, h, s)[..., [1, 0, 2, 3, 4]].copy()).to(dev) for h, s in levels]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(points):
    return [ops.fr.feature_refine(x, b, 1.0 / s, points) for x, b, (_, s) in zip(xs, bs, levels)]


def t(fn, k=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(k):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / k


def multi(points):
    return ops.fr.feature_refine_multi(xs, bs, [1.0 / s for _, s in levels], points)


print("per-level calls: points=1 %.4f ms   points=5 %.4f ms" % (t(lambda: run(1)), t(lambda: run(5))), flush=True)
print("one call       : points=1 %.4f ms   points=5 %.4f ms" % (t(lambda: multi(1)), t(lambda: multi(5))), flush=True)
print("level 0 alone  : points=5 %.4f ms" % t(lambda: ops.fr.feature_refine(xs[0], bs[0], 1.0 / 8, 5)), flush=True)
print("JDET_FR_P5_GATHER=%s" % os.environ.get("JDET_FR_P5_GATHER"), flush=True)

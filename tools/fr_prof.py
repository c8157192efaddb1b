#!/usr/bin/env python
"""feature_refine points=5 at cfg4, the 5 levels in one call, three times: what ncu wraps (tools/fr_ab.sh)."""
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
for _ in range(3):
    ops.fr.feature_refine_multi(xs, bs, [1.0 / s for _, s in levels], 5)
torch.cuda.synchronize()

#!/usr/bin/env python
"""cfg3 NMS (100k x 15, thr 0.1), two calls: what ncu wraps for a source-level look at the mask / exact / scan kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jdet_b200.ops import nms_rotated as N  # noqa: E402
from _inputs import clustered_boxes, dota_boxes, tie_free_scores  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
n = 100000
d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
td, ts, tl = torch.as_tensor(d).to(dev), torch.as_tensor(s).to(dev), torch.as_tensor(l).to(dev)
out = torch.empty((2001, 7), device=dev)
for _ in range(2):
    N.ml_nms_rotated_record(td, ts, tl, 0.1, 2000, out)
torch.cuda.synchronize()
print(int(out[-1, 0].item()))

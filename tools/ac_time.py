#!/usr/bin/env python
"""AlignConv cfg4 (bs 8, 256 -> 256, 5 levels in one launch): time and a hash of the outputs (JDET_ALIGN_CONV_2CTA=1: CTA pairs)."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from jdet_b200.models.roi_heads.s2anet_head import AlignConv  # noqa: E402
from _inputs import s2anet_anchors  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
levels = [(128, 8), (64, 16), (32, 32), (16, 64), (8, 128)]
g = torch.Generator(device=dev).manual_seed(5)
xs = [torch.randn((8, 256, h, h), device=dev, generator=g) for h, _ in levels]
an = [torch.as_tensor(s2anet_anchors(rng, 8, h, h, s)).to(dev) for h, s in levels]
torch.manual_seed(3)
ac = AlignConv(256, 256, 3).to(dev).requires_grad_(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: ac.forward_multi(xs, an, [s for _, s in levels])
outs = fn()
torch.cuda.synchronize()
h = hashlib.sha1()
for o in outs:
    h.update(o.cpu().numpy().tobytes())
tot = 0.0
for it in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    if it >= 3:
        tot += a.elapsed_time(b) / 5
print("align_conv cfg4: %.4f ms  hash %s  2CTA=%s  mean|out| %.5f" % (tot, h.hexdigest()[:12], os.environ.get("JDET_ALIGN_CONV_2CTA"), float(sum(o.abs().mean() for o in outs))), flush=True)

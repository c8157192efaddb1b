#!/usr/bin/env python
"""cfg3 NMS step (100k x 15, thr 0.1): where the time goes between the library calls, eager vs one captured CUDA graph."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jdet_b200.ops as ops  # noqa: E402
from jdet_b200.ops import nms_rotated as N  # noqa: E402
from _inputs import clustered_boxes, dota_boxes, tie_free_scores  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
n = 100000
d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
td, ts, tl = torch.as_tensor(d).to(dev), torch.as_tensor(s).to(dev), torch.as_tensor(l).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty((2001, 7), device=dev)


def t(fn, k=20, fl=True):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(k):
        if fl:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / k


d6 = torch.cat([td, tl.to(torch.float32).unsqueeze(1)], dim=1).contiguous()
order = N.argsort_desc(ts)
keep = N.nms_rotated_cuda(d6, order, 0.1, box_length=6)
print("argsort_desc        %.4f ms" % t(lambda: N.argsort_desc(ts)))
print("nms_rotated_cuda    %.4f ms" % t(lambda: N.nms_rotated_cuda(d6, order, 0.1, box_length=6)))
print("record (3 calls)    %.4f ms" % t(lambda: N.ml_nms_rotated_record(td, ts, tl, 0.1, 2000, out)))
print("record, no L2 flush %.4f ms" % t(lambda: N.ml_nms_rotated_record(td, ts, tl, 0.1, 2000, out), fl=False))
ref = out.clone()
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    for _ in range(3):
        N.ml_nms_rotated_record(td, ts, tl, 0.1, 2000, out)
torch.cuda.current_stream().wait_stream(st)
with torch.cuda.graph(g):
    N.ml_nms_rotated_record(td, ts, tl, 0.1, 2000, out)
out.zero_()
g.replay()
torch.cuda.synchronize()
assert torch.equal(out, ref)
print("record, graph replay %.4f ms" % t(g.replay))
print("record, graph replay, no flush %.4f ms" % t(g.replay, fl=False))

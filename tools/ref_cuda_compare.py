#!/usr/bin/env python
"""Ours vs the reference's own CUDA kernels (oracle/_ref/libref_cuda.so, nvcc sm_100a, reference launch
geometry) on the same B200, same inputs, CUDA-event timing with an L2 flush before every run.
Test infrastructure / evidence only — writes gpurun_out/ref_cuda_compare.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jdet_b200.ops as ops  # noqa: E402
from jdet_b200.models.roi_heads.s2anet_head import AlignConv  # noqa: E402
import _refcuda  # noqa: E402
from _inputs import clustered_boxes, dota_boxes, s2anet_anchors, tie_free_scores  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cu = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush_buf.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


res = {}
# roi_align cfg2
feat = torch.randn((1, 256, 256, 256), device=dev)
rois = cu(np.concatenate([np.zeros((2048, 1), np.float32), dota_boxes(rng, 2048)], 1))
ours = timeit(lambda: ops.roi_align_rotated_v1.roi_align(feat, rois, (7, 7), 0.25, 2))
ref = timeit(lambda: _refcuda.roi_align_rotated(feat, rois, (7, 7), 0.25, 2, 1))
res["roi_align_rotated_v1 cfg2"] = {"ours_ms": ours, "reference_cuda_kernel_ms": ref, "speedup": ref / ours}
# box_iou_rotated 16k (reference kernel: full clip+hull for every pair)
b1, b2 = cu(dota_boxes(rng, 16384)), cu(dota_boxes(rng, 16384))
ours = timeit(lambda: ops.box_iou_rotated(b1, b2), 5)
ref = timeit(lambda: _refcuda.box_iou_rotated(b1, b2, 0), 3, 1)
res["box_iou_rotated 16k x 16k"] = {"ours_ms": ours, "reference_cuda_kernel_ms": ref, "speedup": ref / ours}
b1, b2 = cu(dota_boxes(rng, 1000)), cu(dota_boxes(rng, 1000))
ours = timeit(lambda: ops.box_iou_rotated(b1, b2))
ref = timeit(lambda: _refcuda.box_iou_rotated(b1, b2, 0))
res["box_iou_rotated 1k x 1k (cfg1)"] = {"ours_ms": ours, "reference_cuda_kernel_ms": ref, "speedup": ref / ours}
# nms cfg3: reference = its mask kernel time alone (its host reduce and the 1.25 GB mask copy are NOT counted)
n = 100000
d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
td, ts, tl = cu(d), cu(s), cu(l, torch.int64)
ours = timeit(lambda: ops.nms_rotated.ml_nms_rotated(td, ts, tl, 0.1), 5)
d6 = torch.cat([td, tl.float()[:, None]], 1).contiguous()
order = ops.nms_rotated.argsort_desc(ts)
_, ref_mask_ms = _refcuda.nms_rotated_keep(d6, order, 0.1)
import time
t0 = time.perf_counter()
_refcuda.nms_rotated_keep(d6, order, 0.1)
ref_total = (time.perf_counter() - t0) * 1e3
res["ml_nms_rotated 100k x 15 (cfg3)"] = {"ours_ms_whole_op": ours, "reference_cuda_mask_kernel_only_ms": ref_mask_ms,
                                          "reference_cuda_path_wall_ms_incl_host_reduce": ref_total,
                                          "speedup_vs_mask_kernel_only": ref_mask_ms / ours}
# feature_refine + AlignConv, level 0 of cfg4
x = torch.randn((8, 256, 128, 128), device=dev)
an = cu(s2anet_anchors(rng, 8, 128, 128, 8))
bx = an[..., [1, 0, 2, 3, 4]].contiguous()
for pts in (1, 5):
    ours = timeit(lambda: ops.fr.feature_refine(x, bx, 1 / 8., pts))
    ref = timeit(lambda: _refcuda.feature_refine(x, bx, 1 / 8., pts))
    res["feature_refine points=%d, 8x256x128x128" % pts] = {"ours_ms": ours, "reference_cuda_kernel_ms": ref, "speedup": ref / ours}
m = AlignConv(256, 256, 3).to(dev).requires_grad_(False)
w = m.deform_conv.weight.detach()
ours = timeit(lambda: m(x, an, 8), 5)
off = m.get_offset_batched(an, 8)
ref = timeit(lambda: _refcuda.deform_conv(x, off, w, 1, 1, 1, 1), 3, 1)     # reference im2col kernel + cuBLAS fp32 GEMM (no offsets/ReLU cost)
res["AlignConv 256->256, 8x128x128 (level 0 of cfg4)"] = {"ours_ms": ours, "reference_im2col_plus_sgemm_ms": ref, "speedup": ref / ours}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_compare.json"), "w"), indent=1)
for k, v in res.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})

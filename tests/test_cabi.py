"""CPU-only: the C-ABI library loads and exports exactly what include/jdet_b200.h declares;
the host-side mirrors keep the reference's error behaviour without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jdet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jdet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from jdet_b200 import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), "missing export " + s
        assert s in _lib.SIGNATURES, "no ctypes signature for " + s
    assert b"sm_100a" in L.jdet_version()


def test_workspace_queries_need_no_gpu():
    from jdet_b200 import _lib
    L = _lib.lib()
    assert L.jdet_box_iou_rotated_workspace_bytes(1000, 1000) >= 2 * 1000 * 32
    assert L.jdet_nms_rotated_workspace_bytes(100000, 6) > 100000 * 32
    assert L.jdet_roi_align_rotated_workspace_bytes(1, 256, 256, 256, 2048, 7, 7, 2) >= 256 * 256 * 256 * 4
    assert L.jdet_roi_align_rotated_workspace_bytes(2, 1024, 64, 64, 2, 7, 7, 2) == 256   # direct path
    # argument errors are reported before any CUDA call
    assert L.jdet_box_iou_rotated(None, -1, None, 0, None, 0, None, 0, None) == -1
    assert L.jdet_feature_refine(None, None, 1, 1, 1, 1, 3, 1.0, None, None) == -1
    assert L.jdet_nms_rotated(None, 5, 7, None, 0.1, None, None, 0, None) == -1


def test_no_cpu_fallback():
    import jdet_b200.ops as ops
    b = torch.zeros((3, 5))
    with pytest.raises(NotImplementedError):
        ops.box_iou_rotated(b, b)
    with pytest.raises(NotImplementedError):
        ops.nms_rotated.nms_rotated(b, torch.zeros(3), 0.1)
    with pytest.raises(NotImplementedError):
        ops.roi_align_rotated_v1.roi_align(torch.zeros(1, 4, 8, 8), torch.zeros(2, 6), (7, 7), 1.0, 2)
    with pytest.raises(NotImplementedError):
        ops.fr.feature_refine(torch.zeros(1, 4, 8, 8), torch.zeros(1, 8, 8, 5), 1.0, 1)
    with pytest.raises(NotImplementedError):
        ops.dcn_v1.deform_conv(torch.zeros(1, 4, 8, 8), torch.zeros(1, 18, 8, 8), torch.zeros(4, 4, 3, 3), 1, 1)
    with pytest.raises(NotImplementedError):
        ops.nms_rotated.nms_rotated_cpu(b, torch.zeros(3, dtype=torch.int32), 0.1)


def test_reference_asserts_are_kept():
    import jdet_b200.ops as ops
    with pytest.raises(AssertionError):    # box_iou_rotated.py:503
        ops.box_iou_rotated(torch.zeros((3, 5)), torch.zeros((3, 5), dtype=torch.float64))
    with pytest.raises(AssertionError):    # roi_align_rotated.py:263
        ops.roi_align_rotated.roi_align(torch.zeros(1, 4, 8, 8), torch.zeros(2, 5), (7, 7), 1.0, 2)
    with pytest.raises(AssertionError):    # fr.py:261
        ops.fr.feature_refine(torch.zeros(1, 4, 8, 8), torch.zeros(1, 8, 8, 5), 1.0, 3)
    with pytest.raises(ValueError):        # dcn_v1.py:571-574
        ops.dcn_v1.deform_conv(torch.zeros(4, 8, 8), torch.zeros(1, 18, 8, 8), torch.zeros(4, 4, 3, 3))
    with pytest.raises(AssertionError):    # nms_rotated.py:516
        ops.nms_rotated.ml_nms_rotated(torch.zeros((0, 5)), torch.zeros(0), torch.zeros(0), 0.1)
    assert ops.nms_rotated.nms_rotated(torch.zeros((0, 5)), torch.zeros(0), 0.1).numel() == 0   # :528-529


def test_install_as_jdet():
    import jdet_b200
    jdet_b200.install_as_jdet()
    from jdet.ops import box_iou_rotated, box_iou_rotated_v1, roi_align_rotated_v1   # noqa: F401
    from jdet.models.roi_heads.s2anet_head import AlignConv                          # noqa: F401
    assert hasattr(roi_align_rotated_v1, "ROIAlignRotated_v1")


def test_box_coders_and_anchors_cpu():
    """delta2bbox_rotated / bbox_decode / S2ANet grid anchors against a numpy restatement of
    models/boxes/box_ops.py:229-285, s2anet_head.py:631-654, anchor_generator.py:127-183 (torch ops: CPU is fine)."""
    import numpy as np
    from jdet_b200.models.boxes import AnchorGeneratorRotatedS2ANet, delta2bbox_rotated, norm_angle
    from jdet_b200.models.roi_heads import bbox_decode
    rng = np.random.default_rng(0)
    gen = AnchorGeneratorRotatedS2ANet(4, [4.0], [1.0], angles=[0.0])
    anchors = gen.grid_anchors((6, 7), 8).numpy()
    assert anchors.shape == (42, 5)
    assert np.allclose(anchors[0], [1.5, 1.5, 16, 16, 0]) and np.allclose(anchors[8], [9.5, 9.5, 16, 16, 0])   # row-major, x fastest
    rois = np.concatenate([rng.uniform(0, 100, (50, 2)), rng.uniform(4, 40, (50, 2)), rng.uniform(-1.5, 1.5, (50, 1))], 1).astype(np.float32)
    deltas = (rng.standard_normal((50, 5)) * 0.5).astype(np.float32)
    got = delta2bbox_rotated(torch.from_numpy(rois), torch.from_numpy(deltas), wh_ratio_clip=1e-6).numpy()
    dx, dy, dw, dh, da = deltas.T
    mr = np.abs(np.log(1e-6))
    dw, dh = np.clip(dw, -mr, mr), np.clip(dh, -mr, mr)
    x, y, w, h, a = rois.T
    gx = dx * w * np.cos(a) - dy * h * np.sin(a) + x
    gy = dx * w * np.sin(a) + dy * h * np.cos(a) + y
    ga = (np.pi * da + a + np.pi / 4) % np.pi - np.pi / 4
    want = np.stack([gx, gy, w * np.exp(dw), h * np.exp(dh), ga], 1)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-4)
    assert float(norm_angle(torch.tensor(-np.pi / 2))) == pytest.approx(np.pi / 2, abs=1e-6)
    preds = torch.from_numpy((rng.standard_normal((2, 5, 6, 7)) * 0.3).astype(np.float32))
    out = bbox_decode(preds, gen.grid_anchors((6, 7), 8))
    assert out.shape == (2, 6, 7, 5)
    one = delta2bbox_rotated(gen.grid_anchors((6, 7), 8), preds[1].permute(1, 2, 0).reshape(-1, 5), wh_ratio_clip=1e-6)
    assert torch.equal(out[1].reshape(-1, 5), one)


def test_orn_arf_and_poly_cpu():
    """active_rotating_filter against a loop restatement of ARF_forward_cpu_kernel (ops/orn.py:136-170);
    rotated_box_to_poly against box_ops.py:568-590 (numpy); torch ops, so CPU is fine."""
    import numpy as np
    from jdet_b200.ops.orn import ORConv2d, RotationInvariantPooling, active_rotating_filter, arf_indices
    from jdet_b200.models.boxes import rotated_box_to_poly
    torch.manual_seed(0)
    for (no, ni, nori, nrot) in ((4, 3, 1, 8), (2, 3, 8, 8), (3, 2, 4, 4)):
        w, ind = torch.randn(no, ni, nori, 3, 3), arf_indices(nori, nrot, (3, 3))
        got = active_rotating_filter(w, ind).numpy().ravel()
        nE = nori * 9
        want = np.zeros(no * nrot * ni * nE, np.float32)
        wf, idf = w.numpy().ravel(), ind.numpy().ravel()
        for i in range(no):
            for j in range(ni):
                for l in range(nE):
                    for k in range(nrot):
                        want[i * (nrot * ni * nE) + k * (ni * nE) + j * nE + int(idf[l * nrot + k]) - 1] = wf[i * ni * nE + j * nE + l]
        assert np.array_equal(got, want)
    m = ORConv2d(16, 4, 3, padding=1, arf_config=(1, 8))
    y = m(torch.randn(2, 16, 8, 8))
    assert y.shape == (2, 32, 8, 8) and RotationInvariantPooling(256, 8)(y).shape == (2, 4, 8, 8)
    # rotation 0 of the ARF is the filter itself: the first of every 8 output channels is a plain conv with it
    plain = torch.nn.functional.conv2d(torch.ones(1, 16, 5, 5), m.weight[:, :, 0], None, 1, 1)
    assert torch.allclose(m(torch.ones(1, 16, 5, 5))[:, 0::8] - m.bias[0::8][None, :, None, None], plain, atol=1e-5)
    r = np.array([[10, 20, 8, 4, 0.3], [0, 0, 2, 2, 0]], np.float32)
    got = rotated_box_to_poly(torch.from_numpy(r)).numpy()
    for b, p in zip(r, got):
        x, y_, w, h, a = b
        rect = np.array([[-w / 2, w / 2, w / 2, -w / 2], [-h / 2, -h / 2, h / 2, h / 2]])
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        q = R.dot(rect)
        assert np.allclose(p, np.stack([q[0] + x, q[1] + y_], 1).ravel(), atol=1e-5)


def test_s2anet_head_wiring_cpu(monkeypatch):
    """Layer wiring of the forward-only S2ANetHead (s2anet_head.py:207-252) with AlignConv stubbed out (no GPU here):
    FAM regression -> refined anchors -> [AlignConv] -> ORConv2d -> (pooled) cls / reg towers."""
    import torch
    import jdet_b200.models.roi_heads.s2anet_head as H
    monkeypatch.setattr(H.AlignConv, "forward", lambda self, x, anchors, stride: x)
    head = H.S2ANetHead(16, 256, test_cfg=dict(nms_pre=200, score_thr=0.3, nms=dict(iou_thr=0.1), max_per_img=100)).eval()
    assert len(list(head.fam_reg_convs[0].children())) == 2          # conv + relu, nothing applied twice
    fam, refine, cls, reg = head.forward_single(torch.randn(2, 256, 16, 16), 8)
    assert fam.shape == (2, 5, 16, 16) and refine.shape == (2, 16, 16, 5)
    assert cls.shape == (2, 15, 16, 16) and reg.shape == (2, 5, 16, 16)


def test_oriented_rcnn_coders_cpu():
    """Decoders of the Oriented R-CNN inference path (coder.py:377-438, 486-519): exact inverses on constructed cases."""
    import math
    import torch
    from jdet_b200.models.boxes import (midpoint_offset_decode, oriented_delta_xywht_decode, obb2hbb, obb2poly, rectpoly2obb,
                                        regular_obb)
    g = torch.Generator().manual_seed(0)
    n = 200
    obb = torch.stack([torch.rand(n, generator=g) * 500, torch.rand(n, generator=g) * 500,
                       40 + torch.rand(n, generator=g) * 100, 5 + torch.rand(n, generator=g) * 30,
                       (torch.rand(n, generator=g) - 0.5) * 3.0], 1)
    obb = regular_obb(obb)
    poly = obb2poly(obb).reshape(n, 4, 2)
    assert torch.allclose(rectpoly2obb(poly.reshape(n, 8)), obb, atol=1e-3)
    # midpoint offsets of a true rectangle w.r.t. its own bounding box decode back to the rectangle
    hbb = obb2hbb(obb)
    assert torch.allclose(hbb[:, :2], poly.min(1)[0], atol=1e-3) and torch.allclose(hbb[:, 2:], poly.max(1)[0], atol=1e-3)
    gx, gy = (hbb[:, 0] + hbb[:, 2]) / 2, (hbb[:, 1] + hbb[:, 3]) / 2
    gw, gh = hbb[:, 2] - hbb[:, 0], hbb[:, 3] - hbb[:, 1]
    top = poly[torch.arange(n), poly[..., 1].argmin(1)]          # vertex on the top edge
    right = poly[torch.arange(n), poly[..., 0].argmax(1)]        # vertex on the right edge
    deltas = torch.zeros(n, 6)
    deltas[:, 4] = (top[:, 0] - gx) / gw / 0.5
    deltas[:, 5] = (right[:, 1] - gy) / gh / 0.5
    got = midpoint_offset_decode(hbb, deltas)
    assert torch.allclose(got[:, :4], obb[:, :4], atol=2e-2)
    dth = torch.remainder(got[:, 4] - obb[:, 4] + math.pi / 2, math.pi) - math.pi / 2
    assert dth.abs().max() < 1e-3
    # OrientedDeltaXYWHT: deltas built with the encoder's formulas decode to the target
    roi = obb
    tgt = regular_obb(obb + torch.tensor([3., -2., 5., 1., 0.1]))
    c, s = torch.cos(-roi[:, 4]), torch.sin(-roi[:, 4])
    dx = (c * (tgt[:, 0] - roi[:, 0]) + s * (tgt[:, 1] - roi[:, 1])) / roi[:, 2]
    dy = (-s * (tgt[:, 0] - roi[:, 0]) + c * (tgt[:, 1] - roi[:, 1])) / roi[:, 3]
    d = torch.stack([dx, dy, torch.log(tgt[:, 2] / roi[:, 2]), torch.log(tgt[:, 3] / roi[:, 3]), tgt[:, 4] - roi[:, 4]], 1)
    stds = torch.tensor([.1, .1, .2, .2, .1])
    back = oriented_delta_xywht_decode(roi, d / stds)
    assert torch.allclose(back[:, :4], tgt[:, :4], atol=1e-2)
    dth = torch.remainder(back[:, 4] - tgt[:, 4] + math.pi / 2, math.pi) - math.pi / 2
    assert dth.abs().max() < 1e-4


def test_horizontal_anchor_generator_cpu():
    """anchor_generator.py:186-420: ratios x scales boxes per cell, anchor index fastest, cells row-major."""
    import torch
    from jdet_b200.models.boxes import AnchorGenerator
    gen = AnchorGenerator(strides=[4, 8], ratios=[0.5, 1.0, 2.0], scales=[8])
    a0, a1 = gen.grid_anchors([(3, 5), (2, 2)])
    assert a0.shape == (3 * 5 * 3, 4) and a1.shape == (2 * 2 * 3, 4) and gen.num_base_anchors == [3, 3]
    w, h = a0[:, 2] - a0[:, 0], a0[:, 3] - a0[:, 1]
    assert torch.allclose(w * h, torch.full_like(w, 32.0 * 32.0), rtol=1e-5)            # area (stride*scale)^2 at every ratio
    assert torch.allclose((h / w)[:3], torch.tensor([0.5, 1.0, 2.0]), rtol=1e-5)
    ctr = (a0[:, :2] + a0[:, 2:]) / 2
    assert torch.allclose(ctr[3:6], torch.tensor([[4.0, 0.0]]).expand(3, 2))            # second cell of the first row
    assert torch.allclose(ctr[15:18], torch.tensor([[0.0, 4.0]]).expand(3, 2))          # first cell of the second row

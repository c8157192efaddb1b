"""CPU-only: the oracle restatement against (a) the committed golden vectors produced by the
reference's own source and (b), when oracle/_ref is built here, the compiled reference directly."""
import ctypes
import os

import numpy as np
import pytest

import oracle
from _inputs import ADVERSARIAL, clustered_boxes, dota_boxes, tie_free_scores

GOLD = os.path.join(os.path.dirname(__file__), "golden")
fp = ctypes.POINTER(ctypes.c_float)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_known_answers():
    k = np.load(os.path.join(GOLD, "known_answers.npz"))
    # K1: ops/box_iou_rotated.py:513-514
    for variant in (oracle.VARIANT_CPU, oracle.VARIANT_CUDA):
        got = oracle.box_iou_rotated(k["k1_boxes"], k["k1_boxes"], 0, variant)
        assert np.array_equal(bits(got), bits(k["k1_iou"]))
        assert abs(got[0, 1] - 0.2) < 1e-7 and got[0, 0] == 1.0
        # K3
        assert np.float32(oracle.single_iou(k["k3_a"], k["k3_b"], 0, variant)) == k["k3_iou"]
        assert abs(float(k["k3_iou"]) - 0.44319284) < 1e-7
        # K2: ops/nms_rotated.py:599-603 keeps [2] for both ml_nms_rotated and nms_rotated
        got2 = oracle.box_iou_rotated(k["k2_dets"], k["k2_dets"], 0, variant)
        assert np.array_equal(bits(got2), bits(k["k2_iou"]))
        assert list(oracle.ml_nms_rotated(k["k2_dets"], k["k2_scores"], k["k2_labels"], 0.3, variant)) == [2]
        assert list(oracle.nms_rotated(k["k2_dets"], k["k2_scores"], 0.3, variant)) == [2]
    # different labels => IoU 0 => everything kept
    assert list(oracle.ml_nms_rotated(k["k2_dets"], k["k2_scores"], [0, 1, 2], 0.3)) == [0, 1, 2]


def test_iou_golden_bit_exact():
    g = np.load(os.path.join(GOLD, "ref_cpu_iou.npz"))
    b1, b2 = g["boxes1"], g["boxes2"]
    for ver in (0, 1):
        L = oracle.lib()
        raw = np.zeros((len(b1), len(b2)), np.float32)
        L.orc_box_iou_rotated(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2), raw.ctypes.data_as(fp),
                              ver, oracle.VARIANT_CPU, 0)
        assert np.array_equal(bits(raw), bits(g["iou_v%d_cpu" % ver])), "cpu-variant restatement drifted"
        L.orc_box_iou_rotated(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2), raw.ctypes.data_as(fp),
                              ver, oracle.VARIANT_CUDA, 0)
        assert np.array_equal(bits(raw), bits(g["iou_v%d_cudavariant" % ver])), "cuda-variant restatement drifted"


def test_nms_golden():
    g = np.load(os.path.join(GOLD, "ref_cpu_nms.npz"))
    d6, order = g["dets6"], g["order"]
    assert np.array_equal(order, oracle.argsort_desc(g["scores"]))
    for thr in (0.1, 0.3, 0.5):
        for bl, d in ((5, np.ascontiguousarray(d6[:, :5])), (6, d6)):
            want = g["keep%d_cpu_thr%02d" % (bl, int(thr * 10))]
            got = oracle.nms_rotated_keep(d, order, thr, oracle.VARIANT_CPU)
            assert np.array_equal(got, want)
            # the CUDA variant (strict >, exchange-sort hull) agrees here: no IoU sits on a threshold
            assert np.array_equal(oracle.nms_rotated_keep(d, order, thr, oracle.VARIANT_CUDA), want)


def test_gt_vs_ge_delta():
    """The one documented CPU/CUDA difference: `>=` vs `>` (nms_rotated.py:444 vs :403-404)."""
    d = np.array([[0, 0, 2, 2, 0], [0, 0, 2, 1, 0]], np.float32)        # IoU exactly 0.5
    order = np.array([0, 1], np.int32)
    assert oracle.single_iou(d[0], d[1]) == 0.5
    assert list(oracle.nms_rotated_keep(d, order, 0.5, oracle.VARIANT_CPU)) == [True, False]
    assert list(oracle.nms_rotated_keep(d, order, 0.5, oracle.VARIANT_CUDA)) == [True, True]


@pytest.mark.skipif(oracle.ref_cpu() is None, reason="oracle/_ref not built on this box")
def test_against_compiled_reference_random():
    rng = np.random.default_rng(7)
    RC = oracle.ref_cpu()
    b1 = np.concatenate([dota_boxes(rng, 150, 300.0), clustered_boxes(rng, 100, 10, 300.0), ADVERSARIAL])
    b2 = np.concatenate([dota_boxes(rng, 120, 300.0), b1[:60]])
    for ver, fn in ((0, RC.ref_box_iou_rotated_cpu), (1, RC.ref_box_iou_rotated_v1_cpu)):
        want = np.zeros((len(b1), len(b2)), np.float32)
        fn(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2), want.ctypes.data_as(fp))
        got = np.zeros_like(want)
        oracle.lib().orc_box_iou_rotated(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2),
                                         got.ctypes.data_as(fp), ver, oracle.VARIANT_CPU, 0)
        assert np.array_equal(bits(got), bits(want))
    # NMS via the reference nms_rotated_cpu source
    n = 400
    d = clustered_boxes(rng, n, 20, 200.0)
    d6 = np.concatenate([d, rng.integers(0, 3, n).astype(np.float32)[:, None]], 1).astype(np.float32)
    order = oracle.argsort_desc(tie_free_scores(rng, n))
    keep, sup = np.zeros(n, np.bool_), np.zeros(n, np.uint8)
    RC.ref_nms_rotated_cpu(d6.ctypes.data_as(fp), n, 6, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                           ctypes.c_float(0.2), sup.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                           keep.ctypes.data_as(ctypes.POINTER(ctypes.c_bool)))
    assert np.array_equal(keep, oracle.nms_rotated_keep(d6, order, 0.2, oracle.VARIANT_CPU))


def test_reference_hull_overshoot_pair():
    """The reference's own arithmetic overshoots the true IoU on this lattice pair (b lies inside a with two collinear
    edges: true IoU 0.25): both of its builds return 0.26354942 in (a, b) order and 0.25 in (b, a) order.  The oracle
    must reproduce that, because the product's NMS pruning exemption for parallel pairs exists for exactly this."""
    a = np.array([3, 2, 4, 3, 3 * np.pi / 4], np.float32)
    b = np.array([4, 1, 3, 1, -3 * np.pi / 4], np.float32)
    for variant in (oracle.VARIANT_CPU, oracle.VARIANT_CUDA):
        assert np.float32(oracle.single_iou(a, b, 0, variant)) == np.float32(0.26354942)
        assert np.float32(oracle.single_iou(b, a, 0, variant)) == np.float32(0.25)
    RC = oracle.ref_cpu()
    if RC is not None:   # the reference's compiled CPU source says the same
        assert np.float32(RC.ref_single_iou_v0_cpu(a.ctypes.data_as(fp), b.ctypes.data_as(fp))) == np.float32(0.26354942)
        assert np.float32(RC.ref_single_iou_v0_cpu(b.ctypes.data_as(fp), a.ctypes.data_as(fp))) == np.float32(0.25)
    RG = oracle.ref_cuda()
    if RG is not None and hasattr(RG, "ref_single_iou_v0_cudavariant_host"):   # and so does its CUDA-variant source run on the host
        f = RG.ref_single_iou_v0_cudavariant_host
        f.restype = ctypes.c_float
        assert np.float32(f(a.ctypes.data_as(fp), b.ctypes.data_as(fp))) == np.float32(0.26354942)


def test_v1_small_box_zeroing():
    b = np.array([[0, 0, 4, 4, 0.1], [1, 1, 5e-4, 4, 0.2], [0.5, 0, 3, 3, -0.1]], np.float32)
    out = oracle.box_iou_rotated(b, b, version=1)
    assert (out[1] == 0).all() and (out[:, 1] == 0).all() and out[0, 2] > 0.3


def test_sampling_ops_shapes_and_identities():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 16, 16)).astype(np.float32)
    # a RoI exactly covering integer pixel centres: v1, sampling grid lands on pixel centres
    rois = np.array([[0, 8.0, 8.0, 8.0, 8.0, 0.0], [1, 5.0, 9.0, 6.0, 3.0, 0.7]], np.float32)
    out = oracle.roi_align_rotated(x, rois, 4, 1.0, 2, version=1)
    assert out.shape == (2, 4, 4, 4) and np.isfinite(out).all()
    out0 = oracle.roi_align_rotated(x, rois, (4, 4), 1.0, 2, version=0)
    assert not np.allclose(out, out0)
    # feature_refine with the box centred on its own pixel (row=bbox[0], col=bbox[1]) doubles the map
    N, C, H, W = 1, 3, 8, 8
    f = rng.standard_normal((N, C, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    boxes = np.zeros((N, H, W, 5), np.float32)
    boxes[0, :, :, 0] = ys
    boxes[0, :, :, 1] = xs
    assert np.allclose(oracle.feature_refine(f, boxes, 1.0, 1), 2 * f, atol=1e-6)
    # AlignConv with axis-aligned anchors of size 3*stride == a plain 3x3 conv (offsets all zero)
    from _inputs import s2anet_anchors
    stride = 8
    a = s2anet_anchors(rng, 1, 8, 8, stride, jitter=False)
    a[..., 2:4] = 3 * stride
    off = oracle.align_conv_offset(a, stride)
    assert np.abs(off).max() < 1e-5
    w = rng.standard_normal((5, 3, 3, 3)).astype(np.float32) * 0.1
    got = oracle.align_conv(f, a, stride, w)
    import torch
    want = torch.relu(torch.nn.functional.conv2d(torch.from_numpy(f), torch.from_numpy(w), padding=1)).numpy()
    assert np.allclose(got, want, atol=1e-5)

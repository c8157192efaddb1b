"""GPU parity tests: every C-ABI entry point against the oracle (oracle/oracle.cpp), the committed
golden vectors and — where oracle/_ref/libref_cuda.so travelled to the box — the reference's own
CUDA kernels on the same GPU.

Bars (BASELINE.json north_star): IoU and NMS bit-exact against the oracle's CUDA-variant
arithmetic; sampled features <= 1e-4 abs."""
import os

import numpy as np
import pytest
import torch

import oracle
import _refcuda
from _inputs import ADVERSARIAL, clustered_boxes, dota_boxes, s2anet_anchors, tie_free_scores

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4   # abs tolerance for sampled features / IoU floats (north_star)


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def ops():
    import jdet_b200.ops as o
    return o


# ------------------------------------------------------------------------------------ IoU
@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("shape", [(700, 900), (257, 131), (1, 1), (65, 4)])
def test_iou_bit_exact_vs_oracle(version, shape):
    rng = np.random.default_rng(11 + version)
    n1, n2 = shape
    b1 = np.concatenate([dota_boxes(rng, n1, 400.0), clustered_boxes(rng, n1 // 2 + 1, 10, 400.0), ADVERSARIAL])[:max(n1, 1)]
    b2 = np.concatenate([ADVERSARIAL, dota_boxes(rng, n2, 400.0), b1[: n1 // 3]])[:max(n2, 1)]
    fn = ops().box_iou_rotated if version == 0 else ops().box_iou_rotated_v1
    got = fn(cu(b1), cu(b2)).cpu().numpy()
    want = oracle.box_iou_rotated(b1, b2, version, oracle.VARIANT_CUDA)
    bad = np.nonzero(bits(got) != bits(want))
    assert len(bad[0]) == 0, (len(bad[0]), got[bad][:5], want[bad][:5])


def _lattice_boxes(rng, n):
    """Integer-lattice boxes at multiples of 45 degrees: shared corners, collinear / parallel edges, exact zeros in the
    edge tests, up to 24 (duplicate) clip points — the exact routine's fallback and slot-overflow paths, and the pairs
    on which the reference's own hull overshoots the true IoU."""
    return np.stack([rng.integers(0, 6, n), rng.integers(0, 6, n), rng.integers(1, 5, n), rng.integers(1, 5, n),
                     rng.integers(-4, 5, n) * (np.pi / 4)], 1).astype(np.float32)


@pytest.mark.parametrize("version", [0, 1])
def test_iou_dense_tiles_and_degenerate_pairs(version):
    """Tiles whose 8192 pairs ALL survive the reject stages (several rounds of the 4096-entry survivor queue) and
    lattice boxes (comparison-only unit tests leave their ordinary range; more clip points than shared-memory slots)."""
    rng = np.random.default_rng(21 + version)
    heap = clustered_boxes(rng, 450, 450, 100.0)              # one cluster: every pair overlaps
    lat = _lattice_boxes(rng, 300)
    fn = ops().box_iou_rotated if version == 0 else ops().box_iou_rotated_v1
    for b1, b2 in ((heap[:192], heap[192:]), (lat[:130], lat), (lat, heap[:140])):
        got = fn(cu(b1), cu(b2)).cpu().numpy()
        want = oracle.box_iou_rotated(b1, b2, version, oracle.VARIANT_CUDA)
        bad = np.nonzero(bits(got) != bits(want))
        assert len(bad[0]) == 0, (len(bad[0]), got[bad][:5], want[bad][:5])


def test_iou_candidate_queue_overflow():
    """4200 x 4200 boxes of ONE cluster: 17.6 M overlapping pairs, more than the 16 M slots of the device-wide
    candidate queue, so the last tile CTAs evaluate their own candidates in place (compact routine) and the exact
    kernel skips the reserved-but-unused slots.  Bit-exact against the oracle all the same."""
    rng = np.random.default_rng(3)
    b = clustered_boxes(rng, 4200, 4200, 100.0)
    got = ops().box_iou_rotated(cu(b), cu(b)).cpu().numpy()
    want = oracle.box_iou_rotated(b, b, 0, oracle.VARIANT_CUDA)
    assert (want > 0).mean() > 0.99
    bad = np.nonzero(bits(got) != bits(want))
    assert len(bad[0]) == 0, (len(bad[0]), got[bad][:5], want[bad][:5])


def test_iou_golden_and_known_answers():
    g = np.load(os.path.join(GOLD, "ref_cpu_iou.npz"))
    b1, b2 = cu(g["boxes1"]), cu(g["boxes2"])
    for ver, fn in ((0, ops().box_iou_rotated), (1, ops().box_iou_rotated_v1)):
        got = fn(b1, b2).cpu().numpy()
        want = g["iou_v%d_cudavariant" % ver].copy()
        if ver == 1:   # the fixture holds the raw kernel output; the Python post-pass zeroes small boxes
            want[np.minimum(g["boxes1"][:, 2], g["boxes1"][:, 3]) < 1e-3, :] = 0
            want[:, np.minimum(g["boxes2"][:, 2], g["boxes2"][:, 3]) < 1e-3] = 0
        assert np.array_equal(bits(got), bits(want))
        # reference CPU build (std::sort hull): same values up to hull tie-breaking
        assert np.abs(got - np.where(want == 0, 0, g["iou_v%d_cpu" % ver])).max() <= 1e-6
    k = np.load(os.path.join(GOLD, "known_answers.npz"))
    got = ops().box_iou_rotated(cu(k["k1_boxes"]), cu(k["k1_boxes"])).cpu().numpy()
    assert np.array_equal(bits(got), bits(k["k1_iou"]))
    got = ops().box_iou_rotated(cu(k["k3_a"][None]), cu(k["k3_b"][None])).cpu().numpy()
    assert got[0, 0] == k["k3_iou"]


def test_iou_empty_and_dtype():
    o = ops()
    e = torch.zeros((0, 5), device="cuda")
    b = cu(ADVERSARIAL)
    assert o.box_iou_rotated(e, b).shape == (0, len(ADVERSARIAL))
    assert o.box_iou_rotated(b, e).shape == (len(ADVERSARIAL), 0)
    with pytest.raises(AssertionError):
        o.box_iou_rotated(b, b.double())


@pytest.mark.skipif(not _refcuda.available(), reason="oracle/_ref/libref_cuda.so not present")
@pytest.mark.parametrize("version", [0, 1])
def test_iou_vs_reference_cuda_kernel(version):
    rng = np.random.default_rng(5)
    b1 = cu(np.concatenate([dota_boxes(rng, 1000), clustered_boxes(rng, 500, 10)]))
    b2 = cu(np.concatenate([dota_boxes(rng, 1000), clustered_boxes(rng, 300, 10)]))
    fn = ops().box_iou_rotated if version == 0 else ops().box_iou_rotated_v1
    got, ref = fn(b1, b2), _refcuda.box_iou_rotated(b1, b2, version)
    # the reference kernel is compiled with nvcc's default FMA contraction: last-bit differences only
    assert (got - ref).abs().max().item() <= 1e-5
    assert ((got > 0) != (ref > 0)).sum().item() == 0


def test_iou_full_size_properties():
    """16k x 16k (1 GiB of IoUs): size-independent properties + spot checks against the oracle."""
    rng = np.random.default_rng(0)
    n = 16384
    b = dota_boxes(rng, n)
    tb = cu(b)
    m = ops().box_iou_rotated(tb, tb)
    d = m.diagonal()
    assert (d - 1).abs().max().item() <= 2e-6                      # IoU(b,b) = 1
    assert m.min().item() >= 0 and m.max().item() <= 1 + 1e-6
    assert (m - m.t()).abs().max().item() <= 1e-5                   # symmetric up to rounding
    ii, jj = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    nz = torch.nonzero(m[:2048] > 0).cpu().numpy()[:4000]
    ii, jj = np.concatenate([ii, nz[:, 0]]), np.concatenate([jj, nz[:, 1]])
    got = m[torch.as_tensor(ii).cuda(), torch.as_tensor(jj).cuda()].cpu().numpy()
    want = np.array([oracle.single_iou(b[i], b[j]) for i, j in zip(ii, jj)], np.float32)
    assert np.array_equal(bits(got), bits(want))


# ------------------------------------------------------------------------------------ NMS
@pytest.mark.parametrize("thr", [0.1, 0.26, 0.5])
def test_nms_lattice_boxes_where_the_reference_hull_overshoots(thr):
    """Parallel / perpendicular lattice pairs are exempt from the IoU-upper-bound pruning because the reference's IoU
    can exceed the true one there (0.2635 vs 0.25 straddles thr = 0.26): keep flags must still equal the oracle's."""
    rng = np.random.default_rng(9)
    d = np.concatenate([_lattice_boxes(rng, 900), np.array([[3, 2, 4, 3, 3 * np.pi / 4], [4, 1, 3, 1, -3 * np.pi / 4]], np.float32)])
    s = tie_free_scores(rng, len(d))
    s[-2], s[-1] = 2.0, 1.5                                   # the overshooting pair first, in (a, b) order
    l = rng.integers(0, 2, len(d)).astype(np.int64)
    l[-2:] = 0
    got = ops().nms_rotated.ml_nms_rotated(cu(d), cu(s), cu(l, torch.int64), thr).cpu().numpy()
    assert np.array_equal(got, oracle.ml_nms_rotated(d, s, l, thr, oracle.VARIANT_CUDA))
    got5 = ops().nms_rotated.nms_rotated(cu(d), cu(s), thr).cpu().numpy()
    assert np.array_equal(got5, oracle.nms_rotated(d, s, thr, oracle.VARIANT_CUDA))


def test_nms_candidate_queue_overflow():
    """4400 boxes of one cluster, one class, thr 0.05: ~9.7 M same-class pairs that survive every reject stage, more
    than the 8 M slots of the exact-IoU candidate queue — the mask kernel then evaluates its own candidates."""
    rng = np.random.default_rng(4)
    d, s = clustered_boxes(rng, 4400, 4400, 100.0), tie_free_scores(rng, 4400)
    for thr in (0.05, 0.3, 0.6):   # (the higher thresholds prune more pairs before the queue and keep more boxes)
        got = ops().nms_rotated.nms_rotated(cu(d), cu(s), thr).cpu().numpy()
        assert np.array_equal(got, oracle.nms_rotated(d, s, thr, oracle.VARIANT_CUDA)), thr


def _nms_case(rng, n, ncls, clustered=True):
    d = clustered_boxes(rng, n, 25, 512.0) if clustered else dota_boxes(rng, n, 512.0)
    return d, tie_free_scores(rng, n), rng.integers(0, ncls, n).astype(np.int64)


@pytest.mark.parametrize("thr", [0.1, 0.3, 0.5])
@pytest.mark.parametrize("n,ncls", [(3000, 5), (777, 1), (130, 40), (64, 2), (65, 3), (1, 1)])
def test_ml_nms_bit_exact_vs_oracle(n, ncls, thr):
    rng = np.random.default_rng(n + ncls)
    d, s, l = _nms_case(rng, n, ncls)
    got = ops().nms_rotated.ml_nms_rotated(cu(d), cu(s), cu(l, torch.int64), thr).cpu().numpy()
    want = oracle.ml_nms_rotated(d, s, l, thr, oracle.VARIANT_CUDA)
    assert np.array_equal(got, want)
    got5 = ops().nms_rotated.nms_rotated(cu(d), cu(s), thr).cpu().numpy()
    assert np.array_equal(got5, oracle.nms_rotated(d, s, thr, oracle.VARIANT_CUDA))


def test_nms_golden_and_k2():
    g = np.load(os.path.join(GOLD, "ref_cpu_nms.npz"))
    d6, order = g["dets6"], cu(g["order"], torch.int32)
    nr = ops().nms_rotated
    for thr in (0.1, 0.3, 0.5):
        for bl, d in ((5, d6[:, :5]), (6, d6)):
            keep = nr.nms_rotated_cuda(cu(d), order, thr, box_length=bl).cpu().numpy()
            assert np.array_equal(keep, g["keep%d_cpu_thr%02d" % (bl, int(thr * 10))])
    k = np.load(os.path.join(GOLD, "known_answers.npz"))
    assert nr.ml_nms_rotated(cu(k["k2_dets"]), cu(k["k2_scores"]), cu(k["k2_labels"], torch.int64), 0.3).tolist() == [2]
    assert nr.nms_rotated(cu(k["k2_dets"]), cu(k["k2_scores"]), 0.3).tolist() == [2]
    assert nr.ml_nms_rotated(cu(k["k2_dets"]), cu(k["k2_scores"]), cu([0, 1, 2], torch.int64), 0.3).tolist() == [0, 1, 2]


def test_nms_strict_threshold_and_corner_cases():
    nr = ops().nms_rotated
    d = cu([[0, 0, 2, 2, 0], [0, 0, 2, 1, 0]])                   # IoU exactly 0.5
    s = cu([0.9, 0.8])
    assert nr.nms_rotated(d, s, 0.5).tolist() == [0, 1]          # strict >, the CUDA path (nms_rotated.py:403-404)
    assert nr.nms_rotated(d, s, 0.4999).tolist() == [0]
    # negative threshold: 0 > thr, so even disjoint / cross-class boxes suppress (reference semantics)
    rng = np.random.default_rng(1)
    dd, ss, ll = _nms_case(rng, 200, 3)
    got = nr.ml_nms_rotated(cu(dd), cu(ss), cu(ll, torch.int64), -0.5).cpu().numpy()
    assert np.array_equal(got, oracle.ml_nms_rotated(dd, ss, ll, -0.5, oracle.VARIANT_CUDA))
    assert len(got) == 1
    # ties in score: stable order, lower index first
    st = np.full(200, 0.5, np.float32)
    got = nr.ml_nms_rotated(cu(dd), cu(st), cu(ll, torch.int64), 0.2).cpu().numpy()
    assert np.array_equal(got, oracle.ml_nms_rotated(dd, st, ll, 0.2, oracle.VARIANT_CUDA))
    # more distinct labels than grouping keys (8-bit hash), non-integer and negative label values:
    # boxes that share a key but not a label must still never interact
    d4 = clustered_boxes(rng, 3000, 60, 300.0)
    lab = rng.choice(np.concatenate([np.arange(0, 700, dtype=np.float32), np.array([-3.0, 0.5, 1e9, -0.0], np.float32),
                                     (np.arange(40) + 0.25).astype(np.float32)]), 3000).astype(np.float32)
    d6 = np.concatenate([d4, lab[:, None]], 1).astype(np.float32)
    order = oracle.argsort_desc(tie_free_scores(rng, 3000))
    got = nr.nms_rotated_cuda(cu(d6), cu(order, torch.int32), 0.15, box_length=6).cpu().numpy()
    assert np.array_equal(got, oracle.nms_rotated_keep(d6, order, 0.15, oracle.VARIANT_CUDA))
    # degenerate boxes never suppress nor get suppressed
    deg = np.array([[5, 5, 0, 0, 0], [5, 5, 0, 0, 0], [5, 5, 3, 3, 0], [5, 5, 3, 3, 0.01]], np.float32)
    assert nr.nms_rotated(cu(deg), cu([0.9, 0.8, 0.7, 0.6]), 0.1).tolist() == [0, 1, 2]


def test_multiclass_nms_rotated_matches_oracle():
    rng = np.random.default_rng(21)
    n, C = 1500, 15
    boxes = clustered_boxes(rng, n, 30, 512.0)
    scores = rng.random((n, C + 1)).astype(np.float32) ** 4
    nr = ops().nms_rotated
    for max_num in (-1, 2000, 100):
        gd, gl = nr.multiclass_nms_rotated(cu(boxes), cu(scores), 0.05, dict(type="nms_rotated", iou_thr=0.1), max_num)
        wd, wl = oracle.multiclass_nms_rotated(boxes, scores, 0.05, dict(iou_thr=0.1), max_num)
        assert gd.shape == wd.shape and np.array_equal(gd.cpu().numpy(), wd) and np.array_equal(gl.cpu().numpy(), wl)
    # per-class boxes (n, 5*(C+1)) layout and the empty return
    mb = np.tile(boxes, (1, C + 1)).astype(np.float32)
    gd, gl = nr.multiclass_nms_rotated(cu(mb), cu(scores), 0.05, dict(iou_thr=0.1), 50)
    wd, wl = oracle.multiclass_nms_rotated(mb, scores, 0.05, dict(iou_thr=0.1), 50)
    assert np.array_equal(gd.cpu().numpy(), wd) and np.array_equal(gl.cpu().numpy(), wl)
    gd, gl = nr.multiclass_nms_rotated(cu(boxes), cu(scores), 2.0, dict(iou_thr=0.1))
    assert gd.shape == (0, 6) and gl.shape == (0,)


@pytest.mark.skipif(not _refcuda.available(), reason="oracle/_ref/libref_cuda.so not present")
def test_nms_vs_reference_cuda_kernel_100k():
    """BASELINE cfg3: 100k proposals x 15 classes, thr 0.1 — keep mask equal to the reference CUDA path."""
    rng = np.random.default_rng(0)
    n = 100000
    d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
    s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
    d6 = cu(np.concatenate([d, l[:, None].astype(np.float32)], 1))
    nr = ops().nms_rotated
    order = nr.argsort_desc(cu(s))
    assert np.array_equal(order.cpu().numpy(), oracle.argsort_desc(s))
    keep = nr.nms_rotated_cuda(d6, order, 0.1, box_length=6).cpu().numpy()
    ref, _ = _refcuda.nms_rotated_keep(d6, order, 0.1)
    assert np.array_equal(keep, ref), int((keep != ref).sum())
    # size-independent properties: survivors of one class are pairwise below threshold;
    # NMS is idempotent on its own output
    kept = np.nonzero(keep)[0]
    kd, ks, kl = d[kept], s[kept], l[kept]
    again = nr.ml_nms_rotated(cu(kd), cu(ks), cu(kl, torch.int64), 0.1).cpu().numpy()
    assert np.array_equal(again, np.arange(len(kept)))
    c0 = kd[kl == 0][:4000]
    m = ops().box_iou_rotated(cu(c0), cu(c0))
    m.fill_diagonal_(0)
    assert m.max().item() <= 0.1 + 1e-6


# ------------------------------------------------------------------------------------ RoIAlign
def _rois(rng, R, B, extent, lo=8, hi=256):
    b = dota_boxes(rng, R, extent, lo, hi)
    return np.concatenate([rng.integers(0, B, (R, 1)).astype(np.float32), b], 1)


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("cfg", [
    dict(B=2, C=64, H=40, W=56, R=300, out=(7, 7), sr=2, scale=0.25),      # staged (channel-last) path
    dict(B=1, C=128, H=32, W=32, R=200, out=(5, 3), sr=3, scale=0.125),    # staged, non-square bins
    dict(B=1, C=256, H=24, W=24, R=150, out=(4, 4), sr=1, scale=0.25),     # staged, whole-RoI CTAs, 4 slots/bin, even bin count
    dict(B=2, C=256, H=32, W=32, R=260, out=(6, 6), sr=2, scale=0.125),    # staged, even bin count (padded smem rows)
    dict(B=1, C=128, H=20, W=20, R=120, out=(2, 2), sr=4, scale=0.25),     # staged, 64 slots/bin (no tap merging)
    dict(B=2, C=24, H=40, W=56, R=100, out=(7, 7), sr=2, scale=0.25),      # direct path (C % 64 != 0)
    dict(B=2, C=64, H=64, W=64, R=8, out=(7, 7), sr=2, scale=0.0625),      # direct path (few RoIs)
    dict(B=1, C=16, H=48, W=48, R=40, out=(7, 7), sr=0, scale=0.25),       # adaptive grid (sampling_ratio 0)
])
def test_roi_align_vs_oracle(version, cfg):
    rng = np.random.default_rng(cfg["R"] + version)
    x = rng.standard_normal((cfg["B"], cfg["C"], cfg["H"], cfg["W"])).astype(np.float32)
    extent = cfg["W"] / cfg["scale"]
    rois = _rois(rng, cfg["R"], cfg["B"], extent, 8, extent / 2)
    rois[:5, 1:3] = [[-20, -20], [extent + 30, 5], [0, 0], [extent, extent], [extent / 2, -3]]   # border / outside
    rois[5, 3:5] = [1.0, 1.0]                                                                    # malformed -> 1x1
    mod = ops().roi_align_rotated_v1 if version == 1 else ops().roi_align_rotated
    got = mod.roi_align(cu(x), cu(rois), cfg["out"], cfg["scale"], cfg["sr"]).cpu().numpy()
    want = oracle.roi_align_rotated(x, rois, cfg["out"], cfg["scale"], cfg["sr"], version)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()
    if _refcuda.available():
        ref = _refcuda.roi_align_rotated(cu(x), cu(rois), cfg["out"], cfg["scale"], cfg["sr"], version).cpu().numpy()
        assert np.abs(got - ref).max() <= TOL, np.abs(got - ref).max()


@pytest.mark.parametrize("version", [0, 1])
def test_roi_align_channels_last_input(version):
    """A torch.channels_last map takes the no-relayout entry point (jdet_roi_align_rotated_nhwc); same numbers as NCHW.
    A shape the channel-last kernel refuses (C % 64 != 0) silently takes the NCHW entry point."""
    rng = np.random.default_rng(31)
    for C, sr in ((256, 2), (64, 2), (24, 2)):
        x = rng.standard_normal((2, C, 40, 48)).astype(np.float32)
        rois = _rois(rng, 300, 2, 160.0, 4, 96)
        mod = ops().roi_align_rotated_v1 if version == 1 else ops().roi_align_rotated
        xcl = cu(x).contiguous(memory_format=torch.channels_last)
        assert not xcl.is_contiguous()
        got = mod.roi_align(xcl, cu(rois), (7, 7), 0.25, sr).cpu().numpy()
        want = oracle.roi_align_rotated(x, rois, (7, 7), 0.25, sr, version)
        assert np.abs(got - want).max() <= TOL
        assert np.array_equal(got, mod.roi_align(cu(x), cu(rois), (7, 7), 0.25, sr).cpu().numpy()) or C == 24


def test_roi_align_module_and_reference_selftest_shape():
    """ops/roi_align_rotated_v1.py:376-386: feature (2,1024,64,64), the two literal RoIs, 7x7, 1/16."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 1024, 64, 64)).astype(np.float32)
    rois = np.array([[0, 20, 120, 80, 195.5, 0.3], [1, 23, 56, 200, 300.5, 0.2]], np.float32)
    for mod, cls, ver in ((ops().roi_align_rotated_v1, "ROIAlignRotated_v1", 1), (ops().roi_align_rotated, "ROIAlignRotated", 0)):
        layer = getattr(mod, cls)((7, 7), 1 / 16.)
        got = layer(cu(x), cu(rois)).cpu().numpy()
        want = oracle.roi_align_rotated(x, rois, (7, 7), 1 / 16., 0, ver)
        assert got.shape == (2, 1024, 7, 7) and np.abs(got - want).max() <= TOL
        assert layer.execute(cu(x), cu(rois)).shape == (2, 1024, 7, 7)


@pytest.mark.parametrize("shape", [dict(B=2, C=256, H=48, W=64, R=600, out=(7, 7), sr=2, scale=0.25),
                                   dict(B=1, C=64, H=40, W=40, R=300, out=(6, 6), sr=2, scale=0.125),
                                   dict(B=2, C=128, H=32, W=32, R=200, out=(5, 3), sr=3, scale=0.125)])
def test_roi_align_tma_staged_path(shape, monkeypatch):
    """JDET_ROI_TMA=1: small RoIs go through the tensor-map gather4 staging kernel (roi_gather_tma.cuh), the rest through
    the LSU kernel.  Same numbers as the default path bit for bit (same taps, same order of the per-bin sums)."""
    rng = np.random.default_rng(77)
    x = rng.standard_normal((shape["B"], shape["C"], shape["H"], shape["W"])).astype(np.float32)
    extent = shape["W"] / shape["scale"]
    rois = _rois(rng, shape["R"], shape["B"], extent, 4, extent / 3)
    rois[:4, 1:3] = [[-40, -40], [extent + 30, 5], [0, 0], [extent, extent]]          # outside / on the border
    for version, mod in ((1, ops().roi_align_rotated_v1), (0, ops().roi_align_rotated)):
        monkeypatch.delenv("JDET_ROI_TMA", raising=False)
        base = mod.roi_align(cu(x), cu(rois), shape["out"], shape["scale"], shape["sr"]).cpu().numpy()
        monkeypatch.setenv("JDET_ROI_TMA", "1")
        got = mod.roi_align(cu(x), cu(rois), shape["out"], shape["scale"], shape["sr"]).cpu().numpy()
        want = oracle.roi_align_rotated(x, rois, shape["out"], shape["scale"], shape["sr"], version)
        assert np.abs(got - want).max() <= TOL
        assert np.array_equal(bits(got), bits(base))


def test_roi_align_full_size_cfg2_tma_staged(monkeypatch):
    monkeypatch.setenv("JDET_ROI_TMA", "1")
    test_roi_align_full_size_cfg2()


def test_roi_align_full_size_cfg2():
    """BASELINE cfg2: 256-ch 256x256 map, 2048 RoIs, 7x7, sampling 2 — whole output vs the oracle."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 256, 256, 256)).astype(np.float32)
    rois = _rois(rng, 2048, 1, 1024.0)
    got = ops().roi_align_rotated_v1.roi_align(cu(x), cu(rois), (7, 7), 0.25, 2).cpu().numpy()
    want = oracle.roi_align_rotated(x, rois, (7, 7), 0.25, 2, 1)
    err = np.abs(got - want)
    assert err.max() <= TOL, err.max()


# ------------------------------------------------------------------------------------ feature_refine
@pytest.mark.parametrize("points", [1, 5])
@pytest.mark.parametrize("W", [40, 30])        # W % 4 == 0 takes the 16-B vector kernel for points=1
def test_feature_refine_vs_oracle(points, W):
    rng = np.random.default_rng(points)
    N, C, H, stride = 2, 48, 32, 8.0
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    a = s2anet_anchors(rng, N, H, W, stride)
    boxes = a[..., [1, 0, 2, 3, 4]].copy()       # the op reads bbox[0] as the row coordinate
    boxes[0, 0, :8, :2] = [[-50, -50]] * 4 + [[1e4, 3]] * 4     # out-of-map samples
    boxes[1, 3, :6, :2] = [[(H - 1) * stride, 2 * stride], [0, (W - 1) * stride]] * 3   # in-map but far from the pixel's row band
    boxes[1, 20, 5:9, 2:4] *= 6                                  # huge boxes: corner samples leave the band
    got = ops().fr.feature_refine(cu(x), cu(boxes), 1 / stride, points).cpu().numpy()
    want = oracle.feature_refine(x, boxes, 1 / stride, points)
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()
    if _refcuda.available():
        ref = _refcuda.feature_refine(cu(x), cu(boxes), 1 / stride, points).cpu().numpy()
        assert np.abs(got - ref).max() <= TOL
    fr = ops().fr.FR(1 / stride, points)
    assert np.array_equal(fr(cu(x), cu(boxes)).cpu().numpy(), got)


def test_feature_refine_module():
    rng = np.random.default_rng(9)
    m = ops().fr.FeatureRefineModule(16, [8, 16]).cuda()
    feats = [cu(rng.standard_normal((2, 16, 16, 16))), cu(rng.standard_normal((2, 16, 8, 8)))]
    boxes = [[cu(s2anet_anchors(rng, 1, 16, 16, 8).reshape(-1, 5)), cu(s2anet_anchors(rng, 1, 8, 8, 16).reshape(-1, 5))]
             for _ in range(2)]
    with torch.no_grad():
        out = m(feats, boxes)
    assert [tuple(o.shape) for o in out] == [(2, 16, 16, 16), (2, 16, 8, 8)]


# ------------------------------------------------------------------------------------ backward
BWD_TOL = 2e-4   # float atomics: summation order differs from the fp64-accumulating oracle (and run to run)


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("cfg", [
    dict(B=2, C=64, H=40, W=56, R=300, out=(7, 7), sr=2, scale=0.25),      # staged: channel-last scratch + vector atomics
    dict(B=1, C=128, H=32, W=32, R=200, out=(5, 3), sr=3, scale=0.125),
    dict(B=1, C=256, H=24, W=24, R=150, out=(4, 4), sr=1, scale=0.25),
    dict(B=2, C=24, H=40, W=56, R=60, out=(7, 7), sr=2, scale=0.25),       # direct NCHW atomics
    dict(B=1, C=16, H=48, W=48, R=30, out=(7, 7), sr=0, scale=0.25),       # adaptive grid
])
def test_roi_align_backward(version, cfg):
    rng = np.random.default_rng(cfg["R"] + 10 * version)
    x = rng.standard_normal((cfg["B"], cfg["C"], cfg["H"], cfg["W"])).astype(np.float32)
    extent = cfg["W"] / cfg["scale"]
    rois = _rois(rng, cfg["R"], cfg["B"], extent, 8, extent / 2)
    rois[:3, 1:3] = [[-20, -20], [extent + 30, 5], [0, 0]]
    go = rng.standard_normal((cfg["R"], cfg["C"]) + cfg["out"]).astype(np.float32)
    mod = ops().roi_align_rotated_v1 if version == 1 else ops().roi_align_rotated
    xt = cu(x).requires_grad_(True)
    out = mod.roi_align(xt, cu(rois), cfg["out"], cfg["scale"], cfg["sr"])
    out.backward(cu(go))
    got = xt.grad.cpu().numpy()
    want = oracle.roi_align_rotated_backward(go, rois, x.shape, cfg["scale"], cfg["sr"], version)
    assert np.abs(got - want).max() <= BWD_TOL, np.abs(got - want).max()
    if _refcuda.available():
        ref = _refcuda.roi_align_rotated_backward(cu(go), cu(rois), x.shape, cfg["scale"], cfg["sr"], version).cpu().numpy()
        assert np.abs(got - ref).max() <= BWD_TOL
    # forward under autograd is the same forward
    assert np.array_equal(out.detach().cpu().numpy(),
                          mod.roi_align(cu(x), cu(rois), cfg["out"], cfg["scale"], cfg["sr"]).cpu().numpy())


@pytest.mark.parametrize("points", [1, 5])
def test_feature_refine_backward(points):
    rng = np.random.default_rng(40 + points)
    N, C, H, W, stride = 2, 24, 20, 28, 8.0
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    a = s2anet_anchors(rng, N, H, W, stride)
    boxes = a[..., [1, 0, 2, 3, 4]].copy()
    boxes[0, 0, :4, :2] = [[-50, -50]] * 4
    go = rng.standard_normal(x.shape).astype(np.float32)
    xt = cu(x).requires_grad_(True)
    ops().fr.feature_refine(xt, cu(boxes), 1 / stride, points).backward(cu(go))
    got = xt.grad.cpu().numpy()
    want = oracle.feature_refine_backward(go, boxes, 1 / stride, points)
    assert np.abs(got - want).max() <= BWD_TOL, np.abs(got - want).max()
    if _refcuda.available():
        ref = _refcuda.feature_refine_backward(cu(go), cu(boxes), 1 / stride, points).cpu().numpy()
        assert np.abs(got - ref).max() <= BWD_TOL


# ------------------------------------------------------------------------------------ DeformConv / AlignConv
@pytest.mark.parametrize("cfg", [
    dict(B=2, C=16, H=12, W=14, Co=24, k=3, stride=1, pad=1, dil=1, groups=1, dg=1),
    dict(B=1, C=8, H=15, W=13, Co=12, k=3, stride=2, pad=1, dil=2, groups=1, dg=2),
    dict(B=3, C=6, H=9, W=9, Co=70, k=1, stride=1, pad=0, dil=1, groups=1, dg=1),
])
def test_deform_conv_vs_oracle(cfg):
    rng = np.random.default_rng(cfg["Co"])
    B, C, H, W, Co, k = cfg["B"], cfg["C"], cfg["H"], cfg["W"], cfg["Co"], cfg["k"]
    Ho = (H + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    Wo = (W + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((B, cfg["dg"] * 2 * k * k, Ho, Wo)) * 2).astype(np.float32)
    w = (rng.standard_normal((Co, C, k, k)) * 0.2).astype(np.float32)
    got = ops().dcn_v1.deform_conv(cu(x), cu(off), cu(w), cfg["stride"], cfg["pad"], cfg["dil"], cfg["groups"],
                                   cfg["dg"]).cpu().numpy()
    want = oracle.deform_conv(x, off, w, cfg["stride"], cfg["pad"], cfg["dil"], cfg["dg"])
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()
    if _refcuda.available():
        ref = _refcuda.deform_conv(cu(x), cu(off), cu(w), cfg["stride"], cfg["pad"], cfg["dil"], cfg["dg"]).cpu().numpy()
        assert np.abs(got - ref).max() <= TOL


def test_deform_conv_groups_match_per_group_oracle():
    rng = np.random.default_rng(4)
    B, C, H, W, Co, g = 2, 8, 10, 10, 12, 2
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = rng.standard_normal((B, 18, H, W)).astype(np.float32)
    w = (rng.standard_normal((Co, C // g, 3, 3)) * 0.2).astype(np.float32)
    got = ops().dcn_v1.deform_conv(cu(x), cu(off), cu(w), 1, 1, 1, g, 1).cpu().numpy()
    for gi in range(g):
        want = oracle.deform_conv(x[:, gi * 4:(gi + 1) * 4], off, w[gi * 6:(gi + 1) * 6], 1, 1, 1, 1)
        assert np.abs(got[:, gi * 6:(gi + 1) * 6] - want).max() <= TOL


@pytest.mark.parametrize("shape", [(2, 64, 16, 16, 64, 8), (1, 256, 32, 32, 256, 16), (2, 32, 9, 7, 48, 32),
                                   (3, 64, 13, 11, 96, 8), (8, 128, 64, 64, 256, 16), (5, 32, 8, 8, 32, 128)])
def test_align_conv_vs_oracle(shape):
    N, C, H, W, Co, stride = shape
    rng = np.random.default_rng(C + H)
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    m = AlignConv(C, Co, 3).cuda()
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    a = s2anet_anchors(rng, N, H, W, stride)
    w = m.deform_conv.weight.detach().cpu().numpy()
    with torch.no_grad():
        got = m(cu(x), cu(a), stride).cpu().numpy()
    want = oracle.align_conv(x, a, stride, w)
    assert got.shape == (N, Co, H, W) and (got >= 0).all()
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()
    off = m.get_offset_batched(cu(a), stride).cpu().numpy()
    assert np.abs(off - oracle.align_conv_offset(a, stride)).max() <= 1e-4
    assert np.array_equal(m.get_offset(cu(a[0].reshape(-1, 5)), (H, W), stride).cpu().numpy(), off[0])


@pytest.mark.parametrize("cfg", [
    dict(B=2, C=16, H=12, W=14, Co=24, k=3, stride=1, pad=1, dil=1, dg=1),
    dict(B=3, C=8, H=15, W=13, Co=12, k=3, stride=2, pad=1, dil=2, dg=2),
])
def test_deform_conv_backward(cfg):
    rng = np.random.default_rng(cfg["Co"] + 1)
    B, C, H, W, Co, k = cfg["B"], cfg["C"], cfg["H"], cfg["W"], cfg["Co"], cfg["k"]
    Ho = (H + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    Wo = (W + 2 * cfg["pad"] - (cfg["dil"] * (k - 1) + 1)) // cfg["stride"] + 1
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    off = (rng.standard_normal((B, cfg["dg"] * 2 * k * k, Ho, Wo)) * 1.5).astype(np.float32)
    w = (rng.standard_normal((Co, C, k, k)) * 0.2).astype(np.float32)
    go = rng.standard_normal((B, Co, Ho, Wo)).astype(np.float32)
    xt, ot, wt = cu(x).requires_grad_(True), cu(off).requires_grad_(True), cu(w).requires_grad_(True)
    out = ops().dcn_v1.deform_conv(xt, ot, wt, cfg["stride"], cfg["pad"], cfg["dil"], 1, cfg["dg"])
    out.backward(cu(go))
    wx, wo_, ww = oracle.deform_conv_backward(x, off, w, go, cfg["stride"], cfg["pad"], cfg["dil"], cfg["dg"])
    assert np.abs(xt.grad.cpu().numpy() - wx).max() <= BWD_TOL
    assert np.abs(ot.grad.cpu().numpy() - wo_).max() <= 5e-4        # sums C products of O(1) terms in fp32
    assert np.abs(wt.grad.cpu().numpy() - ww).max() <= 5e-4
    if _refcuda.available():   # the reference's own col2im / col2im_coord kernels on the same column gradient
        colg = (cu(w).reshape(Co, -1).t() @ cu(go).permute(1, 0, 2, 3).reshape(Co, -1)).contiguous()
        r_x = _refcuda.deform_col2im(colg, cu(off), B, C, H, W, k, cfg["stride"], cfg["pad"], cfg["dil"], cfg["dg"]).cpu().numpy()
        r_o = _refcuda.deform_col2im_coord(colg, cu(x), cu(off), k, cfg["stride"], cfg["pad"], cfg["dil"], cfg["dg"]).cpu().numpy()
        assert np.abs(xt.grad.cpu().numpy() - r_x).max() <= 5e-4
        assert np.abs(ot.grad.cpu().numpy() - r_o).max() <= 5e-4


def test_align_conv_backward():
    rng = np.random.default_rng(77)
    N, C, H, W, Co, stride = 2, 32, 10, 12, 64, 8
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    m = AlignConv(C, Co, 3).cuda()
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    a = s2anet_anchors(rng, N, H, W, stride)
    go = rng.standard_normal((N, Co, H, W)).astype(np.float32)
    w = m.deform_conv.weight.detach().cpu().numpy()
    xt = cu(x).requires_grad_(True)
    out = m(xt, cu(a), stride)
    out.backward(cu(go))
    off = oracle.align_conv_offset(a, stride)
    fwd = oracle.deform_conv(x, off, w, 1, 1, 1, 1, relu=True)
    assert np.abs(out.detach().cpu().numpy() - fwd).max() <= TOL
    g = go * (fwd > 0)
    near = np.abs(oracle.deform_conv(x, off, w, 1, 1, 1, 1)) < 1e-4      # ReLU kink: sign may differ within tolerance
    g[near] = 0
    go2 = go.copy()
    go2[near] = 0
    xt2 = cu(x).requires_grad_(True)
    m.zero_grad()
    m(xt2, cu(a), stride).backward(cu(go2))
    wx, _, ww = oracle.deform_conv_backward(x, off, w, g, 1, 1, 1, 1)
    assert np.abs(xt2.grad.cpu().numpy() - wx).max() <= 5e-4
    assert np.abs(m.deform_conv.weight.grad.cpu().numpy() - ww).max() <= 2e-3   # sums N*H*W = 240 products per weight


# ------------------------------------------------------------------------------------ callers
def test_roi_extractor_and_iou_calculator():
    rng = np.random.default_rng(6)
    from jdet_b200.models.roi_extractors import OrientedSingleRoIExtractor, RboxSingleRoIExtractor
    from jdet_b200.models.boxes import BboxOverlaps2D_rotated, BboxOverlaps2D_rotated_v1
    strides = [4, 8, 16, 32]
    feats = [rng.standard_normal((2, 64, 256 // s, 256 // s)).astype(np.float32) for s in strides]
    rois = _rois(rng, 400, 2, 256.0, 8, 200)
    ext = OrientedSingleRoIExtractor(dict(type="ROIAlignRotated_v1", output_size=7, sampling_ratio=2), 64, strides,
                                     extend_factor=(1.4, 1.2))
    got = ext([cu(f) for f in feats], cu(rois)).cpu().numpy()
    r2 = rois.copy()
    r2[:, 3] *= 1.2
    r2[:, 4] *= 1.4
    lvl = np.clip(np.floor(np.log2(np.sqrt(r2[:, 3] * r2[:, 4]) / 56 + 1e-6)), 0, 3).astype(int)
    want = np.zeros_like(got)
    for i, s in enumerate(strides):
        if (lvl == i).any():
            want[lvl == i] = oracle.roi_align_rotated(feats[i], r2[lvl == i], 7, 1 / s, 2, 1)
    assert np.abs(got - want).max() <= TOL
    # the fused one-call path (jdet_roi_align_rotated_fpn) and the per-level path are the same kernels on the same tables
    fc = [cu(f) for f in feats]
    stretched = ext.roi_rescale(cu(rois), ext.extend_factor)
    per_level = ext._pool_by_level(fc, stretched, ext.map_roi_levels(stretched, 4)).cpu().numpy()
    assert np.array_equal(got, per_level)
    ext0 = RboxSingleRoIExtractor(dict(type="ROIAlignRotated", output_size=7, sampling_ratio=2), 64, strides)
    got0 = ext0(fc, cu(rois)).cpu().numpy()
    lvl0 = np.clip(np.floor(np.log2(np.sqrt(rois[:, 3] * rois[:, 4]) / 56 + 1e-6)), 0, 3).astype(int)
    want0 = np.zeros_like(got0)
    for i, s in enumerate(strides):
        if (lvl0 == i).any():
            want0[lvl0 == i] = oracle.roi_align_rotated(feats[i], rois[lvl0 == i], 7, 1 / s, 2, 0)
    assert got0.shape == (400, 64, 7, 7) and np.abs(got0 - want0).max() <= TOL
    b = dota_boxes(rng, 50, 200.0)
    b6 = np.concatenate([b, rng.random((50, 1)).astype(np.float32)], 1)
    assert np.array_equal(BboxOverlaps2D_rotated()(cu(b6), cu(b)).cpu().numpy(), oracle.box_iou_rotated(b, b, 0))
    assert np.array_equal(BboxOverlaps2D_rotated_v1()(cu(b), cu(b6)).cpu().numpy(), oracle.box_iou_rotated(b, b, 1))


def test_s2anet_head_forward_against_oracle_ops():
    """the forward-only S2ANetHead calls bbox_decode -> AlignConv -> ORConv2d/RIP -> multiclass NMS.
    The same layers are re-run with the two hot-path ops swapped for their oracle restatements."""
    from jdet_b200.models.roi_heads import S2ANetHead, bbox_decode
    torch.manual_seed(3)
    head = S2ANetHead(16, 256, test_cfg=dict(nms_pre=200, score_thr=0.3, nms=dict(type="nms_rotated", iou_thr=0.1),
                                              max_per_img=100)).cuda().eval()
    for m in head.modules():                       # the reference init (std 0.01) puts every score below score_thr
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.normal_(m.weight, 0, 0.05)
    torch.nn.init.normal_(head.or_conv.weight, 0, 0.05)
    torch.nn.init.normal_(head.align_conv.deform_conv.weight, 0, 0.03)
    torch.nn.init.constant_(head.odm_cls.bias, 0.0)
    strides = head.anchor_strides
    feats = [torch.randn(2, 256, 128 // s if 128 // s else 1, 128 // s if 128 // s else 1, device="cuda") for s in strides]
    res = head(feats)
    assert len(res) == 2
    # per-level: AlignConv inside the head == oracle AlignConv on the same refined anchors
    with torch.no_grad():
        x, s0 = feats[0], strides[0]
        f = x
        for conv in head.fam_reg_convs:
            f = conv(f)
        anchors = head.anchor_generators[0].grid_anchors(tuple(x.shape[-2:]), s0, device=x.device)
        refine = bbox_decode(head.fam_reg(f), anchors)
        got = head.align_conv(x, refine, s0).cpu().numpy()
    want = oracle.align_conv(x.cpu().numpy(), refine.cpu().numpy(), s0, head.align_conv.deform_conv.weight.detach().cpu().numpy())
    assert np.abs(got - want).max() <= 5e-4          # |activations| ~ 10 with the inflated test weights
    # final NMS inside the head == oracle multiclass NMS on the head's own pre-NMS boxes
    outs = [head.forward_single(x, s) for x, s in zip(feats, strides)]
    from jdet_b200.models.boxes import delta2bbox_rotated, rotated_box_to_poly
    for i in range(2):
        boxes, scores = [], []
        for o in outs:
            sc = o[2][i].permute(1, 2, 0).reshape(-1, 15).sigmoid()
            bp, an = o[3][i].permute(1, 2, 0).reshape(-1, 5), o[1][i].reshape(-1, 5)
            if sc.shape[0] > 200:
                _, idx = sc.max(1)[0].topk(200)
                sc, bp, an = sc[idx], bp[idx], an[idx]
            boxes.append(delta2bbox_rotated(an, bp))
            scores.append(sc)
        boxes, scores = torch.cat(boxes), torch.cat(scores)
        scores = torch.cat([scores.new_zeros((scores.shape[0], 1)), scores], 1)
        wd, wl = oracle.multiclass_nms_rotated(boxes.cpu().numpy(), scores.cpu().numpy(), 0.3, dict(iou_thr=0.1), 100)
        polys, sc, lab = res[i]
        assert polys.shape == (wd.shape[0], 8) and wd.shape[0] > 0
        assert np.array_equal(lab.cpu().numpy(), wl) and np.allclose(sc.cpu().numpy(), wd[:, 5])
        assert np.allclose(polys.cpu().numpy(), rotated_box_to_poly(torch.from_numpy(wd[:, :5])).numpy(), atol=1e-3)


@pytest.mark.gpu
def test_oriented_rcnn_heads_forward():
    """forward-only OrientedRPNHead -> OrientedHead on synthetic FPN maps.  The RoI features the head
    consumes are the rotated RoIAlign kernels' output on the RPN's own proposals: checked against the oracle per level."""
    from jdet_b200.models.roi_heads import OrientedHead, OrientedRPNHead
    torch.manual_seed(5)
    strides = [4, 8, 16, 32, 64]
    feats = [torch.randn(2, 256, 256 // s, 256 // s, device="cuda") for s in strides]
    rpn = OrientedRPNHead(256, nms_pre=500, nms_post=300).cuda().eval()
    for m in (rpn.rpn_cls, rpn.rpn_reg):
        torch.nn.init.normal_(m.weight, 0, 0.05)
    props = rpn(feats)
    assert len(props) == 2 and all(p.shape[1] == 6 and 0 < p.shape[0] <= 300 for p in props)
    assert all(torch.isfinite(p).all() and (p[:, 2] >= p[:, 3]).all() for p in props)        # regular obb: long side first
    assert all((p[1:, 5] <= p[:-1, 5]).all() for p in props)                                   # NMS keeps score order
    head = OrientedHead(num_classes=15, score_thresh=0.01).cuda().eval()
    res = head(feats, props)
    assert len(res) == 2
    for polys, scores, labels in res:
        assert polys.shape[1] == 8 and polys.shape[0] == scores.shape[0] == labels.shape[0] and polys.shape[0] > 0
        assert (scores > 0.01).all() and int(labels.max()) < 15
    # the extractor inside the head == oracle RoIAlign v1 on the level the reference formula picks (oriented_single_level.py:91-114)
    rois = head.arb2roi(props)
    got = head.bbox_roi_extractor(feats[:4], rois).cpu().numpy()
    r = rois.cpu().numpy().copy()
    r[:, 3] *= 1.2; r[:, 4] *= 1.4                                   # extend_factor = (h, w) = (1.4, 1.2)
    lvl = np.clip(np.floor(np.log2(np.sqrt(r[:, 3] * r[:, 4]) / 56 + 1e-6)), 0, 3).astype(int)
    for l in range(4):
        sel = np.nonzero(lvl == l)[0][:40]
        if sel.size:
            want = oracle.roi_align_rotated(feats[l].cpu().numpy(), r[sel], (7, 7), 1.0 / strides[l], 2, 1)
            assert np.abs(got[sel] - want).max() <= TOL


@pytest.mark.gpu
def test_ops_are_cuda_graph_capturable():
    """no host sync / allocation inside the C-ABI calls — the ops can be captured in a CUDA graph and replayed."""
    rng = np.random.default_rng(41)
    x = cu(rng.standard_normal((1, 128, 48, 48)).astype(np.float32))
    rois = cu(_rois(rng, 300, 1, 192.0, 8, 96))
    b1, b2 = cu(dota_boxes(rng, 200, 300.0)), cu(dota_boxes(rng, 150, 300.0))
    boxes = cu(s2anet_anchors(rng, 1, 48, 48, 8)[..., [1, 0, 2, 3, 4]].copy())
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    n = 4000
    d6 = cu(np.concatenate([clustered_boxes(rng, n // 2, 20, 400.0), dota_boxes(rng, n - n // 2, 400.0),
                            ], 0))
    d6 = torch.cat([d6, cu(rng.integers(0, 15, n).astype(np.float32))[:, None]], 1).contiguous()
    sc = cu(tie_free_scores(rng, n))
    lab = d6[:, 5].to(torch.int64)
    anchors = cu(s2anet_anchors(rng, 1, 48, 48, 8))
    ac = AlignConv(128, 64, 3).cuda().requires_grad_(False)

    def run_all():
        order = ops().nms_rotated.argsort_desc(sc)
        return (ops().roi_align_rotated_v1.roi_align(x, rois, (7, 7), 0.25, 2), ops().box_iou_rotated(b1, b2),
                ops().fr.feature_refine(x, boxes, 1 / 8., 1), ops().fr.feature_refine(x, boxes, 1 / 8., 5),
                ops().nms_rotated.nms_rotated_cuda(d6, order, 0.1, 6).to(torch.float32),       # argsort + NMS: keep mask, no host sync
                ops().nms_rotated.ml_nms_rotated_record(d6[:, :5].contiguous(), sc, lab, 0.1, 500),   # + the pack into a send record
                ac(x, anchors, 8))                                                              # fused tcgen05 AlignConv

    eager = run_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            cap = run_all()
    for t in cap:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager, cap):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.1, 0.3, 0.0])
def test_nms_rotated_cpu_convention_on_gpu(thr):
    """the reference CPU path suppresses on IoU >= thr (nms_rotated.py:444).  The GPU mirror
    runs the same kernel with the threshold one ulp lower; checked against the oracle's CPU-convention NMS (std::sort hull
    variant of the IoU, `>=`) on clustered boxes with labels, and through the tile-merge wrapper py_cpu_nms_obb."""
    rng = np.random.default_rng(53)
    n = 3000
    d = clustered_boxes(rng, n, 25, 500.0)
    lab = rng.integers(0, 4, n).astype(np.float32)
    s = tie_free_scores(rng, n)
    d6 = np.concatenate([d, lab[:, None]], 1)
    order = oracle.argsort_desc(s)
    want = oracle.nms_rotated_keep(d6, order, thr, oracle.VARIANT_CPU)
    got = ops().nms_rotated.nms_rotated_cpu(cu(d6), cu(order, torch.int32), thr, 6).cpu().numpy()
    assert np.array_equal(got, want)
    if thr > 0:
        strict = ops().nms_rotated.nms_rotated_cuda(cu(d6), cu(order, torch.int32), thr, 6).cpu().numpy()
        assert strict.sum() >= got.sum()                      # ">" never suppresses more than ">="
        from jdet_b200.models.boxes import obb2poly, rectpoly2obb
        polys = obb2poly(cu(d))
        keep = ops().nms_rotated.py_cpu_nms_obb(torch.cat([polys, cu(s)[:, None]], 1), thr).cpu().numpy()
        boxes5 = rectpoly2obb(polys).cpu().numpy()
        want5 = np.nonzero(oracle.nms_rotated_keep(boxes5, order, thr, oracle.VARIANT_CPU))[0]
        assert np.array_equal(keep, want5)


@pytest.mark.gpu
def test_iou_cpu_build_arithmetic_bit_exact():
    """the reference's CPU build sorts the hull with std::sort (box_iou_rotated.py:316-325) and keeps
    stale distances afterwards (:219-224).  jdet_box_iou_rotated_ex(arithmetic=0) restates that on the GPU: bit-equal to the
    golden vectors produced by the reference's own compiled cpu_src, and to the oracle's CPU variant on random, clustered,
    adversarial and many-vertex (shared corners, 45-degree copies: up to 24 hull points -> the introsort branch) inputs."""
    g = np.load(os.path.join(GOLD, "ref_cpu_iou.npz"))
    got = ops().box_iou_rotated(cu(g["boxes1"]), cu(g["boxes2"]), cpu_arithmetic=True).cpu().numpy()
    assert np.array_equal(bits(got), bits(g["iou_v0_cpu"]))
    rng = np.random.default_rng(61)
    star = []
    for w, h in ((10, 10), (10, 4), (7, 7), (12, 3)):
        for k in range(8):
            star.append([50, 50, w, h, k * np.pi / 4])
            star.append([50 + (k % 3) * 0.5 * w, 50, w, h, k * np.pi / 8])
            star.append([50, 50 + 0.5 * h, w, h, -k * np.pi / 4])
    star = np.asarray(star, np.float32)
    sets = [(dota_boxes(rng, 300, 200.0), dota_boxes(rng, 280, 200.0)), (clustered_boxes(rng, 256, 8, 100.0),) * 2,
            (np.asarray(ADVERSARIAL, np.float32),) * 2, (star, star)]
    ndiff = 0
    for b1, b2 in sets:
        got = ops().box_iou_rotated(cu(b1), cu(b2), cpu_arithmetic=True).cpu().numpy()
        want = oracle.box_iou_rotated(b1, b2, 0, oracle.VARIANT_CPU)
        assert np.array_equal(bits(got), bits(want))
        ndiff += int((want != oracle.box_iou_rotated(b1, b2, 0, oracle.VARIANT_CUDA)).sum())
    assert ndiff > 0          # the two reference builds really do differ on these inputs (else the test proves nothing)


@pytest.mark.gpu
def test_nms_cpu_path_golden():
    """nms_rotated_cpu on the GPU == the reference's compiled CPU NMS (golden keep masks, thr 0.1/0.3/0.5, 5 and 6 columns)."""
    g = np.load(os.path.join(GOLD, "ref_cpu_nms.npz"))
    d6, order = g["dets6"], cu(g["order"], torch.int32)
    for thr in (0.1, 0.3, 0.5):
        for bl, d in ((5, d6[:, :5]), (6, d6)):
            keep = ops().nms_rotated.nms_rotated_cpu(cu(d), order, thr, box_length=bl).cpu().numpy()
            assert np.array_equal(keep, g["keep%d_cpu_thr%02d" % (bl, int(thr * 10))])


# ------------------------------------------------------------------------------------ NMS -> record -> all-gather
@pytest.mark.parametrize("n,max_per_img", [(5000, 2000), (3000, 100), (40, 2000), (1, 5)])
def test_nms_record_matches_eager_pack(n, max_per_img):
    """ml_nms_rotated_record (argsort + NMS + one pack launch into the send buffer) == ml_nms_rotated followed by the eager
    torch packing, bit for bit; and == the oracle's multiclass tail (re-sort by score, truncate)."""
    from jdet_b200 import dist as jdist
    rng = np.random.default_rng(n)
    d = np.concatenate([clustered_boxes(rng, n // 2 + 1, 25, 600.0), dota_boxes(rng, n, 600.0)])[:n]
    s, l = tie_free_scores(rng, n), rng.integers(0, 15, n)
    td, ts, tl = cu(d), cu(s), cu(l, torch.int64)
    buf = torch.full((3, max_per_img + 1, 7), 7.0, device="cuda")              # a dirty persistent buffer: every byte is rewritten
    rec = ops().nms_rotated.ml_nms_rotated_record(td, ts, tl, 0.1, max_per_img, out=buf[1])
    keep = ops().nms_rotated.ml_nms_rotated(td, ts, tl, 0.1)
    want = jdist.pack_detections(td, ts, tl, keep, max_per_img)
    assert rec.data_ptr() == buf[1].data_ptr()
    assert np.array_equal(bits(rec.cpu().numpy()), bits(want.cpu().numpy()))
    assert float(buf[0].min()) == 7.0 and float(buf[2].min()) == 7.0          # neighbours untouched
    ko = oracle.ml_nms_rotated(d, s, l, 0.1)
    order = ko[np.argsort(-s[ko], kind="stable")][:max_per_img]
    got = rec.cpu().numpy()
    assert int(got[-1, 0]) == len(order)
    assert np.array_equal(got[:len(order), :5], d[order]) and np.array_equal(got[:len(order), 6], l[order].astype(np.float32))


def _nccl_worker(rank, world, port, q):
    import os
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    from jdet_b200 import dist as jdist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        outs = []
        for step in range(2):
            imgs = []
            for i in range(2):
                d, s, l = _image_dets(100 * step + 2 * rank + i)
                imgs.append((cu(d).cuda(rank), cu(s).cuda(rank), cu(l, torch.int64).cuda(rank)))
            outs.append(jdist.nms_and_gather(imgs, 0.1, 300).cpu().numpy().copy())
        q.put((rank, outs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _image_dets(seed):
    rng = np.random.default_rng(seed)
    n = 1500 + 37 * (seed % 5)
    d = np.concatenate([clustered_boxes(rng, n // 2, 25, 800.0), dota_boxes(rng, n - n // 2, 800.0)])
    return d, tie_free_scores(rng, n), rng.integers(0, 15, n)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_nccl_gathered_detections_equal_single_gpu():
    """SURVEY 8e parity criterion: the detections gathered over NCCL equal the single-GPU run image by image."""
    import socket
    import torch.multiprocessing as mp
    from jdet_b200 import dist as jdist
    sck = socket.socket(); sck.bind(("127.0.0.1", 0)); port = sck.getsockname()[1]; sck.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for step in range(2):
        assert np.array_equal(got[0][step], got[1][step])
        for r in range(2):
            for i in range(2):
                d, s, l = _image_dets(100 * step + 2 * r + i)
                want = ops().nms_rotated.ml_nms_rotated_record(cu(d), cu(s), cu(l, torch.int64), 0.1, 300).cpu().numpy()
                assert np.array_equal(bits(got[0][step][r, i]), bits(want))
                ko = oracle.ml_nms_rotated(d, s, l, 0.1)
                assert int(want[-1, 0]) == min(len(ko), 300)


# ------------------------------------------------------------------------------------ parity of the torch glue on the GPU
def test_feature_refine_module_numeric():
    """FeatureRefineModule (ops/fr.py:291-347) against oracle.glue: float64 numpy convolutions + the oracle's feature_refine."""
    from oracle import glue
    rng = np.random.default_rng(9)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = ops().fr.FeatureRefineModule(16, [8, 16]).cuda()
        with torch.no_grad():
            for p_ in m.parameters():
                p_.copy_(torch.randn_like(p_) * 0.2)
        xs = [rng.standard_normal((2, 16, 16, 16)).astype(np.float32), rng.standard_normal((2, 16, 8, 8)).astype(np.float32)]
        boxes = [[s2anet_anchors(rng, 1, 16, 16, 8).reshape(-1, 5)[:, [1, 0, 2, 3, 4]].copy(),
                  s2anet_anchors(rng, 1, 8, 8, 16).reshape(-1, 5)[:, [1, 0, 2, 3, 4]].copy()] for _ in range(2)]
        with torch.no_grad():
            out = m([cu(x) for x in xs], [[cu(b) for b in img] for img in boxes])
        g = lambda t: t.detach().cpu().double().numpy()
        want = glue.feature_refine_module(xs, boxes, [8, 16], g(m.conv_5_1.weight), g(m.conv_5_1.bias), g(m.conv_1_5.weight),
                                          g(m.conv_1_5.bias), g(m.conv_1_1.weight), g(m.conv_1_1.bias), oracle.feature_refine)
        for o, w in zip(out, want):
            assert tuple(o.shape) == w.shape
            assert np.abs(o.cpu().numpy() - w).max() <= TOL, np.abs(o.cpu().numpy() - w).max()
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_bbox_decode_and_delta2bbox_on_gpu():
    """bbox_decode / delta2bbox_rotated (s2anet_head.py:631-654, box_ops.py:229-285) run on the GPU against oracle.glue."""
    from oracle import glue
    from jdet_b200.models.boxes import AnchorGeneratorRotatedS2ANet, delta2bbox_rotated
    from jdet_b200.models.roi_heads import bbox_decode
    rng = np.random.default_rng(3)
    gen = AnchorGeneratorRotatedS2ANet(4, [4.0], [1.0], angles=[0.0])
    anchors = gen.grid_anchors((20, 24), 8)
    preds = (rng.standard_normal((3, 5, 20, 24)) * 0.4).astype(np.float32)
    preds[0, 2:4, :3, :3] = 30.0                                            # clipped by wh_ratio_clip = 1e-6 (|dw| <= 13.8)
    got = bbox_decode(cu(preds), anchors.cuda()).cpu().numpy()
    want = glue.bbox_decode(preds, anchors.numpy())
    assert got.shape == want.shape == (3, 20, 24, 5)
    assert np.allclose(got, want, rtol=2e-5, atol=2e-4)
    rois = np.concatenate([rng.uniform(0, 800, (500, 2)), rng.uniform(4, 300, (500, 2)), rng.uniform(-1.6, 1.6, (500, 1))], 1).astype(np.float32)
    deltas = (rng.standard_normal((500, 5)) * 0.5).astype(np.float32)
    stds = (0.1, 0.1, 0.2, 0.2, 0.1)
    got = delta2bbox_rotated(cu(rois), cu(deltas), stds=stds, wh_ratio_clip=16 / 1000).cpu().numpy()
    want = glue.delta2bbox_rotated(rois, deltas, stds=stds, wh_ratio_clip=16 / 1000)
    ang = np.abs(np.mod(got[:, 4] - want[:, 4] + np.pi / 2, np.pi) - np.pi / 2)   # the angle wraps at the range ends
    assert np.allclose(got[:, :4], want[:, :4], rtol=2e-5, atol=2e-4) and ang.max() < 1e-4


def test_orn_on_gpu_against_compiled_reference():
    """ORConv2d's ARF rotation and RIE on the GPU == the reference's compiled CPU sources (golden: tests/golden/ref_cpu_orn.npz)."""
    from jdet_b200.ops.orn import active_rotating_filter, rotation_invariant_encoding
    z = np.load(os.path.join(GOLD, "ref_cpu_orn.npz"))
    for t in range(4):
        got = active_rotating_filter(cu(z["arf%d_w" % t]), torch.as_tensor(z["arf%d_ind" % t]).cuda()).cpu().numpy()
        assert np.array_equal(bits(got), bits(z["arf%d_out" % t]))
    for t in range(3):
        f, nori = z["rie%d_f" % t], int(z["rie%d_nori" % t])
        out, d = rotation_invariant_encoding(cu(f)[:, :, None, None], nori)
        assert np.array_equal(d.cpu().numpy(), z["rie%d_dir" % t])
        assert np.array_equal(bits(out.cpu().numpy()[:, :, 0, 0]), bits(z["rie%d_out" % t]))


def test_nms_scan_handoff_stress():
    """nms_scan_kernel publishes kept(cb) between warps through shared memory (volatile + __threadfence_block; racecheck
    flags it, profiles/r01_sanitizer_reentry.txt).  1000 repetitions over many classes x many column blocks, dense clusters:
    every keep mask must hash identically (and equal the oracle's)."""
    rng = np.random.default_rng(123)
    n = 24000
    d = np.concatenate([clustered_boxes(rng, n * 3 // 4, 60, 900.0), dota_boxes(rng, n - n * 3 // 4, 900.0)])
    s, l = tie_free_scores(rng, n), rng.integers(0, 5, n)                  # 5 classes: ~75 column blocks per class
    d6 = cu(np.concatenate([d, l[:, None].astype(np.float32)], 1))
    order = ops().nms_rotated.argsort_desc(cu(s))
    wts = torch.arange(1, n + 1, device="cuda", dtype=torch.float64)
    first = None
    for it in range(1000):
        keep = ops().nms_rotated.nms_rotated_cuda(d6, order, 0.1, 6)
        h = (keep.to(torch.float64) * wts).sum()
        if first is None:
            first = h.clone()
            want = np.zeros(n, bool)
            want[oracle.ml_nms_rotated(d, s, l, 0.1)] = True
            assert np.array_equal(keep.cpu().numpy(), want)
        else:
            assert bool(h == first), it


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("thr", [0.1, 0.3])
def test_nms_poly_merge_vs_oracle(fast, thr):
    """py_cpu_nms_poly_fast / py_cpu_nms_poly (data/devkits/result_merge.py:69-131, :33-66) on the GPU == the oracle's
    restatement: same kept indices in the same (descending score) order.  Rectangles, jittered convex quadrilaterals,
    clockwise inputs and duplicates."""
    from oracle import glue
    from jdet_b200.models.boxes import rotated_box_to_poly
    rng = np.random.default_rng(17 + int(fast))
    n = 700
    b = np.concatenate([clustered_boxes(rng, n // 2, 12, 500.0), dota_boxes(rng, n - n // 2, 500.0)]).astype(np.float32)
    polys = rotated_box_to_poly(torch.from_numpy(b)).numpy()
    polys[::3] += rng.uniform(-0.15, 0.15, polys[::3].shape).astype(np.float32) * b[::3, 2:3].clip(max=20)   # convex, not rectangles
    polys[1::7] = polys[1::7][:, [6, 7, 4, 5, 2, 3, 0, 1]]                           # clockwise
    polys[5] = polys[4]                                                               # exact duplicate
    dets = np.concatenate([polys, tie_free_scores(rng, n)[:, None]], 1).astype(np.float32)
    fn = ops().nms_rotated.py_cpu_nms_poly_fast if fast else ops().nms_rotated.py_cpu_nms_poly
    got = fn(cu(dets), thr).cpu().numpy()
    want = glue.py_cpu_nms_poly_fast(dets, thr, fast=fast)
    assert np.array_equal(got, want), (len(got), len(want))
    assert fn(cu(dets[:0]), thr).numel() == 0
    one = fn(cu(dets[:1]), thr).cpu().numpy()
    assert np.array_equal(one, [0])


def test_oriented_rcnn_batched_heads_equal_per_image_path():
    """cfg5 tail: the batched RPN (one decode, one horizontal NMS call over (image, level) groups) and the batched RoI head
    (one rotated NMS call for all images with image-major labels, per-image range pack) give, image by image, what the
    per-image paths give."""
    from jdet_b200.models.roi_heads import OrientedHead, OrientedRPNHead
    from jdet_b200.models.boxes import rectpoly2obb
    torch.manual_seed(3)
    rpn = OrientedRPNHead(64, nms_pre=600, nms_post=400).cuda().eval()
    torch.nn.init.normal_(rpn.rpn_cls.weight, 0, 0.05); torch.nn.init.normal_(rpn.rpn_reg.weight, 0, 0.02)
    head = OrientedHead(num_classes=15, in_channels=64, fc_out_channels=128).cuda().eval()
    torch.nn.init.normal_(head.fc_cls.weight, 0, 0.05)
    feats = [torch.randn(3, 64, 512 // s, 512 // s, device="cuda") for s in (4, 8, 16, 32, 64)]
    single = rpn(feats)
    props, counts = rpn.forward_batched(feats)
    for i in range(3):
        k = int(counts[i])
        assert k == single[i].shape[0] and torch.equal(props[i, :k], single[i]) and float(props[i, k:].abs().max() if k < 400 else 0) == 0
    rec = head.detect_records(feats, props, counts, 0.1, 300)
    # the tail on the SAME class scores / regressions, image by image (the FC GEMMs are library calls whose last bits move
    # with the batch size, so they are computed once): identical records
    N, P = props.shape[:2]
    rois = torch.cat([torch.arange(N, device="cuda", dtype=torch.float32)[:, None, None].expand(N, P, 1), props[..., :5]], 2).reshape(N * P, 6)
    cls_score, bbox_pred = head.forward_single(feats, rois)
    again = head.records_from_scores(rois, cls_score, bbox_pred, counts, N, P, 0.1, 300)
    for i in range(3):
        r_i = rois[i * P:(i + 1) * P].clone(); r_i[:, 0] = 0
        one = head.records_from_scores(r_i, cls_score[i * P:(i + 1) * P], bbox_pred[i * P:(i + 1) * P], counts[i:i + 1], 1, P, 0.1, 300)
        assert np.array_equal(bits(again[i].cpu().numpy()), bits(one[0].cpu().numpy()))
        assert int(rec[i, -1, 0]) == int(again[i, -1, 0])
    # and against the reference-shaped path: per-image get_bboxes -> per-class rotated NMS on the decoded boxes
    dets = head(feats, [props[i, :int(counts[i])] for i in range(3)])
    for i, (polys, sc, lab) in enumerate(dets):
        k = int(rec[i, -1, 0])
        if polys.shape[0] == 0:
            assert k == 0
            continue
        keep = ops().nms_rotated.ml_nms_rotated(rectpoly2obb(polys), sc, lab, 0.1)
        assert k == min(int(keep.numel()), 300)
        top = torch.sort(sc[keep], descending=True, stable=True)[0][:k]
        assert torch.allclose(rec[i, :k, 5], top, atol=0, rtol=0)


def test_oriented_rcnn_network_runs_end_to_end():
    """tiles in -> records out on the R50-FPN stand-in (random weights): shapes, counts within bounds, scores sorted."""
    from jdet_b200.models.networks import OrientedRCNN
    torch.manual_seed(0)
    net = OrientedRCNN(max_per_img=500).cuda().eval()
    torch.nn.init.normal_(net.rpn.rpn_cls.weight, 0, 0.05); torch.nn.init.normal_(net.roi_head.fc_cls.weight, 0, 0.05)
    img = torch.randint(0, 256, (2, 3, 512, 512), dtype=torch.uint8, device="cuda")
    rec = net(img)
    assert tuple(rec.shape) == (2, 501, 7)
    for i in range(2):
        k = int(rec[i, -1, 0])
        assert 0 <= k <= 500 and bool((rec[i, :k - 1, 5] >= rec[i, 1:k, 5]).all()) and float(rec[i, k:500].abs().max() if k < 500 else 0) == 0


def test_horizontal_nms_matches_torchvision():
    """The RPN's horizontal proposal NMS runs on the rotated-NMS kernels with theta = 0; same survivors as torchvision.ops.nms
    with the groups kept apart by a coordinate offset (random boxes: no pair within 1e-6 of the threshold)."""
    from torchvision.ops import nms
    from jdet_b200.models.roi_heads.oriented_rpn_head import horizontal_nms
    rng = np.random.default_rng(8)
    n = 6000
    xy = rng.uniform(0, 600, (n, 2)); wh = rng.uniform(8, 120, (n, 2))
    hbb = cu(np.concatenate([xy, xy + wh], 1))
    sc = cu(tie_free_scores(rng, n))
    grp = cu(rng.integers(0, 6, n), torch.int64)
    got = horizontal_nms(hbb, sc, grp, 0.7)
    want = nms(hbb + (grp.float() * 2000)[:, None], sc, 0.7)
    assert torch.equal(got, want)


def test_argsort_desc_signed_zero_ties_and_rescale():
    s = cu(np.array([0.0, -0.0, 0.5, -0.0, 0.0, -1.0], np.float32))
    assert ops().nms_rotated.argsort_desc(s).cpu().numpy().tolist() == [2, 0, 1, 3, 4, 5]


@pytest.mark.parametrize("kind", ["clustered", "lattice", "big"])
def test_nms_pruning_bound_never_changes_the_mask(kind, monkeypatch):
    """ADVICE: the IoU-upper-bound pruning of the mask kernel assumes reference IoU <= bound.  JDET_NMS_NO_PRUNE=1 sends every
    SAT survivor through the exact routine; the keep masks must be identical with and without pruning."""
    rng = np.random.default_rng(5)
    if kind == "lattice":
        n = 3000
        d = _lattice_boxes(rng, n)
    else:
        n = 100000 if kind == "big" else 20000
        d = np.concatenate([clustered_boxes(rng, n // 2, 50), dota_boxes(rng, n - n // 2)])
    s, l = tie_free_scores(rng, n), rng.integers(0, 15 if kind != "lattice" else 2, n)
    d6 = cu(np.concatenate([d, l[:, None].astype(np.float32)], 1))
    order = ops().nms_rotated.argsort_desc(cu(s))
    for thr in (0.1, 0.5):
        monkeypatch.delenv("JDET_NMS_NO_PRUNE", raising=False)
        a = ops().nms_rotated.nms_rotated_cuda(d6, order, thr, 6)
        monkeypatch.setenv("JDET_NMS_NO_PRUNE", "1")
        b = ops().nms_rotated.nms_rotated_cuda(d6, order, thr, 6)
        assert torch.equal(a, b)


def test_candidate_queue_full_path(monkeypatch):
    """ADVICE: the device-wide candidate queues of IoU and NMS (now 64-bit reservation counters) fall back to in-CTA
    evaluation when full.  JDET_TEST_QUEUE_CAP forces that path on a dense cluster; results must not change."""
    rng = np.random.default_rng(2)
    b = clustered_boxes(rng, 1500, 200, 300.0)
    l = rng.integers(0, 3, b.shape[0])
    d6 = cu(np.concatenate([b, l[:, None].astype(np.float32)], 1))
    order = ops().nms_rotated.argsort_desc(cu(tie_free_scores(rng, b.shape[0])))
    monkeypatch.delenv("JDET_TEST_QUEUE_CAP", raising=False)
    iou0 = ops().box_iou_rotated(cu(b), cu(b[:700]))
    keep0 = ops().nms_rotated.nms_rotated_cuda(d6, order, 0.3, 6)
    monkeypatch.setenv("JDET_TEST_QUEUE_CAP", "1000")
    iou1 = ops().box_iou_rotated(cu(b), cu(b[:700]))
    keep1 = ops().nms_rotated.nms_rotated_cuda(d6, order, 0.3, 6)
    assert torch.equal(keep0, keep1)
    assert np.array_equal(bits(iou0.cpu().numpy()), bits(iou1.cpu().numpy()))


def test_align_conv_multi_level_equals_per_level():
    """jdet_align_conv_forward_multi (one persistent tcgen05 launch over all FPN levels) == the per-level calls, bit for bit
    (same tiles, same K order), and <= 1e-4 from the oracle; level-0 sized map included (cfg4: 8 x 256 x 128 x 128 is checked
    in test_full_size_cfg4_level0)."""
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    rng = np.random.default_rng(12)
    levels = [(40, 8), (20, 16), (10, 32), (5, 64), (3, 128)]
    m = AlignConv(64, 96, 3).cuda().requires_grad_(False)
    with torch.no_grad():
        m.deform_conv.weight.copy_(torch.randn_like(m.deform_conv.weight) * 0.05)
    xs = [cu(rng.standard_normal((2, 64, hw, hw + 4))) for hw, _ in levels]
    an = [cu(s2anet_anchors(rng, 2, hw, hw + 4, st)) for hw, st in levels]
    multi = m.forward_multi(xs, an, [st for _, st in levels])
    multi_cl = m.forward_multi([x.contiguous(memory_format=torch.channels_last) for x in xs], an, [st for _, st in levels])
    for o, o2 in zip(multi, multi_cl):                 # channels_last maps are sampled in place: same bits
        assert torch.equal(o, o2)
    for x, a, (hw, st), o in zip(xs, an, levels, multi):
        single = m(x, a, st)
        assert np.array_equal(bits(o.cpu().numpy()), bits(single.cpu().numpy()))
        want = oracle.align_conv(x.cpu().numpy(), a.cpu().numpy(), st, m.deform_conv.weight.cpu().numpy())
        assert np.abs(o.cpu().numpy() - want).max() <= TOL


def test_full_size_cfg4_level0():
    """cfg4 level 0 (8 x 256 x 128 x 128) for feature_refine (points 1 and 5: band / stage arithmetic at the bench shape) and
    AlignConv (1024 tiles on 148 CTAs) against the oracle on sampled positions / a channel slice."""
    from jdet_b200.models.roi_heads.s2anet_head import AlignConv
    rng = np.random.default_rng(21)
    N, C, H, W, st = 8, 256, 128, 128, 8
    x = torch.randn((N, C, H, W), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    a = s2anet_anchors(rng, N, H, W, st)
    boxes = cu(a[..., [1, 0, 2, 3, 4]].copy())
    xn = x.cpu().numpy()
    for points in (1, 5):
        got = ops().fr.feature_refine(x, boxes, 1 / st, points)
        for n, cs in ((0, slice(0, 8)), (7, slice(248, 256)), (3, slice(100, 104))):        # channel slices of three images
            want = oracle.feature_refine(xn[n:n + 1, cs], a[n:n + 1][..., [1, 0, 2, 3, 4]], 1 / st, points)
            assert np.abs(got[n:n + 1, cs].cpu().numpy() - want).max() <= TOL
    m = AlignConv(256, 256, 3).cuda().requires_grad_(False)
    got = m(x, cu(a), st)
    w = m.deform_conv.weight.detach().cpu().numpy()
    for n in (0, 7):                                   # the oracle on a 16-row band of two images (full width, all channels)
        rows = slice(56, 72)
        full = oracle.align_conv(xn[n:n + 1], a[n:n + 1], st, w)
        assert np.abs(got[n:n + 1, :, rows].cpu().numpy() - full[:, :, rows]).max() <= TOL


@pytest.mark.parametrize("points", [1, 5])
def test_feature_refine_staged_bands_vs_oracle(points):
    """cfg4 level-0 geometry (128 x 128: 16 row bands, first / interior / last) on a few channels: the staged kernels' halo, zero
    pad and clamped last row / column, the refined anchors' corner samples leaving the map on every side"""
    rng = np.random.default_rng(40 + points)
    N, C, H, W, stride = 1, 12, 128, 128, 8.0
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    boxes = s2anet_anchors(rng, N, H, W, stride)[..., [1, 0, 2, 3, 4]].copy()
    boxes[0, 64, 10:20, 2:4] *= 5                                # corners beyond the 8 halo rows: the global-tap path
    boxes[0, :, W - 1, 1] = (W - 1) * stride + 3.0               # centres between the last column and the map's edge (clamped taps)
    boxes[0, H - 1, :, 0] = (H - 1) * stride + 5.0
    got = ops().fr.feature_refine(cu(x), cu(boxes), 1 / stride, points).cpu().numpy()
    want = oracle.feature_refine(x, boxes, 1 / stride, points)
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()


@pytest.mark.parametrize("points", [1, 5])
@pytest.mark.parametrize("hw", [(1, 8), (2, 4), (3, 12), (40, 4)])
def test_feature_refine_tiny_maps(points, hw):
    """maps of one or two rows / four columns: every sample is clamped onto the last row or column"""
    H, W = hw
    rng = np.random.default_rng(H * 100 + W + points)
    x = rng.standard_normal((2, 5, H, W)).astype(np.float32)
    boxes = s2anet_anchors(rng, 2, H, W, 8.0)[..., [1, 0, 2, 3, 4]].copy()
    got = ops().fr.feature_refine(cu(x), cu(boxes), 1 / 8.0, points).cpu().numpy()
    want = oracle.feature_refine(x, boxes, 1 / 8.0, points)
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max()


@pytest.mark.parametrize("points", [1, 5])
def test_feature_refine_multi_level_equals_per_level(points):
    """jdet_feature_refine_multi (staged levels share one launch; W % 4 != 0 levels fall back) == per-level feature_refine, bit for bit"""
    rng = np.random.default_rng(30 + points)
    shapes = [(40, 48, 8), (20, 24, 16), (10, 12, 32), (5, 6, 64), (3, 3, 128)]       # the last two: W % 4 != 0 -> per-level kernels
    xs = [cu(rng.standard_normal((2, 48, h, w))) for h, w, _ in shapes]
    bs = [cu(s2anet_anchors(rng, 2, h, w, st)[..., [1, 0, 2, 3, 4]].copy()) for h, w, st in shapes]
    multi = ops().fr.feature_refine_multi(xs, bs, [1.0 / st for _, _, st in shapes], points)
    for x, b, (_, _, st), o in zip(xs, bs, shapes, multi):
        assert torch.equal(o, ops().fr.feature_refine(x, b, 1.0 / st, points))


# ------------------------------------------------------------------------------------ tile -> image merge
@pytest.mark.gpu
def test_result_merge_files_match_the_oracle(tmp_path):
    """mergebypoly / mergebyobb on result files of overlapping tiles == the same merge with the oracle's NMS, line for line"""
    from oracle import glue
    from jdet_b200.data.devkits import result_merge as rm
    from jdet_b200.models.boxes import rotated_box_to_poly
    rng = np.random.default_rng(77)
    src = tmp_path / "raw"; src.mkdir()
    names = ["Task1_plane", "Task1_ship"]
    for name in names:
        rows = []
        for img in ("P0001", "P0007"):
            boxes = clustered_boxes(rng, 240, 6, 700.0)                       # image coordinates; duplicates = objects seen by two tiles
            polys = rotated_box_to_poly(torch.as_tensor(boxes)).numpy().astype(np.float64)
            scores = tie_free_scores(rng, len(boxes))
            for p, s in zip(polys, scores):
                tx, ty = int(rng.integers(0, 2)) * 512, int(rng.integers(0, 2)) * 512
                tile = "%s__1__%d___%d" % (img, tx, ty)
                rows.append((tile, s, [round(float(v - (tx, ty)[i & 1]), 1) for i, v in enumerate(p)]))
        with open(src / (name + ".txt"), "w") as f:
            for tile, s, p in rows:
                f.write(tile + " " + repr(float(s)) + " " + " ".join(repr(v) for v in p) + "\n")
    for merge, tag in ((rm.mergebypoly, "poly"), (rm.mergebyobb, "obb")):
        dst = tmp_path / ("merged_" + tag); dst.mkdir()
        merge(str(src), str(dst))
        for name in names:
            parsed = rm.parse_result_file(str(src / (name + ".txt")))
            want = []
            for img, dets in parsed.items():
                d = np.array(dets)
                if tag == "poly":
                    keep = glue.py_cpu_nms_poly_fast(d.astype(np.float32).astype(np.float64), 0.1)
                else:
                    from jdet_b200.models.boxes.coder import rectpoly2obb
                    b = rectpoly2obb(torch.as_tensor(d[:, :8], dtype=torch.float32)).numpy()
                    order = oracle.argsort_desc(d[:, 8].astype(np.float32))
                    keep = np.nonzero(oracle.nms_rotated_keep(b, order, 0.1, oracle.VARIANT_CPU))[0]
                want += [img + " " + str(dets[int(i)][-1]) + " " + " ".join(map(str, dets[int(i)][:-1])) for i in keep]
            got = open(dst / (name + ".txt")).read().splitlines()
            assert got == want, (tag, name, len(got), len(want))
            assert 0 < len(got) < sum(len(v) for v in parsed.values())


@pytest.mark.gpu
@pytest.mark.parametrize("points", [1, 5])
def test_feature_refine_fuzz_shapes_and_boxes(points):
    """random map shapes (staged and unstaged widths, one-row maps, bands that end mid-stage), boxes from sub-pixel to larger than
    the map, centres far (and absurdly far) outside it: staged kernels == oracle"""
    rng = np.random.default_rng(900 + points)
    for case in range(28):
        N, C = int(rng.integers(1, 4)), int(rng.integers(1, 10))
        H = int(rng.integers(1, 72))
        W = int(rng.integers(1, 40)) * 4 if case % 4 else int(rng.integers(1, 70))
        stride = float(rng.choice([4.0, 8.0, 16.0]))
        x = rng.standard_normal((N, C, H, W)).astype(np.float32)
        boxes = s2anet_anchors(rng, N, H, W, stride)[..., [1, 0, 2, 3, 4]].copy()
        boxes[..., 2:4] *= np.exp(rng.uniform(-3, 3, boxes[..., 2:4].shape)).astype(np.float32)      # 1/20 x ... 20 x the anchor
        far = rng.random(boxes.shape[:3]) < 0.03
        boxes[far, 0:2] += rng.uniform(-3, 3, (int(far.sum()), 2)).astype(np.float32) * stride * max(H, W)
        bad = rng.random(boxes.shape[:3]) < 0.01
        boxes[bad, int(rng.integers(0, 2))] = rng.choice([1e30, -1e30, 3e9])     # (NaN / Inf coordinates index out of bounds in the reference: undefined there)
        got = ops().fr.feature_refine(cu(x), cu(boxes), 1 / stride, points).cpu().numpy()
        want = oracle.feature_refine(x, boxes, 1 / stride, points)
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok), (case, H, W)
        assert np.abs(got[ok] - want[ok]).max(initial=0.0) <= TOL, (case, N, C, H, W, np.abs(got[ok] - want[ok]).max())


@pytest.mark.gpu
def test_align_conv_cta_pair_variant_is_bit_identical():
    """JDET_ALIGN_CONV_2CTA=1 (tcgen05 cta_group::2: M = 256 MMAs over a CTA pair, half the weights per CTA) == the default kernel,
    bit for bit, on a multi-level call with an odd tile count (one dummy tile in the last pair).  The switch is read once per
    process, so the variant runs in a child process."""
    import hashlib
    import subprocess
    import sys
    code = r'''
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
from jdet_b200.models.roi_heads.s2anet_head import AlignConv
from _inputs import s2anet_anchors
dev = torch.device("cuda:0")
rng = np.random.default_rng(4)
levels = [(24, 8), (12, 16), (5, 32)]
g = torch.Generator(device=dev).manual_seed(9)
xs = [torch.randn((3, 64, h, h + 4), device=dev, generator=g) for h, _ in levels]
an = [torch.as_tensor(s2anet_anchors(rng, 3, h, h + 4, s)).to(dev) for h, s in levels]
torch.manual_seed(2)
ac = AlignConv(64, 128, 3).to(dev).requires_grad_(False)
outs = ac.forward_multi(xs, an, [s for _, s in levels])
torch.cuda.synchronize()
h = hashlib.sha1()
for o in outs:
    h.update(o.cpu().numpy().tobytes())
print("HASH", h.hexdigest())
''' % (ROOT, ROOT)
    hashes = []
    for pair in (False, True):
        env = dict(os.environ)
        env.pop("JDET_ALIGN_CONV_2CTA", None)
        if pair:
            env["JDET_ALIGN_CONV_2CTA"] = "1"
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes.append([l for l in out.stdout.splitlines() if l.startswith("HASH")][0])
    assert hashes[0] == hashes[1]


@pytest.mark.gpu
def test_result_merge_py_cpu_nms_horizontal():
    """py_cpu_nms (result_merge.py:147-178: pixel-counting areas, x2 - x1 + 1) on the library's NMS == a numpy restatement"""
    from jdet_b200.data.devkits import result_merge as rm
    rng = np.random.default_rng(5)
    n = 600
    ctr = rng.uniform(0, 400, (n, 2))
    wh = np.exp(rng.uniform(np.log(6), np.log(90), (n, 2)))
    dets = np.concatenate([np.round(ctr - wh / 2), np.round(ctr + wh / 2), tie_free_scores(rng, n)[:, None]], 1).astype(np.float64)

    def restated(d, thresh):                                # the reference's greedy loop, in float64
        x1, y1, x2, y2, sc = d.T
        areas = (x2 - x1 + 1) * (y2 - y1 + 1)
        order = sc.argsort()[::-1]
        keep = []
        while order.size > 0:
            i = order[0]
            keep.append(int(i))
            w = np.maximum(0.0, np.minimum(x2[i], x2[order[1:]]) - np.maximum(x1[i], x1[order[1:]]) + 1)
            h = np.maximum(0.0, np.minimum(y2[i], y2[order[1:]]) - np.maximum(y1[i], y1[order[1:]]) + 1)
            ovr = w * h / (areas[i] + areas[order[1:]] - w * h)
            order = order[np.where(ovr <= thresh)[0] + 1]
        return keep

    for thr in (0.1, 0.3, 0.5):
        got = rm.py_cpu_nms(dets, thr)
        assert list(got) == restated(dets, thr), thr

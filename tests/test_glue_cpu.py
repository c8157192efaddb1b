"""CPU: the torch mirrors of the reference's glue (ORN ARF / RIE, box coders) against oracle.glue and the golden vectors
produced by the reference's own compiled CPU sources (tests/golden/ref_cpu_orn.npz, tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import glue

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_arf_and_rie_match_compiled_reference_golden():
    from jdet_b200.ops.orn import RotationInvariantEncoding, active_rotating_filter, rotation_invariant_encoding
    z = np.load(os.path.join(GOLD, "ref_cpu_orn.npz"))
    for t in range(4):
        w, ind, want = z["arf%d_w" % t], z["arf%d_ind" % t], z["arf%d_out" % t]
        assert np.array_equal(active_rotating_filter(torch.from_numpy(w), torch.from_numpy(ind)).numpy(), want)
        assert np.array_equal(glue.arf_forward(w, ind), want)                      # the restatement is pinned too
        live = glue.ref_arf_forward(w, ind)
        assert live is None or np.array_equal(live, want)
    for t in range(3):
        f, nori = z["rie%d_f" % t], int(z["rie%d_nori" % t])
        out, d = rotation_invariant_encoding(torch.from_numpy(f)[:, :, None, None], nori)
        assert np.array_equal(d.numpy(), z["rie%d_dir" % t]) and np.array_equal(out.numpy()[:, :, 0, 0], z["rie%d_out" % t])
        d2, a2 = glue.rie_forward(f, nori)
        assert np.array_equal(d2, z["rie%d_dir" % t]) and np.array_equal(a2, z["rie%d_out" % t])
        m = RotationInvariantEncoding(nori, return_direction=True)
        o3, d3 = m(torch.from_numpy(f)[:, :, None, None])
        assert torch.equal(o3, out) and torch.equal(d3, d)


def test_rie_gradient_is_the_inverse_roll():
    """rie_backward (ops/orn.py:332-360): gradInput[.., (l + d) mod nOri] = gradOutput[.., l]"""
    from jdet_b200.ops.orn import rotation_invariant_encoding
    g = torch.Generator().manual_seed(0)
    f = torch.randn((4, 24, 1, 1), generator=g, requires_grad=True)
    out, d = rotation_invariant_encoding(f, 8)
    go = torch.randn(out.shape, generator=g)
    out.backward(go)
    want = torch.zeros(4, 3, 8)
    gov, dv = go.reshape(4, 3, 8), d.long()
    for i in range(4):
        for j in range(3):
            for l in range(8):
                want[i, j, (l + int(dv[i, j])) % 8] = gov[i, j, l]
    assert torch.equal(f.grad.reshape(4, 3, 8), want)


def test_box_coders_match_glue_restatement():
    from jdet_b200.models.boxes import AnchorGeneratorRotatedS2ANet, delta2bbox_rotated
    from jdet_b200.models.roi_heads import bbox_decode
    rng = np.random.default_rng(1)
    rois = np.concatenate([rng.uniform(0, 500, (200, 2)), rng.uniform(4, 200, (200, 2)), rng.uniform(-1.6, 1.6, (200, 1))], 1).astype(np.float32)
    deltas = (rng.standard_normal((200, 5)) * 0.5).astype(np.float32)
    got = delta2bbox_rotated(torch.from_numpy(rois), torch.from_numpy(deltas), stds=(.1, .1, .2, .2, .1)).numpy()
    want = glue.delta2bbox_rotated(rois, deltas, stds=(.1, .1, .2, .2, .1))
    ang = np.abs(np.mod(got[:, 4] - want[:, 4] + np.pi / 2, np.pi) - np.pi / 2)
    assert np.allclose(got[:, :4], want[:, :4], rtol=2e-5, atol=2e-4) and ang.max() < 1e-4
    gen = AnchorGeneratorRotatedS2ANet(4, [4.0], [1.0], angles=[0.0])
    anchors = gen.grid_anchors((6, 7), 8)
    preds = (rng.standard_normal((2, 5, 6, 7)) * 0.3).astype(np.float32)
    assert np.allclose(bbox_decode(torch.from_numpy(preds), anchors).numpy(), glue.bbox_decode(preds, anchors.numpy()), rtol=2e-5, atol=2e-4)


def test_iou_poly_restatement_known_answers():
    """oracle.glue.iou_poly (the restated shapely arithmetic, ops/nms_poly.py:247-252): analytic cases, orientation
    independence, and agreement with the pinned box_iou_rotated oracle on rectangles."""
    import oracle
    from jdet_b200.models.boxes import rotated_box_to_poly
    sq = np.array([0, 0, 2, 0, 2, 2, 0, 2], np.float64)
    assert glue.iou_poly(sq, sq) == 1.0
    assert glue.iou_poly(sq, sq + np.tile([1.0, 0.0], 4)) == 2.0 / 6.0              # half overlap: 2 / (4 + 4 - 2)
    assert glue.iou_poly(sq, sq + np.tile([2.0, 0.0], 4)) == 0.0                    # touching edges
    assert glue.iou_poly(sq, sq[[6, 7, 4, 5, 2, 3, 0, 1]] + np.tile([1.0, 0.0], 4)) == 2.0 / 6.0   # clockwise input
    tri_like = np.array([0, 0, 4, 0, 4, 4, 0, 0.001], np.float64)                   # general convex quadrilateral
    assert abs(glue.iou_poly(sq, tri_like) - glue.iou_poly(tri_like, sq)) < 1e-15
    tiny = sq * 1e-3                                                                 # union < 0.01: the max(., 0.01) clamp
    assert glue.iou_poly(tiny, tiny) == (4e-6) / 0.01
    rng = np.random.default_rng(4)
    b = np.concatenate([rng.uniform(0, 60, (40, 2)), rng.uniform(4, 40, (40, 2)), rng.uniform(-1.5, 1.5, (40, 1))], 1).astype(np.float32)
    polys = rotated_box_to_poly(torch.from_numpy(b)).numpy()
    want = oracle.box_iou_rotated(b, b, 0, oracle.VARIANT_CUDA)
    got = np.array([[glue.iou_poly(polys[i], polys[j]) for j in range(40)] for i in range(40)])
    assert np.abs(got - want).max() < 2e-5                                           # float32 corners vs the fp32 rectangle routine


def test_s2anet_head_state_dict_keys_match_reference_names():
    """A reference S2ANetHead checkpoint names its ConvModule parameters `<list>.<i>.conv.{weight,bias}` and carries the unused
    `or_pool.conv.*` (Conv2d + BatchNorm2d); the mirror must use the same keys so such a checkpoint loads as it is."""
    from jdet_b200.models.roi_heads.s2anet_head import S2ANetHead
    head = S2ANetHead(16, 256)
    keys = set(head.state_dict().keys())
    want = set()
    for lst in ("fam_reg_convs", "odm_reg_convs", "odm_cls_convs"):
        for i in range(2):
            want |= {"%s.%d.conv.weight" % (lst, i), "%s.%d.conv.bias" % (lst, i)}
    want |= {"fam_reg.weight", "fam_reg.bias", "align_conv.deform_conv.weight", "or_conv.weight", "or_conv.bias",
             "odm_cls.weight", "odm_cls.bias", "odm_reg.weight", "odm_reg.bias",
             "or_pool.conv.0.weight", "or_pool.conv.0.bias", "or_pool.conv.1.weight", "or_pool.conv.1.bias",
             "or_pool.conv.1.running_mean", "or_pool.conv.1.running_var"}
    assert want <= keys, sorted(want - keys)
    extra = {k for k in keys - want if not k.endswith("num_batches_tracked") and k != "or_conv.indices"}
    assert not extra, sorted(extra)
    assert tuple(head.state_dict()["or_conv.weight"].shape) == (32, 256, 1, 3, 3)


def test_argsort_key_treats_signed_zero_as_one_score():
    """documented tie rule of jdet_argsort_desc (lower index first) — the host-side statement of it, for the CPU suite"""
    s = np.array([0.0, -0.0, 0.5, -0.0, 0.0], np.float32)
    order = np.argsort(-np.where(s == 0, 0.0, s), kind="stable")
    assert order.tolist() == [2, 0, 1, 3, 4]


# ------------------------------------------------------------------------------------ tile -> image merge (result_merge.py)
def _write_result_file(path, rows):
    with open(path, "w") as f:
        for name, score, poly in rows:
            f.write(name + " " + repr(float(score)) + " " + " ".join(repr(float(v)) for v in poly) + "\n")


def test_result_merge_parsing_grouping_and_output_format(tmp_path):
    """jdet.data.devkits.result_merge mirror: tile names -> (image, scale, offset), poly2origpoly, grouping, the output lines —
    with a stand-in NMS (no GPU here)."""
    from jdet_b200.data.devkits import dota_utils as util
    from jdet_b200.data.devkits import result_merge as rm
    assert rm.poly2origpoly([1, 2, 3, 4, 5, 6, 7, 8], 100, 200, "0.5") == [202.0, 404.0, 206.0, 408.0, 210.0, 412.0, 214.0, 416.0]
    src, dst = tmp_path / "raw", tmp_path / "merged"
    src.mkdir(); dst.mkdir()
    rows = [("P0001__1__0___0", 0.9, [10, 10, 20, 10, 20, 20, 10, 20]),
            ("P0001__1__512___1024", 0.8, [1, 2, 3, 2, 3, 4, 1, 4]),
            ("P0002__0.5__100___0", 0.7, [0, 0, 8, 0, 8, 8, 0, 8]),
            ("P0001__1__0___0", 0.6, [11, 10, 21, 10, 21, 20, 11, 20])]
    _write_result_file(src / "Task1_plane.txt", rows)
    assert util.custombasename(str(src / "Task1_plane.txt")) == "Task1_plane"
    assert util.GetFileFromThisRootDir(str(src)) == [str(src / "Task1_plane.txt")]
    parsed = rm.parse_result_file(str(src / "Task1_plane.txt"))
    assert list(parsed) == ["P0001", "P0002"] and len(parsed["P0001"]) == 3
    assert parsed["P0001"][1] == [513.0, 1026.0, 515.0, 1026.0, 515.0, 1028.0, 513.0, 1028.0, 0.8]
    assert parsed["P0002"][0] == [200.0, 0.0, 216.0, 0.0, 216.0, 16.0, 200.0, 16.0, 0.7]
    seen = []

    def keep_all_reversed(dets, thresh):
        seen.append((dets.shape, thresh))
        return list(range(len(dets)))[::-1]

    rm.mergesingle(str(dst), keep_all_reversed, str(src / "Task1_plane.txt"))
    assert seen == [((3, 9), rm.nms_threshold_0), ((1, 9), rm.nms_threshold_0)]
    lines = open(dst / "Task1_plane.txt").read().splitlines()
    assert lines[0] == "P0001 0.6 11.0 10.0 21.0 10.0 21.0 20.0 11.0 20.0"
    assert lines[3] == "P0002 0.7 200.0 0.0 216.0 0.0 216.0 16.0 200.0 16.0" and len(lines) == 4
    # per-class thresholds (the reference's cfg.merge_nms_threshold_type == 1)
    _write_result_file(src / "plane.txt", rows[:1])
    seen.clear()
    rm.mergesingle(str(dst), keep_all_reversed, str(src / "plane.txt"), nms_threshold_type=1)
    assert seen == [((1, 9), rm.nms_threshold_1["plane"])]
    import jdet_b200
    jdet_b200.install_as_jdet()
    from jdet.data.devkits.result_merge import mergebypoly, mergebyobb, mergebyrec   # noqa: F401  (reference import paths)

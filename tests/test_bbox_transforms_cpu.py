"""CPU-only: jdet_b200.ops.bbox_transforms against plain numpy restatements of the reference formulas
(python/jdet/ops/bbox_transforms.py:499-704) and round-trip properties."""
import math

import numpy as np
import pytest
import torch

from jdet_b200.ops import bbox_transforms as bt
from _inputs import dota_boxes


def _np_obb2poly(o):
    x, y, w, h, t = [o[..., i] for i in range(5)]
    c, s = np.cos(t), np.sin(t)
    v1 = np.stack([w / 2 * c, -w / 2 * s], -1)
    v2 = np.stack([-h / 2 * s, -h / 2 * c], -1)
    ctr = np.stack([x, y], -1)
    return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2], -1)


def test_types_and_dims():
    assert bt.get_bbox_type(torch.zeros(3, 4)) == "hbb" and bt.get_bbox_type(torch.zeros(3, 6), with_score=True) == "obb"
    assert bt.get_bbox_type(torch.zeros(2, 8)) == "poly" and bt.get_bbox_type(torch.zeros(2, 7)) == "notype"
    assert bt.get_bbox_dim("poly", with_score=True) == 9
    with pytest.raises(ValueError):
        bt.get_bbox_dim("circle")
    with pytest.raises(ValueError):
        bt.bbox2type(torch.zeros(2, 7), "obb")


def test_obb_poly_hbb_conversions_match_numpy():
    rng = np.random.default_rng(0)
    o = dota_boxes(rng, 500, 800.0).astype(np.float64)
    t = torch.from_numpy(o)
    poly = bt.obb2poly(t).numpy()
    assert np.allclose(poly, _np_obb2poly(o), atol=1e-9)
    hbb = bt.obb2hbb(t).numpy()
    pts = poly.reshape(-1, 4, 2)
    assert np.allclose(hbb, np.concatenate([pts.min(1), pts.max(1)], -1), atol=1e-9)   # hbb of an obb == hbb of its corners
    assert np.allclose(bt.poly2hbb(torch.from_numpy(poly)).numpy(), hbb, atol=1e-9)
    assert np.allclose(bt.bbox2type(t, "poly").numpy(), poly) and bt.bbox2type(t, "obb") is t
    # areas agree across the three representations of the same box
    assert np.allclose(bt.get_bbox_areas(t).numpy(), o[:, 2] * o[:, 3])
    assert np.allclose(bt.get_bbox_areas(torch.from_numpy(poly)).numpy(), o[:, 2] * o[:, 3], rtol=1e-9)
    assert np.allclose(bt.get_bbox_areas(torch.from_numpy(hbb)).numpy(), (hbb[:, 2] - hbb[:, 0]) * (hbb[:, 3] - hbb[:, 1]))


def test_rectpoly2obb_inverts_obb2poly_up_to_regularisation():
    rng = np.random.default_rng(1)
    o = torch.from_numpy(dota_boxes(rng, 400, 800.0).astype(np.float64))
    back = bt.rectpoly2obb(bt.obb2poly(o))
    want = bt.regular_obb(o)
    assert torch.allclose(back[:, :4], want[:, :4], atol=1e-8)
    d = torch.remainder(back[:, 4] - want[:, 4] + math.pi / 2, math.pi) - math.pi / 2      # angles equal modulo pi
    assert d.abs().max() < 1e-8
    assert bool((back[:, 2] >= back[:, 3]).all()) and bool((back[:, 4] >= -math.pi / 2).all()) and bool((back[:, 4] < math.pi / 2).all())
    with pytest.raises(NotImplementedError):
        bt.poly2obb(bt.obb2poly(o))


def test_hbb_conversions():
    h = torch.tensor([[0., 0., 4., 2.], [1., 1., 2., 5.], [3., 3., 5., 5.]], dtype=torch.float64)
    o = bt.hbb2obb(h)
    assert torch.allclose(o[0], torch.tensor([2., 1., 4., 2., 0.], dtype=torch.float64))
    assert torch.allclose(o[1], torch.tensor([1.5, 3., 4., 1., -math.pi / 2], dtype=torch.float64))      # taller than wide
    assert torch.allclose(o[2], torch.tensor([4., 4., 2., 2., 0.], dtype=torch.float64))                   # square: w >= h branch
    assert torch.allclose(bt.obb2hbb(o), h, atol=1e-12)
    p = bt.hbb2poly(h)
    assert p.shape == (3, 8) and torch.allclose(bt.poly2hbb(p), h)
    assert torch.allclose(bt.regular_theta(torch.tensor([math.pi, -math.pi, 0.3])), torch.tensor([0., 0., 0.3]), atol=1e-6)

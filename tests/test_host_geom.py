"""CPU-only: the PRODUCT's pair geometry (jdet_b200/csrc/rbox_geom.cuh: circle / SAT rejects, the straight-line
exact-IoU routine and its reference-control-flow fallback) compiled for the host and run against the oracle, bit for
bit.  This is what lets the device routine be restructured without a GPU in the loop; the GPU parity tests remain the
proof for the device build itself."""
import ctypes
import os
import sys

import numpy as np
import pytest

import oracle
from _inputs import ADVERSARIAL, clustered_boxes, dota_boxes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host"))
import host_geom_build as host_build  # noqa: E402

fp = ctypes.POINTER(ctypes.c_float)
ip = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def hg():
    so = host_build.build()
    if so is None:
        pytest.skip("nvcc not on PATH")
    return ctypes.CDLL(so)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run(hg, b1, b2, version, variant, mode):
    b1, b2 = np.ascontiguousarray(b1, np.float32), np.ascontiguousarray(b2, np.float32)
    out = np.full((len(b1), len(b2)), -1, np.float32)
    counts = np.zeros(3, np.int32)
    hg.host_geom_iou(b1.ctypes.data_as(fp), len(b1), b2.ctypes.data_as(fp), len(b2), out.ctypes.data_as(fp), version,
                     variant, mode, counts.ctypes.data_as(ip))
    return out, counts


def degenerate_boxes(rng, n):
    """Boxes on a coarse integer lattice with angles that are multiples of 45 degrees: shared corners, collinear and
    parallel edges, exact zeros in the edge tests — the cases that leave the straight-line routine's ordinary range."""
    return np.stack([rng.integers(0, 6, n), rng.integers(0, 6, n), rng.integers(1, 5, n), rng.integers(1, 5, n),
                     rng.integers(-4, 5, n) * (np.pi / 4)], 1).astype(np.float32)


def box_sets():
    rng = np.random.default_rng(11)
    yield "dota", dota_boxes(rng, 500, 400.0), dota_boxes(rng, 400, 400.0)
    c = clustered_boxes(rng, 600, 30, 300.0)
    yield "clustered", c[:300], c[300:]
    yield "self", c[:200], c[:200]
    yield "adversarial", ADVERSARIAL, np.concatenate([ADVERSARIAL, dota_boxes(rng, 30, 10.0, 0.5, 8.0)])
    d = degenerate_boxes(rng, 260)
    yield "lattice", d[:130], d
    huge = dota_boxes(rng, 60, 1e9, 1e8, 1e9)          # products near 1e18: outside the comparison-only range
    tiny = dota_boxes(rng, 60, 1e-5, 1e-7, 1e-5)
    yield "huge", huge, huge[::-1].copy()
    yield "tiny", tiny, tiny[::-1].copy()
    nan = dota_boxes(rng, 8, 50.0)
    nan[1, 0] = np.nan
    nan[3, 4] = np.inf
    nan[5, 2] = np.inf
    yield "nonfinite", nan, dota_boxes(rng, 8, 50.0)


@pytest.mark.parametrize("version", [0, 1])
@pytest.mark.parametrize("variant", [oracle.VARIANT_CUDA, oracle.VARIANT_CPU])
def test_host_build_of_product_geometry_matches_oracle(hg, version, variant):
    exact_pairs = 0
    for name, b1, b2 in box_sets():
        want = oracle.box_iou_rotated(b1, b2, version, variant)
        for mode in (0, 1, 2) + ((3, 4) if variant == oracle.VARIANT_CUDA else ()):
            got, counts = run(hg, b1, b2, version, variant, mode)
            same = (bits(got) == bits(want)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (name, version, variant, mode, np.argwhere(~same)[:5])
            if mode == 2:
                exact_pairs += int(counts[2])
    assert exact_pairs > 20000      # the exact routine really ran (the reject stages did not swallow the test)


def test_upper_bound_never_below_exact(hg):
    """nms_rotated.cu prunes a pair when iou_upper_bound < thr; the bound must dominate the REFERENCE's IoU — which on
    lattice boxes at multiples of 45 degrees (collinear overlapping edges, duplicate hull points) exceeds the true
    IoU by percents; such parallel pairs are exempt from pruning (bound = +inf)."""
    rng = np.random.default_rng(5)
    c = clustered_boxes(rng, 500, 25, 200.0)
    horiz = clustered_boxes(rng, 150, 25, 100.0)
    horiz[:, 4] = 0.0
    horiz[:, :4] = np.round(horiz[:, :4] * 2) / 2
    sets = [np.concatenate([c, degenerate_boxes(rng, 100), dota_boxes(rng, 100, 100.0)]),
            degenerate_boxes(rng, 700), degenerate_boxes(rng, 700), horiz]
    overlapping = exempt = 0
    for b in sets:
        b = np.ascontiguousarray(b, np.float32)
        ub = np.zeros((len(b), len(b)), np.float32)
        hg.host_geom_upper_bound(b.ctypes.data_as(fp), len(b), b.ctypes.data_as(fp), len(b), ub.ctypes.data_as(fp))
        iou = oracle.box_iou_rotated(b, b, 0, oracle.VARIANT_CUDA)
        ok = ub >= iou
        assert ok.all(), [(b[i], b[j], ub[i, j], iou[i, j]) for i, j in np.argwhere(~ok)[:3]]
        overlapping += int((iou > 0).sum())
        exempt += int(np.isinf(ub).sum())
    assert overlapping > 300000 and exempt > 1000

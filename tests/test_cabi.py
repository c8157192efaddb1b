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

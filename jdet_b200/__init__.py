"""jdet_b200 — B200-native (sm_100a) kernels for JDet's oriented-box geometry hot path.

Drop-in for the reference's ``jdet.ops`` call signatures on ``torch.Tensor`` (CUDA, fp32):
box_iou_rotated[_v1], nms_rotated / ml_nms_rotated / multiclass_nms_rotated, ROIAlignRotated[_v1],
feature_refine, DeformConv v1 and S2ANet's AlignConv.  Host code is Python over the C ABI in
``include/jdet_b200.h`` (libjdet_b200.so, hand-written CUDA).  No CPU fallback.
"""
__version__ = "0.1.0"


def install_as_jdet():
    """Alias this package as ``jdet`` in sys.modules so reference-style imports
    (``from jdet.ops import box_iou_rotated``; ``from jdet.ops import roi_align_rotated_v1``;
    ``from jdet.models.roi_heads.s2anet_head import AlignConv``) resolve to the B200 kernels."""
    import importlib
    import sys
    pkg = sys.modules[__name__]
    sys.modules.setdefault("jdet", pkg)
    for sub in ("ops", "ops.box_iou_rotated", "ops.box_iou_rotated_v1", "ops.nms_rotated", "ops.roi_align_rotated",
                "ops.roi_align_rotated_v1", "ops.fr", "ops.dcn_v1", "ops.orn", "ops.bbox_transforms", "models", "models.roi_heads",
                "models.roi_heads.s2anet_head", "models.roi_extractors", "models.boxes",
                "models.boxes.iou_calculator", "data", "data.devkits", "data.devkits.result_merge", "data.devkits.dota_utils"):
        try:
            sys.modules.setdefault("jdet." + sub, importlib.import_module(__name__ + "." + sub))
        except ImportError:
            pass
    return pkg

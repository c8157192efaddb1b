"""RboxSingleRoIExtractor (reference: python/jdet/models/roi_extractors/rbox_single_level.py:8-96):
RoITransformer's extractor over ROIAlignRotated (v0).  As in the reference, w_enlarge / h_enlarge are stored
but not applied when pooling."""
from ...ops import roi_align_rotated
from ._rotated_base import RotatedSingleLevelBase


class RboxSingleRoIExtractor(RotatedSingleLevelBase):
    ops_module = roi_align_rotated

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, w_enlarge=1.2, h_enlarge=1.4):
        super().__init__(roi_layer, out_channels, featmap_strides, finest_scale)
        self.w_enlarge, self.h_enlarge = w_enlarge, h_enlarge

    def forward(self, feats, rois):
        if len(feats) == 1:
            return self.roi_layers[0](feats[0], rois)
        fused = self._fused(0, feats, rois, (1., 1.), (1., 1.))
        if fused is not None:
            return fused
        return self._pool_by_level(feats, rois, self.map_roi_levels(rois, len(feats)))

    execute = forward

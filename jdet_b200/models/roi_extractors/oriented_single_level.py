"""OrientedSingleRoIExtractor (reference: python/jdet/models/roi_extractors/oriented_single_level.py:8-114):
Oriented R-CNN's extractor — RoIs are stretched by `extend_factor` before the level is chosen and before
pooling with ROIAlignRotated_v1."""
from ...ops import roi_align_rotated_v1
from ._rotated_base import RotatedSingleLevelBase, pair


class OrientedSingleRoIExtractor(RotatedSingleLevelBase):
    ops_module = roi_align_rotated_v1

    def __init__(self, roi_layer, out_channels, featmap_strides, extend_factor=(1., 1.), finest_scale=56):
        super().__init__(roi_layer, out_channels, featmap_strides, finest_scale)
        self.extend_factor = extend_factor

    def roi_rescale(self, rois, scale_factor):
        """(h_factor, w_factor) multiply columns 4 and 3 of (n,6) RoIs; None leaves them untouched."""
        if scale_factor is None:
            return rois
        fh, fw = pair(scale_factor)
        scaled = rois.clone()
        scaled[:, 3] *= fw
        scaled[:, 4] *= fh
        return scaled

    def forward(self, feats, rois, roi_scale_factor=None):
        if len(feats) == 1:
            return self.roi_layers[0](feats[0], rois)
        fh, fw = pair(self.extend_factor)
        sh, sw = (1., 1.) if roi_scale_factor is None else pair(roi_scale_factor)
        fused = self._fused(1, feats, rois, (fw, fh), (sw, sh))        # one call: level choice, re-layouts, one gather
        if fused is not None:
            return fused
        stretched = self.roi_rescale(rois, self.extend_factor)
        lvls = self.map_roi_levels(stretched, len(feats))
        return self._pool_by_level(feats, self.roi_rescale(stretched, roi_scale_factor), lvls)

    execute = forward

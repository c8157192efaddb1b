"""RboxSingleRoIExtractor mirror (reference: python/jdet/models/roi_extractors/rbox_single_level.py:8-96).

RoITransformer's extractor: level by sqrt(w*h), ROIAlignRotated (v0) per level.  As in the
reference, w_enlarge / h_enlarge are stored but not applied in execute().
"""
import torch
from torch import nn

from ...ops import roi_align_rotated


class RboxSingleRoIExtractor(nn.Module):
    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, w_enlarge=1.2, h_enlarge=1.4):
        super().__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale
        self.w_enlarge = w_enlarge
        self.h_enlarge = h_enlarge

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert hasattr(roi_align_rotated, layer_type)
        layer_cls = getattr(roi_align_rotated, layer_type)
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def map_roi_levels(self, rois, num_levels):
        scale = torch.sqrt(rois[:, 3] * rois[:, 4])
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def forward(self, feats, rois):
        if len(feats) == 1:
            return self.roi_layers[0](feats[0], rois)
        out_size = self.roi_layers[0].output_size
        num_levels = len(feats)
        target_lvls = self.map_roi_levels(rois, num_levels)
        roi_feats = torch.zeros((rois.shape[0], self.out_channels, out_size[0], out_size[1]), dtype=torch.float32,
                                device=rois.device)
        for i in range(num_levels):
            inds = target_lvls == i
            if inds.any():
                roi_feats[inds] += self.roi_layers[i](feats[i], rois[inds, :])
        return roi_feats

    execute = forward

"""OrientedSingleRoIExtractor mirror (reference: python/jdet/models/roi_extractors/oriented_single_level.py:8-114).

Maps each RoI to an FPN level by sqrt(w*h) after extending (w,h), runs ROIAlignRotated_v1 per level
and scatters the results back into RoI order.
"""
import torch
from torch import nn

from ...ops import roi_align_rotated_v1


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class OrientedSingleRoIExtractor(nn.Module):
    def __init__(self, roi_layer, out_channels, featmap_strides, extend_factor=(1., 1.), finest_scale=56):
        super().__init__()
        self.roi_layers = self.build_roi_layers(roi_layer, featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.extend_factor = extend_factor
        self.finest_scale = finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):
        cfg = layer_cfg.copy()
        layer_type = cfg.pop('type')
        assert hasattr(roi_align_rotated_v1, layer_type)
        layer_cls = getattr(roi_align_rotated_v1, layer_type)
        return nn.ModuleList([layer_cls(spatial_scale=1 / s, **cfg) for s in featmap_strides])

    def map_roi_levels(self, rois, num_levels):
        scale = torch.sqrt(rois[:, 3] * rois[:, 4])
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def roi_rescale(self, rois, scale_factor):
        if scale_factor is None:
            return rois
        h_scale_factor, w_scale_factor = _pair(scale_factor)
        new_rois = rois.clone()
        new_rois[:, 3] = w_scale_factor * new_rois[:, 3]
        new_rois[:, 4] = h_scale_factor * new_rois[:, 4]
        return new_rois

    def forward(self, feats, rois, roi_scale_factor=None):
        if len(feats) == 1:
            return self.roi_layers[0](feats[0], rois)
        out_size = self.roi_layers[0].output_size[0]
        num_levels = len(feats)
        roi_feats = torch.zeros((rois.shape[0], self.out_channels, out_size, out_size), dtype=torch.float32,
                                device=rois.device)
        rois = self.roi_rescale(rois, self.extend_factor)
        target_lvls = self.map_roi_levels(rois, num_levels)
        rois = self.roi_rescale(rois, roi_scale_factor)
        for i in range(num_levels):
            inds = target_lvls == i
            if inds.any():
                roi_feats[inds] += self.roi_layers[i](feats[i], rois[inds, :])
        return roi_feats

    execute = forward

"""Shared machinery of the two rotated single-level RoI extractors.

Reference behaviour (python/jdet/models/roi_extractors/{oriented,rbox}_single_level.py): every RoI is
assigned to one FPN level by floor(log2(sqrt(w*h)/finest_scale + 1e-6)) clamped to the level range, the
level's rotated RoIAlign is run on that subset, and results land at the RoI's original row.

Here the subsets are formed by ONE stable sort of the level ids (contiguous slices per level, no boolean
masks), and each level's result is written with a single index_copy_ instead of zeros + masked "+=".
"""
import torch
from torch import nn


def pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class RotatedSingleLevelBase(nn.Module):
    ops_module = None            # jdet_b200.ops.roi_align_rotated[_v1], set by the subclass

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale):
        super().__init__()
        spec = dict(roi_layer)
        kind = spec.pop('type')
        assert hasattr(self.ops_module, kind)
        maker = getattr(self.ops_module, kind)
        self.roi_layers = nn.ModuleList(maker(spatial_scale=1 / s, **spec) for s in featmap_strides)
        self.out_channels = out_channels
        self.featmap_strides = featmap_strides
        self.finest_scale = finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def build_roi_layers(self, layer_cfg, featmap_strides):     # kept for API compatibility
        spec = dict(layer_cfg)
        maker = getattr(self.ops_module, spec.pop('type'))
        return nn.ModuleList(maker(spatial_scale=1 / s, **spec) for s in featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        """scale < finest: 0; [finest, 2 finest): 1; ... ; clamp to num_levels - 1.  rois: (k,6)."""
        side = (rois[:, 3] * rois[:, 4]).sqrt()
        lvl = torch.log2(side / self.finest_scale + 1e-6).floor()
        return lvl.clamp(0, num_levels - 1).long()

    def _pool_by_level(self, feats, rois, lvls):
        out_hw = pair(self.roi_layers[0].output_size)
        pooled = torch.zeros((rois.shape[0], self.out_channels) + tuple(out_hw), dtype=torch.float32, device=rois.device)
        if rois.shape[0] == 0:
            return pooled
        order = torch.argsort(lvls, stable=True)
        counts = torch.bincount(lvls, minlength=len(feats)).tolist()      # one host sync, like the reference's .any_()
        start = 0
        for level, n in enumerate(counts):
            if n:
                idx = order[start:start + n]
                pooled.index_copy_(0, idx, self.roi_layers[level](feats[level], rois.index_select(0, idx)))
            start += n
        return pooled

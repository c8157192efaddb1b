"""jdet.ops.roi_align_rotated mirror (reference: python/jdet/ops/roi_align_rotated.py:257-322).

v0 convention (RboxSingleRoIExtractor / RoITransformer): no -0.5, x = xx*cos - yy*sin,
count = grid_h*grid_w.  (The reference module also re-exports RiRoIAlign, which is out of scope.)
"""
from torch import nn

from .roi_align_rotated_v1 import _pair, _roi_align_dispatch


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio):
    return _roi_align_dispatch(0, input, rois, output_size, spatial_scale, sampling_ratio)


class ROIAlignRotated(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio=0):
        super().__init__()
        self.output_size = _pair(output_size)
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    execute = forward

    def __repr__(self):
        return (self.__class__.__name__ + "(output_size=" + str(self.output_size) + ", spatial_scale=" +
                str(self.spatial_scale) + ", sampling_ratio=" + str(self.sampling_ratio) + ")")

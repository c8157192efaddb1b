"""jdet.ops.fr mirror — feature_refine / FR / FeatureRefineModule (reference: python/jdet/ops/fr.py:255-347).

feature_refine is R3Det's rotated feature alignment: out = in + sum_i bilinear(in, point_i), the points
being the box centre (points=1) or centre + 4 corners (points=5).  Forward and backward (grad w.r.t. features).
"""
import torch
from torch import nn

from ._common import check, f32c, lib, require_cuda, stream_ptr


def _feature_refine_fwd(features, best_rbboxes, spatial_scale, points):
    assert points in [1, 5]                                    # fr.py:261
    require_cuda(features, best_rbboxes)
    x, b = f32c(features), f32c(best_rbboxes)
    N, C, H, W = x.shape
    assert b.numel() == N * H * W * 5, "best_rbboxes must be (N,H,W,5) (or (N*H*W,5))"
    out = torch.empty_like(x)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        check(lib().jdet_feature_refine(x.data_ptr(), b.data_ptr(), N, C, H, W, points, float(spatial_scale),
                                        out.data_ptr(), stream_ptr(x.device)), "feature_refine")
    return out


def _feature_refine_bwd(grad_output, best_rbboxes, spatial_scale, points):
    g, b = f32c(grad_output), f32c(best_rbboxes)
    N, C, H, W = g.shape
    gi = torch.empty_like(g)
    if gi.numel() == 0:
        return gi
    with torch.cuda.device(g.device):
        check(lib().jdet_feature_refine_backward(g.data_ptr(), b.data_ptr(), N, C, H, W, points, float(spatial_scale),
                                                 gi.data_ptr(), stream_ptr(g.device)), "feature_refine_backward")
    return gi


class FeatureRefineFunction(torch.autograd.Function):
    """fr.py:255-271 (execute/grad -> forward/backward); boxes get no gradient, as in the reference."""

    @staticmethod
    def forward(ctx, features, best_rbboxes, spatial_scale, points=1):
        ctx.spatial_scale, ctx.points = spatial_scale, points
        ctx.save_for_backward(best_rbboxes)
        return _feature_refine_fwd(features, best_rbboxes, spatial_scale, points)

    @staticmethod
    def backward(ctx, grad_output):
        (boxes,) = ctx.saved_tensors
        return _feature_refine_bwd(grad_output, boxes, ctx.spatial_scale, ctx.points), None, None, None


def feature_refine(features, best_rbboxes, spatial_scale, points=1):
    if torch.is_grad_enabled() and isinstance(features, torch.Tensor) and features.requires_grad:
        return FeatureRefineFunction.apply(features, best_rbboxes, spatial_scale, points)
    return _feature_refine_fwd(features, best_rbboxes, spatial_scale, points)


def feature_refine_multi(features_list, best_rbboxes_list, spatial_scales, points=1):
    """feature_refine on every FPN level in ONE library call (inference; no autograd): the levels share a launch
    (jdet_feature_refine_multi).  Level by level the same values as feature_refine."""
    import ctypes
    assert points in [1, 5]
    n = len(features_list)
    require_cuda(*features_list, *best_rbboxes_list)
    xs = [f32c(x) for x in features_list]
    bs = [f32c(b) for b in best_rbboxes_list]
    N, C = xs[0].shape[:2]
    for x, b in zip(xs, bs):
        assert x.shape[0] == N and x.shape[1] == C and b.numel() == N * x.shape[2] * x.shape[3] * 5
    outs = [torch.empty_like(x) for x in xs]
    if n > 8 or n == 0:
        return [feature_refine(x, b, s, points) for x, b, s in zip(xs, bs, spatial_scales)]
    arr = lambda vals, ct: (ct * n)(*vals)
    with torch.cuda.device(xs[0].device):
        check(lib().jdet_feature_refine_multi(arr([x.data_ptr() for x in xs], ctypes.c_void_p), arr([b.data_ptr() for b in bs], ctypes.c_void_p),
                                              n, N, C, arr([x.shape[2] for x in xs], ctypes.c_int), arr([x.shape[3] for x in xs], ctypes.c_int),
                                              arr([float(s) for s in spatial_scales], ctypes.c_float), points,
                                              arr([o.data_ptr() for o in outs], ctypes.c_void_p), stream_ptr(xs[0].device)), "feature_refine_multi")
    return outs


class FR(nn.Module):
    def __init__(self, spatial_scale, points=1):
        super().__init__()
        self.spatial_scale = float(spatial_scale)
        self.points = points

    def forward(self, features, best_rbboxes):
        return feature_refine(features, best_rbboxes, self.spatial_scale, self.points)

    execute = forward

    def __repr__(self):
        return self.__class__.__name__ + '(spatial_scale={}, points={})'.format(self.spatial_scale, self.points)


class FeatureRefineModule(nn.Module):
    """fr.py:291-347: conv_5_1(conv_1_5(x)) + conv_1_1(x) -> FR -> x + refined, per FPN level."""

    def __init__(self, in_channels, featmap_strides, conv_cfg=None, norm_cfg=None):
        super().__init__()
        self.in_channels = in_channels
        self.featmap_strides = featmap_strides
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self.fr = nn.ModuleList([FR(spatial_scale=1 / s) for s in featmap_strides])
        self.conv_5_1 = nn.Conv2d(in_channels, in_channels, kernel_size=(5, 1), stride=1, padding=(2, 0))
        self.conv_1_5 = nn.Conv2d(in_channels, in_channels, kernel_size=(1, 5), stride=1, padding=(0, 2))
        self.conv_1_1 = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.init_weights()

    def init_weights(self):
        for m in (self.conv_5_1, self.conv_1_5, self.conv_1_1):
            nn.init.normal_(m.weight, mean=0.0, std=0.01)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0.0)

    def forward(self, x, best_rbboxes):
        mlvl_rbboxes = [torch.cat(best_rbbox) for best_rbbox in zip(*best_rbboxes)]
        feats = [self.conv_5_1(self.conv_1_5(x_scale)) + self.conv_1_1(x_scale) for x_scale in x]
        if not (torch.is_grad_enabled() and any(f.requires_grad for f in feats)) and len({fr.points for fr in self.fr}) == 1:
            # inference: the per-level FR calls share one launch
            refined = feature_refine_multi(feats, mlvl_rbboxes, [fr.spatial_scale for fr in self.fr], self.fr[0].points)
            return [x_scale + r for x_scale, r in zip(x, refined)]
        return [x_scale + fr_scale(f, b) for x_scale, f, b, fr_scale in zip(x, feats, mlvl_rbboxes, self.fr)]

    execute = forward

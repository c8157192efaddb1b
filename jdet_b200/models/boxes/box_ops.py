"""Rotated box coders used on either side of AlignConv (reference: python/jdet/models/boxes/box_ops.py:176-285).

Elementwise glue in torch; the arithmetic order follows the reference.
`norm_angle` uses Python/floor modulo (`%` on a Jittor Var), i.e. torch.remainder.
"""
import numpy as np
import torch


def norm_angle(angle, range=[float(-np.pi / 4), float(np.pi)]):
    return torch.remainder(angle - range[0], range[1]) + range[0]


def delta2bbox_rotated(rois, deltas, means=(0., 0., 0., 0., 0.), stds=(1., 1., 1., 1., 1.), max_shape=None,
                       wh_ratio_clip=16 / 1000, clip_border=True):
    """rois (N,5), deltas (N, 5*num_classes) -> boxes (N, 5*num_classes)   (box_ops.py:229-285)"""
    reps = deltas.size(1) // 5
    means_t = torch.tensor(means, dtype=deltas.dtype, device=deltas.device).repeat(1, reps)
    stds_t = torch.tensor(stds, dtype=deltas.dtype, device=deltas.device).repeat(1, reps)
    denorm = deltas * stds_t + means_t
    dx, dy, dw, dh, dangle = denorm[:, 0::5], denorm[:, 1::5], denorm[:, 2::5], denorm[:, 3::5], denorm[:, 4::5]
    max_ratio = float(np.abs(np.log(wh_ratio_clip)))
    dw = dw.clamp(min=-max_ratio, max=max_ratio)
    dh = dh.clamp(min=-max_ratio, max=max_ratio)
    roi_x = rois[:, 0].unsqueeze(1).expand_as(dx)
    roi_y = rois[:, 1].unsqueeze(1).expand_as(dy)
    roi_w = rois[:, 2].unsqueeze(1).expand_as(dw)
    roi_h = rois[:, 3].unsqueeze(1).expand_as(dh)
    roi_angle = rois[:, 4].unsqueeze(1).expand_as(dangle)
    gx = dx * roi_w * torch.cos(roi_angle) - dy * roi_h * torch.sin(roi_angle) + roi_x
    gy = dx * roi_w * torch.sin(roi_angle) + dy * roi_h * torch.cos(roi_angle) + roi_y
    gw = roi_w * dw.exp()
    gh = roi_h * dh.exp()
    ga = norm_angle(np.pi * dangle + roi_angle)
    return torch.stack([gx, gy, gw, gh, ga], dim=-1).view_as(deltas)


def rotated_box_to_poly(rrects):
    """(n,5) [x_ctr,y_ctr,w,h,angle] -> (n,8) corner polygon   (box_ops.py:592-613)"""
    n = rrects.shape[0]
    if n == 0:
        return torch.zeros((0, 8), dtype=rrects.dtype, device=rrects.device)
    x_ctr, y_ctr, width, height, angle = rrects[:, 0], rrects[:, 1], rrects[:, 2], rrects[:, 3], rrects[:, 4]
    tl_x, tl_y, br_x, br_y = -width / 2, -height / 2, width / 2, height / 2
    c, s = torch.cos(angle), torch.sin(angle)
    xs = torch.stack([tl_x, br_x, br_x, tl_x], 1)
    ys = torch.stack([tl_y, tl_y, br_y, br_y], 1)
    px = c[:, None] * xs - s[:, None] * ys + x_ctr[:, None]
    py = s[:, None] * xs + c[:, None] * ys + y_ctr[:, None]
    return torch.stack([px, py], 2).reshape(n, 8)

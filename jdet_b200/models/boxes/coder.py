"""Decoders the Oriented R-CNN inference path needs, as plain tensor functions (no registry).

Behaviour follows jdet.models.boxes.coder.{MidpointOffsetCoder, OrientedDeltaXYWHTCoder}.decode
(reference: python/jdet/models/boxes/coder.py:377-438, 486-519) and the small box helpers of
python/jdet/ops/bbox_transforms.py (:499-517 regular_theta / regular_obb, :575-597 rectpoly2obb,
:609-646 obb2poly / obb2hbb).  Encoders (training) are out of scope.
"""
import math

import torch

_MAX_RATIO = abs(math.log(16 / 1000))          # wh_ratio_clip = 16/1000 in both coders


def regular_theta(theta, mode="180", start=-math.pi / 2):
    cycle = 2 * math.pi if mode == "360" else math.pi
    return torch.remainder(theta - start, cycle) + start


def regular_obb(obb):
    """long side first, angle folded into [-pi/2, pi/2)."""
    x, y, w, h, t = obb.unbind(-1)
    swap = ~(w > h)
    return torch.stack([x, y, torch.where(swap, h, w), torch.where(swap, w, h),
                        regular_theta(torch.where(swap, t + math.pi / 2, t))], -1)


def rectpoly2obb(polys):
    """(…,8) rectangle corners -> (…,5) [x,y,w,h,theta]; theta from the first edge, y axis pointing down."""
    pts = polys.reshape(*polys.shape[:-1], 4, 2)
    theta = torch.atan2(-(pts[..., 1, 1] - pts[..., 0, 1]), pts[..., 1, 0] - pts[..., 0, 0])
    c, s = torch.cos(theta), torch.sin(theta)
    ctr = pts.mean(-2, keepdim=True)
    d = pts - ctr
    u = d[..., 0] * c[..., None] - d[..., 1] * s[..., None]         # coordinates in the box frame
    v = d[..., 0] * s[..., None] + d[..., 1] * c[..., None]
    w = u.max(-1)[0] - u.min(-1)[0]
    h = v.max(-1)[0] - v.min(-1)[0]
    return regular_obb(torch.stack([ctr[..., 0, 0], ctr[..., 0, 1], w, h, theta], -1))


def obb2poly(obb):
    ctr, w, h, t = obb[..., :2], obb[..., 2:3], obb[..., 3:4], obb[..., 4:5]
    c, s = torch.cos(t), torch.sin(t)
    a = torch.cat([w / 2 * c, -w / 2 * s], -1)
    b = torch.cat([-h / 2 * s, -h / 2 * c], -1)
    return torch.cat([ctr + a + b, ctr + a - b, ctr - a - b, ctr - a + b], -1)


def obb2hbb(obb):
    ctr, w, h, t = obb[..., :2], obb[..., 2:3], obb[..., 3:4], obb[..., 4:5]
    c, s = torch.cos(t), torch.sin(t)
    bias = torch.cat([(w / 2 * c).abs() + (h / 2 * s).abs(), (w / 2 * s).abs() + (h / 2 * c).abs()], -1)
    return torch.cat([ctr - bias, ctr + bias], -1)


def midpoint_offset_decode(anchors, deltas, means=(0.,) * 6, stds=(1., 1., 1., 1., .5, .5)):
    """anchors (n,4) x1y1x2y2, deltas (n,6) [dx,dy,dw,dh,da,db] -> (n,5) oriented proposals.
    The two midpoint offsets give a parallelogram; its diagonals are equalised to make it a rectangle."""
    d = deltas * deltas.new_tensor(stds) + deltas.new_tensor(means)
    dx, dy, dw, dh, da, db = d.unbind(-1)
    dw, dh = dw.clamp(-_MAX_RATIO, _MAX_RATIO), dh.clamp(-_MAX_RATIO, _MAX_RATIO)
    pw, ph = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
    gx = (anchors[:, 0] + anchors[:, 2]) * 0.5 + pw * dx
    gy = (anchors[:, 1] + anchors[:, 3]) * 0.5 + ph * dy
    gw, gh = pw * dw.exp(), ph * dh.exp()
    x1, y1, x2, y2 = gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5
    da, db = da.clamp(-0.5, 0.5), db.clamp(-0.5, 0.5)
    top, bottom = gx + da * gw, gx - da * gw            # x of the top / bottom edge midpoints' vertices
    right, left = gy + db * gh, gy - db * gh            # y of the right / left ones
    rel = torch.stack([top - gx, y1 - gy, x2 - gx, right - gy, bottom - gx, y2 - gy, x1 - gx, left - gy], -1)
    rel = rel.reshape(-1, 4, 2)
    diag = rel.norm(dim=-1)
    rel = rel * (diag.max(-1, keepdim=True)[0] / diag)[..., None]
    ctr = torch.stack([gx, gy], -1)[:, None, :]
    return rectpoly2obb((rel + ctr).reshape(-1, 8))


def oriented_delta_xywht_decode(rois, deltas, means=(0.,) * 5, stds=(.1, .1, .2, .2, .1)):
    """rois (n,5) [x,y,w,h,theta], deltas (n,5*k) -> (n,5*k) refined oriented boxes (class-agnostic: k = 1)."""
    n = rois.shape[0]
    d = deltas.reshape(n, -1, 5) * deltas.new_tensor(stds) + deltas.new_tensor(means)
    dx, dy, dw, dh, dt = d.unbind(-1)
    dw, dh = dw.clamp(-_MAX_RATIO, _MAX_RATIO), dh.clamp(-_MAX_RATIO, _MAX_RATIO)
    px, py, pw, ph, pt = (rois[:, i:i + 1] for i in range(5))
    c, s = torch.cos(-pt), torch.sin(-pt)
    gx = dx * pw * c - dy * ph * s + px
    gy = dx * pw * s + dy * ph * c + py
    out = regular_obb(torch.stack([gx, gy, pw * dw.exp(), ph * dh.exp(), regular_theta(dt + pt)], -1))
    return out.reshape(n, -1)

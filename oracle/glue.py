"""oracle.glue — numpy restatements of the reference's torch-free glue on either side of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the reference lines it follows
(paths relative to /root/reference/python/jdet/).  Pinned where the reference has a compilable source
(ARF / RIE: oracle/_ref/libref_cpu.so, golden vectors tests/golden/ref_cpu_orn.npz); the box coders are plain
formulas restated in float64 numpy and compared at 1e-5.
"""
import ctypes

import numpy as np


# ---- models/boxes/box_ops.py:176-285 ------------------------------------------------------------------------------------
def norm_angle(angle, lo=-np.pi / 4, rng=np.pi):
    """box_ops.py:176-179: (angle - range[0]) % range[1] + range[0] (Python / floor modulo)"""
    return np.mod(angle - lo, rng) + lo


def delta2bbox_rotated(rois, deltas, means=(0., 0., 0., 0., 0.), stds=(1., 1., 1., 1., 1.), wh_ratio_clip=16 / 1000):
    """box_ops.py:229-285, one class: rois (n,5), deltas (n,5) -> (n,5)"""
    rois, deltas = np.asarray(rois, np.float64), np.asarray(deltas, np.float64)
    d = deltas * np.asarray(stds)[None] + np.asarray(means)[None]                        # :248-250
    dx, dy, dw, dh, da = d.T
    mr = np.abs(np.log(wh_ratio_clip))                                                    # :256
    dw, dh = np.clip(dw, -mr, mr), np.clip(dh, -mr, mr)                                   # :257-258
    x, y, w, h, a = rois.T
    gx = dx * w * np.cos(a) - dy * h * np.sin(a) + x                                      # :265-268
    gy = dx * w * np.sin(a) + dy * h * np.cos(a) + y
    gw, gh = w * np.exp(dw), h * np.exp(dh)                                               # :269-270
    ga = norm_angle(np.pi * da + a)                                                       # :272-273
    return np.stack([gx, gy, gw, gh, ga], 1)


def bbox_decode(bbox_preds, anchors, means=(0., 0., 0., 0., 0.), stds=(1., 1., 1., 1., 1.)):
    """models/roi_heads/s2anet_head.py:631-654: bbox_preds (N,5,H,W), anchors (H*W,5) -> (N,H,W,5), wh_ratio_clip 1e-6"""
    N, _, H, W = bbox_preds.shape
    out = np.zeros((N, H, W, 5))
    for i in range(N):
        delta = np.transpose(bbox_preds[i], (1, 2, 0)).reshape(-1, 5)                     # :647-648
        out[i] = delta2bbox_rotated(anchors, delta, means, stds, wh_ratio_clip=1e-6).reshape(H, W, 5)
    return out


# ---- ops/fr.py:291-347 ---------------------------------------------------------------------------------------------------
def conv2d(x, w, b, pad):
    """plain cross-correlation, stride 1: x (N,C,H,W) f64, w (Co,C,kh,kw), zero padding (ph, pw)"""
    N, C, H, W = x.shape
    Co, _, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad[0], pad[0]), (pad[1], pad[1])))
    out = np.zeros((N, Co, H, W))
    for i in range(kh):
        for j in range(kw):
            out += np.einsum("nchw,oc->nohw", xp[:, :, i:i + H, j:j + W], w[:, :, i, j])
    return out + b[None, :, None, None]


def feature_refine_module(xs, best_rbboxes, strides, w51, b51, w15, b15, w11, b11, feature_refine_fn):
    """FeatureRefineModule.execute (ops/fr.py:331-347): per level conv_5_1(conv_1_5(x)) + conv_1_1(x) -> FR(1/stride) -> x + refined.
    feature_refine_fn = the oracle's feature_refine (oracle.feature_refine)."""
    mlvl = [np.concatenate(lv) for lv in zip(*best_rbboxes)]                              # :339
    outs = []
    for x, boxes, s in zip(xs, mlvl, strides):
        x64 = x.astype(np.float64)
        feat = conv2d(conv2d(x64, w15, b15, (0, 2)), w51, b51, (2, 0)) + conv2d(x64, w11, b11, (0, 0))   # :342-344
        refined = feature_refine_fn(feat.astype(np.float32), boxes.reshape(x.shape[0], x.shape[2], x.shape[3], 5), 1.0 / s, 1)
        outs.append(x + refined)                                                          # :346
    return outs


# ---- ops/orn.py -----------------------------------------------------------------------------------------------------------
def arf_forward(weight, indices):
    """ARF_forward_cpu_kernel (ops/orn.py:136-170) restated: out[i, k, j, indices[l, k] - 1] = weight[i, j, l]"""
    n_out, n_in, n_ori, kh, kw = weight.shape
    n_rot = indices.shape[3]
    n_entry = n_ori * kh * kw
    w = weight.reshape(n_out, n_in, n_entry)
    idx = indices.reshape(n_entry, n_rot).astype(np.int64) - 1
    out = np.zeros((n_out, n_rot, n_in, n_entry), weight.dtype)
    for l in range(n_entry):
        for k in range(n_rot):
            out[:, k, :, idx[l, k]] = w[:, :, l]
    return out.reshape(n_out * n_rot, n_in * n_ori, kh, kw)


def rie_forward(feature, n_ori):
    """RIE_forward_cpu_kernel (ops/orn.py:291-330) restated: (nBatch, nFeature*nOri) -> (mainDirection uint8, aligned)"""
    nb, nc = feature.shape
    nf = nc // n_ori
    f = feature.reshape(nb, nf, n_ori)
    d = np.zeros((nb, nf), np.uint8)
    aligned = np.zeros_like(f)
    fmax = np.finfo(np.float32).max
    for i in range(nb):
        for j in range(nf):
            mx = -fmax
            for l in range(n_ori):
                if f[i, j, l] > mx:
                    mx, d[i, j] = f[i, j, l], l
            for l in range(n_ori):
                aligned[i, j, (l - int(d[i, j]) + n_ori) % n_ori] = f[i, j, l]
    return d, aligned.reshape(nb, nc)


def ref_arf_forward(weight, indices):
    """the reference's own compiled CPU source (oracle/_ref/libref_cpu.so), or None when it did not travel"""
    import oracle
    R = oracle.ref_cpu()
    if R is None or not hasattr(R, "ref_arf_forward_cpu"):
        return None
    w = np.ascontiguousarray(weight, np.float32)
    ind = np.ascontiguousarray(indices, np.uint8)
    n_out, n_in, n_ori, kh, kw = w.shape
    n_rot = ind.shape[3]
    out = np.zeros((n_out * n_rot, n_in * n_ori, kh, kw), np.float32)
    R.ref_arf_forward_cpu(w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ind.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                          n_out, n_in, n_ori, kh, kw, n_rot, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def ref_rie_forward(feature, n_ori):
    import oracle
    R = oracle.ref_cpu()
    if R is None or not hasattr(R, "ref_rie_forward_cpu"):
        return None
    f = np.ascontiguousarray(feature, np.float32)
    nb, nc = f.shape
    d = np.zeros((nb, nc // n_ori), np.uint8)
    al = np.zeros_like(f)
    R.ref_rie_forward_cpu(f.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), d.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                          al.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n_ori, nb, nc // n_ori)
    return d, al


# ---- ops/nms_poly.py:247-252 + data/devkits/result_merge.py:33-131 ------------------------------------------------------
def _clip_area(p, q):
    """area of convex polygon p clipped by convex polygon q (both counter-clockwise), binary64 Sutherland-Hodgman —
    what shapely's Polygon.intersection(...).area returns for valid convex quadrilaterals (GEOS is not under /root/reference;
    PARITY UNPINNED against shapely itself)."""
    pts = [tuple(v) for v in p]
    for e in range(len(q)):
        if not pts:
            break
        ax, ay = q[e]
        bx, by = q[(e + 1) % len(q)]
        ex, ey = bx - ax, by - ay
        out = []
        for i in range(len(pts)):
            x0, y0 = pts[i]
            x1, y1 = pts[(i + 1) % len(pts)]
            si = ex * (y0 - ay) - ey * (x0 - ax)
            sj = ex * (y1 - ay) - ey * (x1 - ax)
            if si >= 0:
                out.append((x0, y0))
            if (si > 0 and sj < 0) or (si < 0 and sj > 0):
                t = si / (si - sj)
                out.append((x0 + t * (x1 - x0), y0 + t * (y1 - y0)))
        pts = out
    if len(pts) < 3:
        return 0.0
    s = 0.0
    for i in range(len(pts)):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % len(pts)]
        s += x0 * y1 - x1 * y0
    return abs(0.5 * s)


def _ccw(poly8):
    p = np.asarray(poly8, np.float64).reshape(4, 2)
    a = 0.5 * sum(p[i, 0] * p[(i + 1) % 4, 1] - p[(i + 1) % 4, 0] * p[i, 1] for i in range(4))
    if a < 0:
        p = p[[0, 3, 2, 1]]
    return p, abs(a)


def iou_poly(poly1, poly2):
    """ops/nms_poly.py:247-252: inter / max(area1 + area2 - inter, 0.01)"""
    p, a1 = _ccw(poly1)
    q, a2 = _ccw(poly2)
    inter = _clip_area(p, q)
    return inter / max(a1 + a2 - inter, 0.01)


def py_cpu_nms_poly_fast(dets, thresh, fast=True):
    """data/devkits/result_merge.py:69-131 (fast=False: py_cpu_nms_poly, :33-66); stable descending order on ties"""
    dets = np.asarray(dets, np.float32)
    obbs = dets[:, :8].astype(np.float64)
    x1, y1 = obbs[:, 0::2].min(1), obbs[:, 1::2].min(1)                                  # :71-74
    x2, y2 = obbs[:, 0::2].max(1), obbs[:, 1::2].max(1)
    order = np.argsort(-dets[:, 8], kind="stable")
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        ovr = np.zeros(rest.size)
        for jj, j in enumerate(rest):
            if fast:
                w = max(0.0, min(x2[i], x2[j]) - max(x1[i], x1[j]))                       # :97-100
                h = max(0.0, min(y2[i], y2[j]) - max(y1[i], y1[j]))
                if not w * h > 0:                                                         # hbb_ovr > 0, :104
                    continue
            ovr[jj] = iou_poly(obbs[i], obbs[j])                                          # :107-109
        order = rest[ovr <= thresh]                                                       # :124-129
    return np.array(keep, np.int64)

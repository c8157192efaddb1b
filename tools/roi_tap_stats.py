#!/usr/bin/env python
"""CPU-only: how many feature-map pixels does a cfg2 RoI really touch?  (numbers quoted in DESIGN.md section 8)

For the bench's 2048 DOTA-shaped RoIs on the 256 x 256 map (7 x 7 bins, 2 x 2 samples, scale 0.25, v1 convention) it
prints, per RoI on average: raw taps (196 samples x 4), taps after merging equal pixels inside a bin (what the gather
loads today), after merging inside a bin row, distinct pixels of the whole RoI, and the bounding-box footprint; then
the share of RoIs / of today's tap traffic that would fit a shared-memory staging buffer of a given size.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _inputs import dota_boxes  # noqa: E402

rng = np.random.default_rng(0)
boxes = dota_boxes(rng, 2048, 1024.0)
scale, PH, PW, g, H, W = 0.25, 7, 7, 2, 256, 256
per_bin, per_row, uniq, footprint = [], [], [], []
for cx, cy, w, h, th in boxes:
    cw, ch, rw, rh = cx * scale - 0.5, cy * scale - 0.5, max(w * scale, 1.0), max(h * scale, 1.0)
    ph, pw, iy, ix = np.meshgrid(np.arange(PH), np.arange(PW), np.arange(g), np.arange(g), indexing="ij")
    yy = -rh / 2 + ph * rh / PH + (iy + .5) * rh / PH / g
    xx = -rw / 2 + pw * rw / PW + (ix + .5) * rw / PW / g
    x = xx * np.cos(th) + yy * np.sin(th) + cw
    y = yy * np.cos(th) - xx * np.sin(th) + ch
    ok = (y >= -1) & (y <= H) & (x >= -1) & (x <= W)
    y, x = np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)
    yl, xl = np.floor(y).astype(int), np.floor(x).astype(int)
    yh, xh = np.minimum(yl + 1, H - 1), np.minimum(xl + 1, W - 1)
    px = np.where(ok[..., None], np.stack([yl * W + xl, yl * W + xh, yh * W + xl, yh * W + xh], -1), -1)
    per_bin.append(sum(len(set(px[i, j].ravel()) - {-1}) for i in range(PH) for j in range(PW)))
    per_row.append(sum(len(set(px[i].ravel()) - {-1}) for i in range(PH)))
    v = px[px >= 0]
    uniq.append(len(np.unique(v)))
    footprint.append(0 if len(v) == 0 else ((v // W).max() - (v // W).min() + 1) * ((v % W).max() - (v % W).min() + 1))
per_bin, per_row, uniq, footprint = map(np.array, (per_bin, per_row, uniq, footprint))
print("per RoI: raw taps 784 | merged per bin %.0f | per bin row %.0f | distinct pixels %.0f (max %d) | footprint %.0f (median %.0f)"
      % (per_bin.mean(), per_row.mean(), uniq.mean(), uniq.max(), footprint.mean(), np.median(footprint)))
for cap in (64, 96, 128, 192, 256, 384):
    m = uniq <= cap
    print("distinct pixels <= %3d: %4.0f %% of the RoIs, %4.0f %% of today's taps, distinct/taps in that class %.2f"
          % (cap, 100 * m.mean(), 100 * per_bin[m].sum() / per_bin.sum(), uniq[m].sum() / max(1, per_bin[m].sum())))

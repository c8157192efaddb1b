"""Seeded synthetic inputs ("DOTA-shaped") shared by tests and bench."""
import numpy as np


def dota_boxes(rng, n, extent=1024.0, lo=8.0, hi=256.0):
    long_side = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    aspect = np.exp(rng.uniform(0.0, np.log(8.0), n))
    short = long_side / aspect
    swap = rng.random(n) < 0.5
    w = np.where(swap, short, long_side)
    h = np.where(swap, long_side, short)
    return np.stack([rng.uniform(0, extent, n), rng.uniform(0, extent, n), w, h,
                     rng.uniform(-np.pi / 2, np.pi / 2, n)], 1).astype(np.float32)


def clustered_boxes(rng, n, per=50, extent=1024.0):
    seeds = (n + per - 1) // per
    base = dota_boxes(rng, seeds, extent)
    b = np.repeat(base, per, 0)[:n].astype(np.float64)
    size = np.sqrt(b[:, 2] * b[:, 3])[:, None]
    b[:, :2] += rng.normal(0, 0.1, (n, 2)) * size
    b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (n, 2)))
    b[:, 4] += rng.normal(0, 0.1, n)
    return b.astype(np.float32)


def tie_free_scores(rng, n):
    return rng.permutation(np.linspace(0.05, 1.0, n)).astype(np.float32)


ADVERSARIAL = np.array([
    [0, 0, 1, 1, 0], [0.5, 0.5, 1, 2, 0], [0, 0, 2, 2, 0], [1, 0, 2, 2, 0], [2, 0, 2, 2, 0],
    [0, 0, 2, 2, np.pi / 2], [0, 0, 2, 2, np.pi / 4], [0, 0, 1e-3, 5, 0.1], [3, 3, 0, 0, 0],
    [0, 0, 2, 2, 1e-7], [10, 10, 20, 8, 0.3], [12, 9, 18, 10, -0.5], [0, 0, 4, 2, 0], [0, 0, 2, 4, np.pi / 2],
    [100, 100, 50, 1e-4, 0.7], [100, 100, 50, 50, 3.0], [100, 100, 50, 50, -3.0], [1e4, 1e4, 30, 10, 0.2],
    [1e4 + 5, 1e4 - 3, 25, 12, -0.4], [0, 0, 2, 2, 0], [2.02, 0, 2, 2, 0], [4.2, 0, 2, 2, 0],
    [0, 0, -3, -2, 0.3], [0.5, 0.2, 3, 2, 0.3], [5, 5, 1e-8, 1e-8, 0], [7, 7, 1e20, 1e20, 0.1]], np.float32)


def s2anet_anchors(rng, n_img, H, W, stride, jitter=True):
    """Grid anchors of size 4*stride (anchor_generator.py:127-152) perturbed like refined anchors."""
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    a = np.zeros((n_img, H, W, 5), np.float32)
    a[..., 0] = xs * stride
    a[..., 1] = ys * stride
    a[..., 2] = 4 * stride
    a[..., 3] = 4 * stride
    if jitter:
        a[..., 0:2] += rng.normal(0, 0.5 * stride, (n_img, H, W, 2))
        a[..., 2:4] *= np.exp(rng.normal(0, 0.4, (n_img, H, W, 2)))
        a[..., 4] = rng.uniform(-np.pi / 2, np.pi / 2, (n_img, H, W))
    return a.astype(np.float32)

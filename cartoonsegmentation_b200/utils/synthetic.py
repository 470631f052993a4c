"""Seeded synthetic inputs shared by tests and bench.py: smooth images and smooth disparity
fields with step edges -- uniform noise has no depth / instance structure."""
import numpy as np


def smooth_image(h, w, seed=1234, n_waves=8, n_ellipses=4):
    """uint8 BGR [h,w,3]: sum of low-frequency sinusoids + a few filled ellipses."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for _ in range(n_waves):
        fx, fy = rng.uniform(-3, 3, 2) * 2 * np.pi / max(h, w)
        ph = rng.uniform(0, 2 * np.pi, 3)
        amp = rng.uniform(0.3, 1.0, 3)
        for c in range(3):
            img[..., c] += amp[c] * np.sin(fx * xx + fy * yy + ph[c])
    img = (img - img.min()) / (img.max() - img.min() + 1e-6)
    for _ in range(n_ellipses):
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        ry, rx = rng.uniform(0.05, 0.2) * h, rng.uniform(0.05, 0.2) * w
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
        img[m] = rng.uniform(0, 1, 3)
    return (img * 255.0).astype(np.uint8)


def smooth_disparity(h, w, seed=4321, lo=4.0, hi=40.0, n_steps=3):
    """float32 [1,1,h,w] disparity in [lo,hi]: smooth ramp + blobs raised by step edges (foreground objects)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = 0.35 + 0.25 * (yy / h) + 0.05 * np.sin(2 * np.pi * xx / w * rng.uniform(0.5, 1.5))
    for _ in range(n_steps):
        cy, cx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w
        ry, rx = rng.uniform(0.1, 0.25) * h, rng.uniform(0.1, 0.25) * w
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
        d[m] += rng.uniform(0.15, 0.3)
    d = (d - d.min()) / (d.max() - d.min() + 1e-6)
    return (lo + (hi - lo) * d).astype(np.float32)[None, None]


def ellipse_masks(h, w, k=8, seed=99):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((k, h, w), bool)
    for i in range(k):
        cy, cx = rng.uniform(0.2, 0.8) * h, rng.uniform(0.2, 0.8) * w
        ry, rx = rng.uniform(0.05, 0.2) * h, rng.uniform(0.05, 0.2) * w
        out[i] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    return out

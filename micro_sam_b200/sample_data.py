"""Seeded synthetic inputs of the BASELINE.json configs; the role micro_sam/sample_data.py:342
(`synthetic_data`) plays for the reference, without skimage."""
from __future__ import annotations

import numpy as np


def lm_tile(shape=(1024, 1024), n_blobs: int = 150, seed: int = 0, dtype="uint16") -> np.ndarray:
    """LM-style tile: bright blobs (radius 8-25 px) on background 100 with shot noise."""
    rng = np.random.default_rng(seed)
    h, w = shape
    img = np.full(shape, 100.0, dtype=np.float32)
    for _ in range(n_blobs):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        r = rng.uniform(8, 25)
        amp = rng.uniform(300, 1500)
        y0, y1 = max(0, int(cy - 3 * r)), min(h, int(cy + 3 * r) + 1)
        x0, x1 = max(0, int(cx - 3 * r)), min(w, int(cx + 3 * r) + 1)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        img[y0:y1, x0:x1] += amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (r / 1.5) ** 2))
    img = rng.poisson(np.maximum(img, 0)).astype(np.float32)
    return np.clip(img, 0, np.iinfo(dtype).max).astype(dtype) if np.issubdtype(np.dtype(dtype), np.integer) else img.astype(dtype)


def em_volume(shape=(4, 512, 512), seed: int = 0) -> np.ndarray:
    """uint8 band-limited noise volume (Gaussian-filtered white noise, sigma 3)."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    v = ndimage.gaussian_filter(rng.standard_normal(shape).astype(np.float32), sigma=(0, 3, 3))
    v = (v - v.min()) / (v.max() - v.min() + 1e-7)
    return (v * 255).astype(np.uint8)


def random_boxes(n: int, shape=(1024, 1024), seed: int = 0) -> np.ndarray:
    """development/benchmark.py:108-116 recipe: xyxy boxes with w,h in [20,100] inside the tile."""
    rng = np.random.default_rng(seed)
    wh = rng.integers(20, 101, size=(n, 2))
    x0 = rng.integers(0, shape[1] - wh[:, 0])
    y0 = rng.integers(0, shape[0] - wh[:, 1])
    return np.stack([x0, y0, x0 + wh[:, 0], y0 + wh[:, 1]], axis=1).astype(np.float64)

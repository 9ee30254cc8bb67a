"""Synthetic disparity frames: seeded numpy, seed = 1000 + frame index.

S1 uniform      u8 uniform in [0,255]
S2 scene        u8 ramp + blobs, 5% rectangular zero holes, 1% salt noise
S3 float-entry  f32 = S1 * 0.125 (what the reference feeds reprojectImageTo3D)
S4 float-stress f32 uniform in [0.1, 32) with sparse zeros (rounding stress, oracle only)
"""
from __future__ import annotations

import numpy as np


def _rng(index: int):
    return np.random.default_rng(1000 + index)


def s1_uniform(h: int, w: int, index: int = 0) -> np.ndarray:
    return _rng(index).integers(0, 256, size=(h, w), dtype=np.uint8)


def s2_scene(h: int, w: int, index: int = 0) -> np.ndarray:
    rng = _rng(index)
    v = np.arange(h, dtype=np.float64)[:, None]
    img = 8.0 * (2.0 + 20.0 * v / h) + np.zeros((1, w))
    for _ in range(8):
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        ry, rx = int(rng.integers(4, max(5, h // 4))), int(rng.integers(4, max(5, w // 4)))
        img[max(0, cy - ry):cy + ry, max(0, cx - rx):cx + rx] += 16.0 * rng.uniform(0.5, 3.0)
    img = np.clip(img, 0, 255).astype(np.uint8)
    holes = 0
    while holes < 0.05 * h * w:
        cy, cx = int(rng.integers(0, h)), int(rng.integers(0, w))
        hh, ww = int(rng.integers(4, max(5, h // 8))), int(rng.integers(4, max(5, w // 8)))
        img[cy:cy + hh, cx:cx + ww] = 0
        holes += hh * ww
    salt = rng.random((h, w)) < 0.01
    img[salt] = rng.integers(0, 256, size=int(salt.sum()), dtype=np.uint8)
    return img


def s3_float(h: int, w: int, index: int = 0) -> np.ndarray:
    return s1_uniform(h, w, index).astype(np.float32) * np.float32(0.125)


def s4_stress(h: int, w: int, index: int = 0) -> np.ndarray:
    rng = _rng(index)
    d = rng.uniform(0.1, 32.0, size=(h, w)).astype(np.float32)
    d[rng.random((h, w)) < 0.002] = 0.0
    return d


def fill_s3(out: np.ndarray, first_index: int = 0) -> None:
    """Fills a (F,H,W) float32 array in place with S3 frames (cheap generator for big pinned buffers)."""
    f, h, w = out.shape
    for i in range(f):
        out[i] = s3_float(h, w, first_index + i)

"""Deterministic synthetic crops and labels of the handposedd shapes.

Real datasets/example depth frames and the trained assets/handposedd.cnnb are absent
from the reference mount, so benchmarks and parity tests use these generators:

* ``uniform_crops``   - uniform[0,1) float32 64x64 crops.
* ``depthlike_crops`` - ~60 % exact zeros plus a smooth blob in (0, 0.5], the value
  distribution the reference's segmented crops have after the normalisation of
  include/handtrack.h:700 (~59 % zeros, max ~0.48, mean ~0.175).
* ``heatmap_labels``  - one Gaussian peak per span (8 spans of 16x16, 16 spans of 16),
  u8-quantised then /255 like include/misc_image.h:248-295 produces, so that span sums
  are NOT exactly 1 (SURVEY.md 8a note 5).
"""
import numpy as np

N_IN = 4096
N_OUT = 2304


def uniform_crops(n, seed=1234):
    return np.random.default_rng(seed).random((n, N_IN), dtype=np.float32)


def depthlike_crops(n, seed=1234):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float32)
    out = np.zeros((n, 64, 64), np.float32)
    for i in range(n):
        cx, cy = rng.uniform(20, 44, 2)
        rx, ry = rng.uniform(14, 22, 2)
        d2 = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2
        blob = np.clip(1.0 - d2, 0.0, None) * rng.uniform(0.3, 0.5)
        blob *= 1.0 + 0.1 * rng.standard_normal((64, 64)).astype(np.float32)
        out[i] = np.where(d2 < 1.0, np.clip(blob, 1e-3, 0.5), 0.0)
    return out.reshape(n, N_IN).astype(np.float32)


def heatmap_labels(n, seed=4321):
    rng = np.random.default_rng(seed)
    t = np.zeros((n, N_OUT), np.float32)
    yy, xx = np.mgrid[0:16, 0:16].astype(np.float32)
    k = np.arange(16, dtype=np.float32)
    for i in range(n):
        for s in range(8):
            cx, cy = rng.uniform(2, 14, 2)
            g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 2.0)
            q = np.floor(g / g.sum() * 255.0)
            t[i, s * 256:(s + 1) * 256] = (q / 255.0).reshape(-1)
        for s in range(16):
            c = rng.uniform(1, 15)
            g = np.exp(-((k - c) ** 2) / 2.0)
            q = np.floor(g / g.sum() * 255.0)
            t[i, 2048 + s * 16:2048 + (s + 1) * 16] = q / 255.0
    return t


def depth_frames(n, h, w, seed, np_=17):
    """Recorded-dataset-shaped frames: 16-bit depth with a hand-range blob, 8-bit IR, np_ poses (xyz + unit quaternion) per frame."""
    rng = np.random.default_rng(seed)
    depth = np.zeros((n, h, w), np.uint16)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):   # a blob of hand-range depth (0.2 .. 0.6 m at depth_scale 0.001) on an empty background
        cy, cx, r = rng.uniform(0.3, 0.7) * h, rng.uniform(0.3, 0.7) * w, rng.uniform(0.2, 0.4) * min(h, w)
        m = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        depth[i][m] = (rng.uniform(200, 600) + 40 * np.sin(xx[m] * 0.3) + rng.uniform(-5, 5, m.sum())).astype(np.uint16)
    ir = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    poses = rng.normal(0, 0.2, (n, np_, 7)).astype(np.float32)
    poses[..., 3:] /= np.linalg.norm(poses[..., 3:], axis=-1, keepdims=True)
    return depth, ir, poses

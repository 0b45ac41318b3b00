"""Deterministic synthetic workloads (BASELINE.json configs 3-5).

Inputs only — no algorithm lives here.  left.png is the KITTI-style fixture shipped in
trackingbench_slam_b200/data (the reference's data/left.png)."""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def load_gray(name):
    import cv2
    img = cv2.imread(os.path.join(_DATA, name), cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(os.path.join(_DATA, name))
    return img


_tile_cache = {}


def _tiled(width, height):
    key = (width, height)
    if key not in _tile_cache:
        base = load_gray('left.png')
        ry, rx = -(-2 * height // base.shape[0]), -(-2 * width // base.shape[1])
        _tile_cache[key] = np.tile(base, (ry, rx))
    return _tile_cache[key]


def synth_frame(k, width=1920, height=1080, seed0=1234):
    """Frame k of the synthetic stream: tiled left.png panned by (k, 3k) px + uniform noise in [-2,2]."""
    t = _tiled(width, height)
    img = np.roll(t, (k % t.shape[0], (3 * k) % t.shape[1]), axis=(0, 1))[:height, :width].astype(np.int16)
    rng = np.random.Generator(np.random.PCG64(seed0 + k))
    img += rng.integers(-2, 3, size=img.shape, dtype=np.int16)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_stream(n, width=1920, height=1080, seed0=1234, out=None):
    out = np.empty((n, height, width), np.uint8) if out is None else out
    for k in range(n):
        out[k] = synth_frame(k, width, height, seed0)
    return out


def synth_descriptors(n, seed=42, frac_matched=0.5, flip_p=0.08):
    """Config 4: Q,T in u8[n,32]; T[perm(i)] = Q[i] ^ mask_i for the first frac*n rows."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    T = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    m = int(n * frac_matched)
    perm = rng.permutation(n)[:m]
    bits = rng.random((m, 256)) < flip_p
    mask = np.packbits(bits, axis=1)
    T[perm] = Q[:m] ^ mask
    return Q, T

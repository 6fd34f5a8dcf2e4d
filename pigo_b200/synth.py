"""Deterministic synthetic grayscale frames (SURVEY.md §8d / BASELINE.md §2 content classes).

  U  uniform noise                 numpy.random.default_rng(seed).integers(0, 256)
  S  smooth: blurred noise (sigma 8 px) min-max rescaled to 0..255 (natural-image-like rejection)
  F  faces: a luma patch (testdata/sample.jpg gray, committed as tests/golden/sample_gray_400x320.u8)
     tiled over the frame with a circular shift, plus optional +-1 LSB noise to break exact ties

Host-side numpy only: these build INPUTS; nothing here is on the detection path.
"""
from __future__ import annotations

import os

import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sample_gray() -> np.ndarray:
    """The 400x320 luma of the reference's testdata/sample.jpg (see tools/make_fixtures.py)."""
    return np.fromfile(os.path.join(_GOLD, "sample_gray_400x320.u8"), dtype=np.uint8).reshape(400, 320)


def frame_noise(rows: int, cols: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(rows, cols), dtype=np.uint8)


def _box_blur(a: np.ndarray, radius: int, passes: int = 3) -> np.ndarray:
    # three box passes approximate a gaussian; wrap-around borders (deterministic, dependency-free)
    for _ in range(passes):
        for ax in (0, 1):
            c = np.cumsum(np.concatenate([a.take(range(-radius - 1, 0), axis=ax), a,
                                          a.take(range(0, radius), axis=ax)], axis=ax), axis=ax, dtype=np.float64)
            n = a.shape[ax]
            hi = c.take(range(2 * radius + 1, 2 * radius + 1 + n), axis=ax)
            lo = c.take(range(0, n), axis=ax)
            a = (hi - lo) / (2 * radius + 1)
    return a


def frame_smooth(rows: int, cols: int, seed: int, sigma: float = 8.0) -> np.ndarray:
    a = np.random.default_rng(seed).random((rows, cols))
    radius = max(1, int(round(sigma * 0.87)))  # 3 box passes of half-width ~0.87 sigma ~ gaussian(sigma)
    b = _box_blur(a, radius)
    b = (b - b.min()) / max(b.max() - b.min(), 1e-12)
    return np.clip(b * 255.0 + 0.5, 0, 255).astype(np.uint8)


def frame_faces(patch: np.ndarray | None, rows: int, cols: int, shift=(0, 0), noise_seed: int | None = None) -> np.ndarray:
    if patch is None:
        patch = sample_gray()
    ph, pw = patch.shape
    reps = (-(-rows // ph) + 1, -(-cols // pw) + 1)
    big = np.tile(patch, reps)
    big = np.roll(big, (int(shift[0]), int(shift[1])), axis=(0, 1))[:rows, :cols]
    if noise_seed is not None:
        n = np.random.default_rng(noise_seed).integers(-1, 2, size=(rows, cols))
        big = np.clip(big.astype(np.int16) + n, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(big, dtype=np.uint8)


def make_batch(nframes: int, rows: int, cols: int, classes: str = "USF", seed0: int = 0, out: np.ndarray | None = None):
    """Frame i has class classes[i % len(classes)] and seed seed0+i.  Returns (nframes, rows, cols) uint8."""
    if out is None:
        out = np.empty((nframes, rows, cols), dtype=np.uint8)
    patch = sample_gray() if "F" in classes else None
    for i in range(nframes):
        k = classes[i % len(classes)]
        if k == "U":
            out[i] = frame_noise(rows, cols, seed0 + i)
        elif k == "S":
            out[i] = frame_smooth(rows, cols, seed0 + i)
        elif k == "F":
            out[i] = frame_faces(patch, rows, cols, shift=((37 * (seed0 + i)) % 400, (53 * (seed0 + i)) % 320),
                                 noise_seed=seed0 + i)
        else:
            raise ValueError(k)
    return out

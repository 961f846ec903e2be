"""Counter-based synthetic images: value(c, y, x) = splitmix64(seed ^ (c<<48 | y<<24 | x)).

u8 images take the top byte; f32/f64 images take the top 24 bits scaled to a 12-bit "medical" range [0, 4096).
Deterministic, seekable by row (so a band of a huge image can be produced without the rest) and uniform --
the worst case for parity, there is no smoothness to hide errors.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def synthetic_image(width: int, height: int, dtype, seed: int, channels: int = 1, y0: int = 0,
                    rows: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """Rows [y0, y0+rows) of the width x height(xchannels) synthetic image with the given seed."""
    dtype = np.dtype(dtype)
    rows = height - y0 if rows is None else rows
    shape = (rows, width) if channels == 1 else (rows, width, channels)
    if out is None:
        out = np.empty(shape, dtype=dtype)
    x = np.arange(width, dtype=np.uint64)[None, :]
    step = max(1, (1 << 22) // max(1, width))
    for c in range(channels):
        for r0 in range(0, rows, step):
            r1 = min(rows, r0 + step)
            y = (np.arange(y0 + r0, y0 + r1, dtype=np.uint64) << np.uint64(24))[:, None]
            key = np.uint64(seed) ^ ((np.uint64(c) << np.uint64(48)) | y | x)
            h = _splitmix64(key)
            if dtype == np.uint8:
                v = (h >> np.uint64(56)).astype(np.uint8)
            else:
                v = ((h >> np.uint64(40)).astype(np.float64) * (4096.0 / 16777216.0)).astype(dtype)
            if channels == 1:
                out[r0:r1] = v
            else:
                out[r0:r1, :, c] = v
    return out

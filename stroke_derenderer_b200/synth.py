"""Synthetic workloads for tests and bench.py.

There is no sample data in the reference (`.gitignore` excludes images/*), so
line images are drawn procedurally: paper 255, ink ~20, glyph-like polylines.
"""

from __future__ import annotations

import cv2
import numpy as np


def synth_line(width: int, seed: int = 0, height: int = 128, density: float = 1.0) -> np.ndarray:
    """RGB u8 (height, width, 3) handwriting-like line image."""
    rng = np.random.default_rng(seed)
    gray = np.full((height, width), 255, np.uint8)
    x = int(rng.integers(2, 12))
    while x < width - 20:
        gw = int(rng.integers(18, 60))
        gh = int(rng.integers(30, 90))
        top = int(rng.integers(10, max(11, height - gh - 5)))
        for _ in range(int(rng.integers(1, 4))):
            pts = np.stack([rng.integers(x, min(x + gw, width - 1) + 1, 4),
                            rng.integers(top, min(top + gh, height - 1) + 1, 4)], axis=1)
            cv2.polylines(gray, [pts.astype(np.int32).reshape(-1, 1, 2)], False, 20,
                          thickness=int(rng.integers(2, 5)), lineType=cv2.LINE_AA)
        x += gw + int(rng.integers(2, 25) / density)
    return np.repeat(gray[:, :, None], 3, axis=2)


def config_widths(n_lines: int, seed: int = 1234, lo: int = 1536, hi: int = 6145) -> np.ndarray:
    """Line widths of BASELINE configs 3 and 4 (n_lines = 512 / 4096)."""
    return np.random.default_rng(seed).integers(lo, hi, n_lines)


def synth_dense_mask(width: int = 16384, p: float = 0.003, seed: int = 0, height: int = 128) -> np.ndarray:
    """Config 5: dense-island stress mask, u8 {0,1} (height, width)."""
    rng = np.random.default_rng(seed)
    m = (rng.random((height, width)) < p).astype(np.uint8)
    return cv2.dilate(m, np.ones((2, 2), np.uint8))


def ink_mask(line_rgb: np.ndarray) -> np.ndarray:
    """Text-like mask straight from a synthetic line (ink darker than 128)."""
    return (line_rgb[:, :, 0] < 128).astype(np.uint8)


def n_tiles_for_width(w: int, tile_w: int = 384, overlap: int = 64) -> int:
    return 1 if w < tile_w else w // (tile_w - overlap) + 1

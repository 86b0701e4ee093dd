"""Seeded synthetic workloads for the BASELINE.json configs.

Nothing here touches the reference or the oracle; both the tests and bench.py draw inputs from
these generators so every side sees identical bytes.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Sequence

import numpy as np


def noise_images(n: int, h: int = 224, w: int = 224, seed: int = 0, channels: int = 3) -> np.ndarray:
    """C1/C2/C3 inputs: uniform uint8 noise, [n, h, w, channels]."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, h, w, channels), dtype=np.uint8)


def mri_like_images(n: int, size: int = 512, seed: int = 0) -> np.ndarray:
    """C4 inputs: smooth gray ellipses + noise on a black background, R==G==B, [n, size, size, 3].

    Mimics the shipped dataset (512x512 JPEGs holding gray content in three identical channels,
    mean ~62, many exact zeros).
    """
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.empty((n, size, size, 3), np.uint8)
    for i in range(n):
        img = np.zeros((size, size), np.float32)
        cy, cx = size * (0.5 + 0.05 * rng.standard_normal(2))
        ry, rx = size * (0.36 + 0.04 * rng.random(2))
        head = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        img += 95.0 * np.clip(1.15 - head, 0.0, 1.0) ** 0.35 * (head < 1.0)
        for _ in range(int(rng.integers(3, 8))):
            by, bx = cy + ry * 0.6 * rng.uniform(-1, 1), cx + rx * 0.6 * rng.uniform(-1, 1)
            sy, sx = size * rng.uniform(0.02, 0.12, 2)
            amp = rng.uniform(-60.0, 110.0)
            img += amp * np.exp(-(((yy - by) / sy) ** 2 + ((xx - bx) / sx) ** 2)) * (head < 1.0)
        img += 9.0 * rng.standard_normal((size, size)).astype(np.float32) * (head < 1.05)
        g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        out[i] = g[:, :, None]
    return out


def ragged_images(shapes: Sequence, seed: int = 0) -> List[np.ndarray]:
    """Variable-size RGB noise images, one per (h, w) in ``shapes``."""
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in shapes]


def write_png_dataset(root: Path, images: Sequence[np.ndarray], n_labeled: int = 32) -> Path:
    """Lay images out as the reference expects (src/feature_extraction.py:72-73,137-170).

    The first ``n_labeled`` go to avec_labels/{cancer,normal}/ alternately, the rest to
    sans_label/.  PNG is lossless, so the file path and the array path see the same pixels.
    """
    from PIL import Image

    root = Path(root)
    for sub in ("avec_labels/cancer", "avec_labels/normal", "sans_label"):
        (root / sub).mkdir(parents=True, exist_ok=True)
    for i, arr in enumerate(images):
        if i < n_labeled:
            sub = "avec_labels/cancer" if i % 2 == 0 else "avec_labels/normal"
        else:
            sub = "sans_label"
        Image.fromarray(arr).save(root / sub / f"img_{i:06d}.png", compress_level=1)
    return root


def dataset_order(n: int, n_labeled: int = 32) -> List[int]:
    """Indices of write_png_dataset's images in discover_image_records order
    (sorted label dirs, then sorted files; labeled before unlabeled)."""
    lab = list(range(min(n, n_labeled)))
    cancer = [i for i in lab if i % 2 == 0]
    normal = [i for i in lab if i % 2 == 1]
    return cancer + normal + list(range(len(lab), n))

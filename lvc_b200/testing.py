"""Seeded synthetic-input generators shared by the parity tests, the golden generator and bench.py
(COCO-shaped boxes -- centres uniform in the image, log-uniform sides in [8, 600], clipped)."""
import numpy as np


def coco_like_boxes(rng, n, W=1333, H=800, min_side=8.0, max_side=600.0):
    cx = rng.uniform(0, W, n)
    cy = rng.uniform(0, H, n)
    w = np.exp(rng.uniform(np.log(min_side), np.log(max_side), n))
    h = np.exp(rng.uniform(np.log(min_side), np.log(max_side), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = b[:, 0::2].clip(0, W)
    b[:, 1::2] = b[:, 1::2].clip(0, H)
    return b.astype(np.float32)


def distinct_scores(rng, n, lo=0.0, hi=1.0):
    """Tie-free fp32 scores (sort order on ties is implementation-defined in the reference)."""
    s = (rng.permutation(n).astype(np.float64) + 0.5) / n
    return (lo + (hi - lo) * s).astype(np.float32)

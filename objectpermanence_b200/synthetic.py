"""Synthetic CATER-shaped inputs.

Mirrors the value distribution the reference loaders produce
(``baselines/datasets.py:130-196`` for the 5-track models, ``:265-336`` for the 6-track
OPNet family): per video 5..10 real objects, slot 0 is the snitch, remaining slots are
all-zero padding; a visible row is ``[x1,y1,x2,y2,1(,is_cone)]`` with integer-pixel
coordinates divided by the 320x240 frame; an invisible row is zeros, except that an
invisible *cone* keeps its ``is_cone`` bit in the 6-track layout (datasets.py:315-318).
Labels are a smooth normalised xyxy track; the mask is constant along the last axis
(datasets.py:487-488).

Pure host-side numpy: no GPU, no reference import.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

MAX_OBJECTS = 15
FRAME_W, FRAME_H = 320, 240


def make_batch(batch: int, frames: int, features: int, seed: int = 1234) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (boxes float32 [B,T,15,F], labels float32 [B,T,4], mask bool [B,T,4])."""
    if features not in (5, 6):
        raise ValueError("features must be 5 (baseline/non-linear/transformer) or 6 (OPNet family)")
    rng = np.random.default_rng(seed)
    boxes = np.zeros((batch, frames, MAX_OBJECTS, features), dtype=np.float32)
    labels = np.zeros((batch, frames, 4), dtype=np.float32)
    mask = np.zeros((batch, frames, 4), dtype=bool)
    for b in range(batch):
        n_obj = int(rng.integers(5, 11))
        is_cone = rng.random(MAX_OBJECTS) < 0.3
        is_cone[0] = False  # the snitch is not a cone
        visible = rng.random((frames, n_obj)) < 0.85
        x1 = rng.integers(0, 280, size=(frames, n_obj))
        y1 = rng.integers(0, 200, size=(frames, n_obj))
        w = rng.integers(5, 41, size=(frames, n_obj))
        h = rng.integers(5, 41, size=(frames, n_obj))
        rows = np.stack([x1 / FRAME_W, y1 / FRAME_H, (x1 + w) / FRAME_W, (y1 + h) / FRAME_H,
                         np.ones_like(x1, dtype=np.float64)], axis=-1)
        rows = rows * visible[..., None]
        boxes[b, :, :n_obj, :5] = rows.astype(np.float32)
        if features == 6:
            boxes[b, :, :n_obj, 5] = is_cone[None, :n_obj].astype(np.float32)
            # a visible non-cone / invisible non-cone keeps 0; an invisible cone keeps its bit
        # smooth ground-truth track: random walk of a box centre, clipped into the frame
        steps = rng.normal(0.0, 2.0, size=(frames, 2)).cumsum(axis=0)
        cx = np.clip(160 + steps[:, 0], 20, 300)
        cy = np.clip(120 + steps[:, 1], 20, 220)
        bw, bh = rng.integers(10, 30), rng.integers(10, 30)
        gt = np.stack([np.floor(cx - bw / 2) / FRAME_W, np.floor(cy - bh / 2) / FRAME_H,
                       np.floor(cx + bw / 2) / FRAME_W, np.floor(cy + bh / 2) / FRAME_H], axis=-1)
        labels[b] = gt.astype(np.float32)
        mask[b] = (rng.random(frames) < 0.8)[:, None]
    return boxes, labels, mask

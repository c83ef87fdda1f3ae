"""Seeded synthetic workload: 2048x1536 RGB "manga pages" with 32 axis-aligned text boxes filled with
dark strokes.  The boxes are the OCR workload (independent of what a random-weight detector emits) and their dilated
union is the inpainting mask.  Page i uses numpy PCG64 seed 20260922 + i, so every rank / run sees the same pages."""
from __future__ import annotations

import cv2
import numpy as np

from .compat import Quadrilateral

BASE_SEED = 20260922


def make_page(index: int, h: int = 2048, w: int = 1536, n_boxes: int = 32):
    """Returns (page uint8 [h,w,3], list of box corner arrays int64 [4,2], mask uint8 [h,w] in {0,255})."""
    rng = np.random.default_rng(BASE_SEED + index)
    grey = np.clip(235 + 10 * rng.standard_normal((h, w)), 0, 255).astype(np.uint8)
    page = np.repeat(grey[:, :, None], 3, axis=2)
    occupied = np.zeros((h, w), bool)
    boxes = []
    tries = 0
    while len(boxes) < n_boxes and tries < 20000:
        tries += 1
        horizontal = len(boxes) < n_boxes // 2
        thick, length = int(rng.integers(48, 65)), int(rng.integers(128, min(769, (w if horizontal else h) - 32)))
        bw, bh = (length, thick) if horizontal else (thick, length)
        x0, y0 = int(rng.integers(8, w - bw - 8)), int(rng.integers(8, h - bh - 8))
        if occupied[max(0, y0 - 12):y0 + bh + 12, max(0, x0 - 12):x0 + bw + 12].any():
            continue
        occupied[y0:y0 + bh, x0:x0 + bw] = True
        # strokes: short dark bars, "characters" every ~thick pixels along the writing direction
        n_chars = max(2, length // thick)
        for c in range(n_chars):
            cx0 = x0 + (c * bw) // n_chars if horizontal else x0
            cy0 = y0 if horizontal else y0 + (c * bh) // n_chars
            cw, ch = (bw // n_chars, bh) if horizontal else (bw, bh // n_chars)
            for _ in range(int(rng.integers(3, 8))):
                sx, sy = cx0 + int(rng.integers(2, max(3, cw - 6))), cy0 + int(rng.integers(2, max(3, ch - 6)))
                if rng.random() < 0.5:
                    page[sy:sy + int(rng.integers(2, 5)), sx:min(sx + int(rng.integers(6, max(7, cw - 4))), cx0 + cw - 1)] = int(rng.integers(0, 50))
                else:
                    page[sy:min(sy + int(rng.integers(6, max(7, ch - 4))), cy0 + ch - 1), sx:sx + int(rng.integers(2, 5))] = int(rng.integers(0, 50))
        boxes.append(np.array([[x0, y0], [x0 + bw, y0], [x0 + bw, y0 + bh], [x0, y0 + bh]], dtype=np.int64))
    mask = np.zeros((h, w), np.uint8)
    for b in boxes:
        cv2.fillPoly(mask, [b.astype(np.int32)], 255)
    mask = cv2.dilate(mask, np.ones((11, 11), np.uint8))      # union dilated by 5 px
    return page, boxes, mask


def make_quads(boxes):
    return [Quadrilateral(b.copy(), "", 1.0) for b in boxes]

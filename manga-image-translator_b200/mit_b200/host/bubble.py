"""`is_ignore` of the OCR stage (utils/bubble.py:28-84), used when OcrConfig.ignore_bubble is in [1, 50]: decide from the 2-pixel
frame of a text-line crop whether it sits in a plain white / black speech bubble; crops whose frame is mixed, or that contain
colour, are skipped by the recogniser (model_48px_ctc.py:90-93)."""
from __future__ import annotations

import cv2
import numpy as np


def has_colour(image: np.ndarray) -> bool:
    """utils/bubble.py:4-26: more than 10 pixels whose squared distance to their own luma (0.299, 0.587, 0.114) exceeds 100."""
    luma = np.dot(image[..., :3], [0.299, 0.587, 0.114])[..., None]
    return bool(np.sum(np.sum((image - luma) ** 2, axis=-1) > 100) > 10)


def is_ignore(region_img: np.ndarray, ignore_bubble: int = 0) -> bool:
    if ignore_bubble < 1 or ignore_bubble > 50:
        return False
    _, binary = cv2.threshold(region_img, 127, 255, cv2.THRESH_BINARY)
    h, w = binary.shape[:2]
    frame = (binary[0:2, 0:w], binary[h - 2:h, 0:w], binary[2:h - 2, 0:2], binary[2:h - 2, w - 2:w])
    dark = sum(int(np.count_nonzero(f == 0)) for f in frame)            # channel values at or below 127, counted per channel
    total = sum(f.size for f in frame)
    ratio = round(dark / total, 6) * 100
    if ignore_bubble <= ratio <= 100 - ignore_bubble:
        return True
    return has_colour(region_img)

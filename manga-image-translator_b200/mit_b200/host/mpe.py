"""LaMa masked positional encoding tables on the host (behaviour of LamaFourier.load_masked_position_encoding,
inpainting/inpainting_lama_mpe.py:751-815) with binary morphology instead of float filter2D passes:
the mask is reduced to 256x256 (INTER_AREA, any coverage counts), the known region is dilated 3x3 step by step;
pos = step at which a hole pixel is reached, direct[k] = reached from the k-th 2x2 diagonal neighbourhood."""
from __future__ import annotations

import cv2
import numpy as np

_BOX = np.ones((3, 3), np.uint8)
_CORNERS = [np.array(k, np.uint8) for k in (
    [[1, 1, 0], [1, 1, 0], [0, 0, 0]], [[0, 0, 0], [1, 1, 0], [1, 1, 0]],
    [[0, 1, 1], [0, 1, 1], [0, 0, 0]], [[0, 0, 0], [0, 1, 1], [0, 1, 1]])]


def _grow(known: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    # correlation with a 0/1 kernel followed by ">0" == cv2.dilate with the same kernel; borders REFLECT_101 like filter2D
    p = cv2.copyMakeBorder(known, 1, 1, 1, 1, cv2.BORDER_REFLECT_101)
    return cv2.dilate(p, kernel, borderType=cv2.BORDER_CONSTANT, borderValue=0)[1:-1, 1:-1]


def mpe_tables_256(mask01: np.ndarray):
    """The tables at the 256x256 working resolution (before the reference's INTER_NEAREST upsampling, which libmitb does on
    the device): (rel_pos int32 [256,256] in [0,127], direct int32 [256,256,4])."""
    return _tables_256(_mask_u8(mask01))


def _mask_u8(mask01) -> np.ndarray:
    """(mask*255).astype(u8) of the reference; integer/bool masks take a u8-only path (no float page)."""
    a = np.asarray(mask01)
    if a.dtype == np.bool_ or (a.dtype.kind in "iu" and a.size and int(a.max()) <= 1 and int(a.min()) >= 0):
        return a.astype(np.uint8) * np.uint8(255)
    return (a.astype(np.float32) * 255).astype(np.uint8)


def small_mask_256(mask01) -> np.ndarray:
    """The reduction step of load_masked_position_encoding (:763-765): (mask*255) u8 -> cv2 INTER_AREA 256x256 (any value > 0 is
    hole).  Stays on the host (cv2's fixed-point area filter is the definition); the distance sweep runs on the device."""
    return cv2.resize(_mask_u8(mask01), (256, 256), interpolation=cv2.INTER_AREA)


def _tables_256(m: np.ndarray):
    small = cv2.resize(m, (256, 256), interpolation=cv2.INTER_AREA)
    known = (small == 0).astype(np.uint8)
    pos = np.zeros((256, 256), np.int32)
    direct = np.zeros((256, 256, 4), np.int32)
    step = 0
    if known.any():
        while not known.all():
            step += 1
            grown = _grow(known, _BOX)
            fresh = (grown > 0) & (known == 0)
            pos[fresh] = step
            for k, ker in enumerate(_CORNERS):
                direct[(_grow(known, ker) > 0) & (known == 0), k] = 1
            known = grown
    rel = np.clip((pos / 128.0 * 128).astype(np.int32), 0, 127)
    return rel.astype(np.int32), direct


def mpe_tables(mask01: np.ndarray):
    """mask01 [H,W] with 1 inside the hole -> (rel_pos int32 [H,W] in [0,127], direct int32 [H,W,4] in {0,1})."""
    m = _mask_u8(mask01)
    H, W = m.shape
    rel, direct = _tables_256(m)
    if (H, W) != (256, 256):
        rel = cv2.resize(rel, (W, H), interpolation=cv2.INTER_NEAREST)
        direct = cv2.resize(direct, (W, H), interpolation=cv2.INTER_NEAREST)
        hole = m != 0
        rel = np.where(hole, rel, 0)
        direct = np.where(hole[..., None], direct, 0)
    return rel.astype(np.int32), direct.astype(np.int32)

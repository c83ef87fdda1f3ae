"""Long-strip handling of the detector: pages that are both very tall (or wide) relative to the detect size and of
extreme aspect ratio are cut into overlapping square-ish patches, several patches are laid side by side into one
square network input, and the outputs are stitched back with overlap averaging.
Behaviour of det_rearrange_forward / square_pad_resize (manga_translator/utils/generic.py:849-997)."""
from __future__ import annotations

import math
from typing import Callable, List, Tuple

import cv2
import numpy as np


def square_pad_resize(img: np.ndarray, tgt: int):
    h, w = img.shape[:2]
    pad_h = pad_w = 0
    if w < h:
        pad_w = h - w
        w = h
    elif h < w:
        pad_h = w - h
        h = w
    if tgt - h > 0:
        pad_h += tgt - h
        pad_w += tgt - h
    if pad_h > 0 or pad_w > 0:
        img = cv2.copyMakeBorder(img, 0, pad_h, 0, pad_w, cv2.BORDER_CONSTANT)
    scale = tgt / img.shape[0]
    assert scale <= 1
    if scale < 1:
        img = cv2.resize(img, (tgt, tgt), interpolation=cv2.INTER_LINEAR)
    return img, scale, pad_h, pad_w


def needs_rearrange(h: int, w: int, tgt: int) -> bool:
    if h < w:
        h, w = w, h
    return (h / tgt) > 2.5 and (h / w) > 3


def rearrange_forward(img: np.ndarray, batch_forward: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], tgt: int = 1280,
                      max_batch: int = 4):
    """Returns (db [1,2,H,W], mask [1,1,H/2,W/2]) stitched at the scale the network ran, or (None, None) when the page
    does not qualify.  ``batch_forward`` maps uint8 [n,S,S,3] to (db [n,2,S,S], mask [n,1,S/2,S/2]) numpy arrays."""
    H0, W0 = img.shape[:2]
    if not needs_rearrange(H0, W0, tgt):
        return None, None
    transpose = H0 < W0
    if transpose:
        img = img.transpose(1, 0, 2)
    h, w = img.shape[:2]
    per_row = max(int(math.floor(2 * tgt / w)), 2)          # patches laid side by side in one network input
    psize = per_row * w                                      # patch height == side of the composed square
    n_patch = int(math.ceil(h / psize))
    step = int((h - psize) / (n_patch - 1)) if n_patch > 1 else 0
    patches = [img[i * step: i * step + psize] for i in range(n_patch)]
    rel_top = [i * step / h for i in range(n_patch)]
    n_group = int(math.ceil(n_patch / per_row))
    pad_num = n_group * per_row - n_patch
    patches += [np.zeros_like(patches[0]) for _ in range(pad_num)]

    squares, scale, pad_size = [], 1.0, 0
    for g in range(n_group):
        grp = patches[g * per_row:(g + 1) * per_row]
        comp = np.concatenate(grp, axis=1)                   # [psize, per_row*w, c]
        if transpose:
            comp = comp.transpose(1, 0, 2)
        sq, scale, ph, pw = square_pad_resize(comp, tgt)
        assert ph == pw
        pad_size = ph
        squares.append(sq)

    dbs: List[np.ndarray] = []
    masks: List[np.ndarray] = []
    for i in range(0, len(squares), max_batch):
        db, mask = batch_forward(np.array(squares[i:i + max_batch]))
        for d, m in zip(db, mask):
            if pad_size > 0:
                pd, pm = int(db.shape[-1] / tgt * pad_size), int(mask.shape[-1] / tgt * pad_size)
                d, m = d[..., :-pd, :-pd], m[..., :-pm, :-pm]
            dbs.append(d)
            masks.append(m)

    def stitch(outs: List[np.ndarray], channels: int):
        side = outs[0].shape[-1]
        ostep = int(step * side / psize)
        pw_out = int(side / per_row)
        h_out = int(pw_out / w * h)
        canvas = np.zeros((channels, h_out, pw_out), dtype=np.float32)
        total = len(outs) * per_row - pad_num
        done = False
        for gi, p in enumerate(outs):
            if transpose:
                p = p.transpose(0, 2, 1)
            for j in range(per_row):
                k = gi * per_row + j
                t = int(round(rel_top[k] * h_out))
                b = min(t + side, h_out)
                canvas[:, t:b, :] += p[:, :b - t, j * pw_out:(j + 1) * pw_out]
                if k > 0:
                    canvas[:, t:t + (side - ostep), :] /= 2.0
                if k >= total - 1:
                    done = True
                    break
            if done:
                break
        if transpose:
            canvas = canvas.transpose(0, 2, 1)
        return canvas[None]

    return stitch(dbs, 2), stitch(masks, 1)

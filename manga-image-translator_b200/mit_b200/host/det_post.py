"""Detector pre/post-processing on the host (CPU, cv2) for stand-alone use; behaviour of
  imgproc.resize_aspect_ratio          detection/default_utils/imgproc.py:37-70
  SegDetectorRepresenter               detection/default_utils/dbnet_utils.py:8-187
  adjustResultCoordinates              detection/default_utils/craft_utils.py:238-244
The reference offsets the box with pyclipper (JT_ROUND) and re-fits a minAreaRect; pyclipper/shapely are third-party and
absent here, so ``unclip`` expands the (integer-truncated, as Clipper does) rectangle analytically by the same distance
area*ratio/perimeter -- identical up to Clipper's integer rounding of the arc points (+-1 px, documented in DESIGN.md).
"""
from __future__ import annotations

import cv2
import numpy as np

from .geometry import polygon_area, polygon_perimeter


def resize_aspect_ratio(img, square_size, interpolation=cv2.INTER_LINEAR, mag_ratio=1):
    """Scale so the longer side equals mag_ratio*square_size (up or down), zero-pad bottom/right to multiples of 256."""
    h, w, c = img.shape
    ratio = (mag_ratio * square_size) / max(h, w)
    th, tw = int(round(h * ratio)), int(round(w * ratio))
    proc = cv2.resize(img, (tw, th), interpolation=interpolation)
    pad_h = (256 - th % 256) % 256
    pad_w = (256 - tw % 256) % 256
    canvas = np.zeros((th + pad_h, tw + pad_w, c), dtype=np.uint8)
    canvas[:th, :tw] = proc
    return canvas, ratio, (int((tw + pad_w) / 2), int((th + pad_h) / 2)), pad_w, pad_h


def mini_box(contour):
    """minAreaRect corners ordered (top-left, top-right, bottom-right, bottom-left) + the short side (dbnet_utils.py:154-172)."""
    rect = cv2.minAreaRect(contour)
    pts = sorted(list(cv2.boxPoints(rect)), key=lambda p: p[0])
    (i1, i4) = (0, 1) if pts[1][1] > pts[0][1] else (1, 0)
    (i2, i3) = (2, 3) if pts[3][1] > pts[2][1] else (3, 2)
    return [pts[i1], pts[i2], pts[i3], pts[i4]], min(rect[1])


def box_score(prob: np.ndarray, contour: np.ndarray) -> float:
    """Mean probability inside the contour polygon (dbnet_utils.py:174-187)."""
    h, w = prob.shape[:2]
    box = contour.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype(np.int32), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype(np.int32), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype(np.int32), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype(np.int32), 0, h - 1)
    m = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] -= xmin
    box[:, 1] -= ymin
    cv2.fillPoly(m, box.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(prob[ymin:ymax + 1, xmin:xmax + 1], m)[0]


def unclip(box4: np.ndarray, unclip_ratio: float) -> np.ndarray:
    """Offset a rectangle outwards by area*ratio/perimeter (dbnet_utils.py:146-152)."""
    pts = np.asarray(box4, dtype=np.float64)
    d = polygon_area(pts) * unclip_ratio / max(polygon_perimeter(pts), 1e-9)
    q = np.trunc(pts)                       # Clipper works on integers; pyclipper truncates the float input
    c = q.mean(axis=0)
    e1, e2 = q[1] - q[0], q[3] - q[0]
    n1, n2 = np.linalg.norm(e1), np.linalg.norm(e2)
    if n1 == 0 or n2 == 0:
        return q.reshape(-1, 1, 2).astype(np.float32)
    u1, u2 = e1 / n1, e2 / n2
    h1, h2 = n1 / 2 + d, n2 / 2 + d
    out = np.array([c - u1 * h1 - u2 * h2, c + u1 * h1 - u2 * h2, c + u1 * h1 + u2 * h2, c - u1 * h1 + u2 * h2])
    return np.round(out).reshape(-1, 1, 2).astype(np.float32)


def boxes_from_prob(prob: np.ndarray, thresh: float, box_thresh: float, unclip_ratio: float, dest_w: int, dest_h: int,
                    max_candidates: int = 1000, min_size: int = 3):
    """SegDetectorRepresenter.boxes_from_bitmap (dbnet_utils.py:96-144): rows of rejected contours stay all-zero."""
    bitmap = prob > thresh
    h, w = bitmap.shape
    # findContours treats every non-zero pixel as 1: the bool map viewed as u8 gives the contours of (bitmap*255) without
    # materialising an int64 page (10 ms on 2048x1536)
    contours, _ = cv2.findContours(np.ascontiguousarray(bitmap).view(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    n = min(len(contours), max_candidates)
    boxes = np.zeros((n, 4, 2), dtype=np.int64)
    scores = np.zeros((n,), dtype=np.float32)
    for i in range(n):
        contour = contours[i].squeeze(1)
        pts, sside = mini_box(contour)
        if sside < min_size:
            continue
        score = box_score(prob, contour)
        if box_thresh > score:
            continue
        box, sside = mini_box(unclip(np.array(pts), unclip_ratio))
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / w * dest_w), 0, dest_w)
        box[:, 1] = np.clip(np.round(box[:, 1] / h * dest_h), 0, dest_h)
        box = np.roll(box, 4 - int(box.sum(axis=1).argmin()), 0)
        boxes[i] = box.astype(np.int64)
        scores[i] = score
    return boxes, scores


def polys_from_boxes(boxes, scores, ratio_w, ratio_h):
    """The zero-row filter + rescale of DBConvNextDetector._infer (dbnet_convnext.py:563-571).  Reference quirk kept: the
    polygons are filtered but later zipped with the UNFILTERED scores."""
    if boxes.size == 0:
        return []
    keep = boxes.reshape(boxes.shape[0], -1).sum(axis=1) > 0
    polys = boxes[keep].astype(np.float64)
    if len(polys):
        polys = polys * np.array([ratio_w, ratio_h], dtype=np.float64)      # adjustResultCoordinates, ratio_net=1
    return polys.astype(np.int64)

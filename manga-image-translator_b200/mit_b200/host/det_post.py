"""Detector pre/post-processing on the host (CPU, cv2) for stand-alone use; behaviour of
  imgproc.resize_aspect_ratio          detection/default_utils/imgproc.py:37-70
  SegDetectorRepresenter               detection/default_utils/dbnet_utils.py:8-187
  adjustResultCoordinates              detection/default_utils/craft_utils.py:238-244
The reference offsets the box with pyclipper (JT_ROUND, ET_CLOSEDPOLYGON) and re-fits a minAreaRect.  pyclipper/shapely are
third-party and absent here (pyclipper is unpinned in requirements.txt; every release since 1.0 bundles Angus Johnson's
Clipper 6.4.2), so ``clipper_offset_round`` restates ClipperOffset (AddPath / FixOrientations / DoOffset / OffsetPoint /
DoRound, clipper.cpp 6.4.2) for one closed path: integer input (pyclipper truncates floats), unit normals, arc step count from
ArcTolerance 0.25, every output vertex rounded half away from zero.  The union pass Clipper runs afterwards only removes
duplicate / collinear vertices of an outward offset of a convex quad, which cannot change the minAreaRect fitted next.
Parity unpinned by a run of the real library (absent offline); pinned by hand-derived vectors in tests/test_host.py.
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


def _c_round(v: float) -> int:
    """Clipper's Round(): half away from zero, via truncation of v +- 0.5."""
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def clipper_offset_round(path, delta: float, arc_tolerance: float = 0.25):
    """ClipperOffset().AddPath(path, JT_ROUND, ET_CLOSEDPOLYGON); Execute(delta) for ONE closed path (Clipper 6.4.2).
    `path`: sequence of (x, y), truncated to integers like pyclipper does.  Returns the offset polygon's vertices as a list of
    integer (x, y) (before Clipper's clean-up union), or [] when Clipper would drop the path (< 3 distinct vertices)."""
    import math
    pts = [(int(p[0]), int(p[1])) for p in path]                     # int(): truncation toward zero
    # ---- AddPath: strip the closing duplicate and consecutive duplicates
    hi = len(pts) - 1
    while hi > 0 and pts[0] == pts[hi]:
        hi -= 1
    src = [pts[0]]
    for i in range(1, hi + 1):
        if src[-1] != pts[i]:
            src.append(pts[i])
    if len(src) < 3:
        return []
    # ---- FixOrientations: closed polygons must have Orientation() == true, i.e. Clipper's Area() >= 0
    a = 0.0
    j = len(src) - 1
    for i in range(len(src)):
        a += (float(src[j][0]) + src[i][0]) * (float(src[j][1]) - src[i][1])
        j = i
    if -a * 0.5 < 0:
        src.reverse()
    n = len(src)
    if delta == 0:
        return list(src)
    # ---- DoOffset: arc discretisation
    y = arc_tolerance
    if arc_tolerance <= 0.0:
        y = 0.25
    elif arc_tolerance > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin, m_cos = math.sin(2 * math.pi / steps), math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0.0:
        m_sin = -m_sin

    def unit_normal(p1, p2):
        if p1 == p2:
            return (0.0, 0.0)
        dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        return (dy * f, -dx * f)

    normals = [unit_normal(src[i], src[(i + 1) % n]) for i in range(n)]
    out = []
    k = n - 1
    for j in range(n):
        # ---- OffsetPoint(j, k, jtRound)
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
            if cos_a > 0:                                         # nearly straight: one vertex
                out.append((_c_round(src[j][0] + normals[k][0] * delta), _c_round(src[j][1] + normals[k][1] * delta)))
                done = True
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if not done:
            if sin_a * delta < 0:                                 # concave corner
                out.append((_c_round(src[j][0] + normals[k][0] * delta), _c_round(src[j][1] + normals[k][1] * delta)))
                out.append(src[j])
                out.append((_c_round(src[j][0] + normals[j][0] * delta), _c_round(src[j][1] + normals[j][1] * delta)))
            else:                                                 # DoRound
                ang = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
                nsteps = max(int(_c_round(steps_per_rad * abs(ang))), 1)
                x, yy = normals[k]
                for _ in range(nsteps):
                    out.append((_c_round(src[j][0] + x * delta), _c_round(src[j][1] + yy * delta)))
                    x2 = x
                    x = x * m_cos - m_sin * yy
                    yy = x2 * m_sin + yy * m_cos
                out.append((_c_round(src[j][0] + normals[j][0] * delta), _c_round(src[j][1] + normals[j][1] * delta)))
        k = j
    return out


def unclip(box4: np.ndarray, unclip_ratio: float) -> np.ndarray:
    """SegDetectorRepresenter.unclip (dbnet_utils.py:146-152): offset the box outwards by area*ratio/perimeter with a round join.
    shapely's Polygon(box).area / .length are the shoelace area and the perimeter of the float box."""
    pts = np.asarray(box4, dtype=np.float64)
    d = polygon_area(pts) * unclip_ratio / max(polygon_perimeter(pts), 1e-9)
    poly = clipper_offset_round(pts, d)
    if not poly:                                                  # Clipper drops degenerate paths; the reference would then fail in minAreaRect
        return np.trunc(pts).reshape(-1, 1, 2).astype(np.float32)
    return np.array(poly, dtype=np.int32).reshape(-1, 1, 2)


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
        if contours[i].shape[0] < 3 and min_size > 0:
            continue        # 1- or 2-point contour: its minAreaRect has a zero side, i.e. sside = 0 < min_size (same row left all-zero)
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

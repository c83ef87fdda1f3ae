"""Host-side geometry used around the three forwards when the real ``manga_translator`` package is not importable
(stand-alone use, tests, bench).  Behavioural mirror of the reference helpers, written without shapely:

  Quadrilateral / sort_pnts           manga_translator/utils/generic.py:324-599
  quadrilateral_can_merge_region      manga_translator/utils/generic.py:653-698
  text-direction grouping             manga_translator/ocr/common.py:12-39

Areas / distances that the reference delegates to shapely (third-party, absent here) are restated: convex-hull
area = shoelace over the monotone-chain hull; polygon distance = 0 when the quads intersect, else the minimum
vertex-to-edge distance (exact for simple polygons).
"""
from __future__ import annotations

import functools
import itertools
from collections import Counter, namedtuple
from typing import List

import cv2
import numpy as np

AABB = namedtuple("AABB", "x y w h")


def order_points(pts: np.ndarray):
    """Canonical corner order + orientation flag, as generic.py:324-353: the two long sides decide whether the line is
    vertical; points come back clockwise from the top-left (horizontal) or sorted top pair / bottom pair (vertical)."""
    pts = np.asarray(pts)
    assert pts.shape == (4, 2)
    diff = (pts[:, None] - pts[None]).reshape(16, -1)
    norms = np.linalg.norm(diff, axis=1)
    longs = diff[np.argsort(norms)[[8, 10]]]
    if (longs[0] * longs[1]).sum() < 0:
        longs[0] = -longs[0]
    s = np.abs(longs.mean(axis=0))
    vertical = bool(s[0] <= s[1])
    if vertical:
        pts = pts[np.argsort(pts[:, 1])]
        pts = pts[[*np.argsort(pts[:2, 0]), *(np.argsort(pts[2:, 0])[::-1] + 2)]]
        return pts, True
    pts = pts[np.argsort(pts[:, 0])]
    out = np.zeros_like(pts)
    out[[0, 3]] = sorted(pts[[0, 1]], key=lambda p: p[1])
    out[[1, 2]] = sorted(pts[[2, 3]], key=lambda p: p[1])
    return out, False


def _hull(points: np.ndarray) -> np.ndarray:
    p = sorted(set(map(tuple, np.asarray(points, dtype=np.float64))))
    if len(p) <= 2:
        return np.array(p, dtype=np.float64).reshape(-1, 2)

    def half(seq):
        h = []
        for q in seq:
            while len(h) >= 2 and ((h[-1][0] - h[-2][0]) * (q[1] - h[-2][1]) - (h[-1][1] - h[-2][1]) * (q[0] - h[-2][0])) <= 0:
                h.pop()
            h.append(q)
        return h
    lo, up = half(p), half(reversed(p))
    return np.array(lo[:-1] + up[:-1], dtype=np.float64)


def polygon_area(poly: np.ndarray) -> float:
    poly = np.asarray(poly, dtype=np.float64)
    if len(poly) < 3:
        return 0.0
    x, y = poly[:, 0], poly[:, 1]
    return float(abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) * 0.5)


def polygon_perimeter(poly: np.ndarray) -> float:
    poly = np.asarray(poly, dtype=np.float64)
    return float(np.linalg.norm(poly - np.roll(poly, -1, axis=0), axis=1).sum())


def hull_area(points) -> float:
    return polygon_area(_hull(points))


def _seg_point_dist(p, a, b):
    ab, ap = b - a, p - a
    den = float(ab @ ab)
    t = 0.0 if den == 0 else min(1.0, max(0.0, float(ap @ ab) / den))
    return float(np.linalg.norm(p - (a + t * ab)))


def _segments_cross(a, b, c, d):
    def orient(p, q, r):
        return (q[0] - p[0]) * (r[1] - p[1]) - (q[1] - p[1]) * (r[0] - p[0])
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    return (o1 * o2 < 0) and (o3 * o4 < 0)


def _inside(poly, p):
    return cv2.pointPolygonTest(poly.astype(np.float32).reshape(-1, 1, 2), (float(p[0]), float(p[1])), False) >= 0


def _orient(p, q, r):
    return (q[..., 0] - p[..., 0]) * (r[..., 1] - p[..., 1]) - (q[..., 1] - p[..., 1]) * (r[..., 0] - p[..., 0])


def _points_to_segments(P, s0, s1) -> float:
    """min over points P [n,2] and segments s0[j] -> s1[j] of the point-to-segment distance (all pairs at once)."""
    ab = (s1 - s0)[None]
    ap = P[:, None] - s0[None]
    den = (ab * ab).sum(-1)
    t = np.where(den == 0, 0.0, np.clip((ap * ab).sum(-1) / np.where(den == 0, 1.0, den), 0.0, 1.0))
    off = P[:, None] - (s0[None] + t[..., None] * ab)
    return float(np.sqrt((off * off).sum(-1)).min())


def polygon_distance(p1: np.ndarray, p2: np.ndarray) -> float:
    """shapely Polygon.distance for two simple polygons: 0 when any edges properly cross or one contains the other, else the
    smallest vertex-to-edge distance.  All edge pairs are evaluated in a few array operations (the text-direction graph calls this
    for every nearby pair of lines)."""
    p1, p2 = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    a, b = p1, np.roll(p1, -1, axis=0)
    c, d = p2, np.roll(p2, -1, axis=0)
    A, B, C, D = a[:, None], b[:, None], c[None], d[None]
    if ((_orient(A, B, C) * _orient(A, B, D) < 0) & (_orient(C, D, A) * _orient(C, D, B) < 0)).any():
        return 0.0
    if _inside(p1, p2[0]) or _inside(p2, p1[0]):
        return 0.0
    return min(_points_to_segments(p1, c, d), _points_to_segments(p2, a, b))


class Quadrilateral:
    """Text-line quad (mirror of generic.py:356-599, only what the detect/OCR path touches)."""

    def __init__(self, pts, text: str, prob: float, fg_r=0, fg_g=0, fg_b=0, bg_r=0, bg_g=0, bg_b=0):
        self.pts, vertical = order_points(pts)
        self.direction = "v" if vertical else "h"
        self.text, self.prob = text, prob
        self.fg_r, self.fg_g, self.fg_b = fg_r, fg_g, fg_b
        self.bg_r, self.bg_g, self.bg_b = bg_r, bg_g, bg_b
        self.assigned_direction = None
        self.textlines: List["Quadrilateral"] = []

    @functools.cached_property
    def structure(self):
        p = self.pts
        return [((p[0] + p[1]) / 2).astype(int), ((p[2] + p[3]) / 2).astype(int),
                ((p[1] + p[2]) / 2).astype(int), ((p[3] + p[0]) / 2).astype(int)]

    def _axes(self):
        a, b, c, d = [q.astype(np.float32) for q in self.structure]
        return b - a, d - c

    @functools.cached_property
    def aspect_ratio(self) -> float:
        v1, v2 = self._axes()
        return np.linalg.norm(v2) / np.linalg.norm(v1)

    @functools.cached_property
    def font_size(self) -> float:
        v1, v2 = self._axes()
        return min(np.linalg.norm(v2), np.linalg.norm(v1))

    @functools.cached_property
    def aabb(self) -> AABB:
        mx, mn = self.pts.max(axis=0), self.pts.min(axis=0)
        return AABB(mn[0], mn[1], mx[0] - mn[0], mx[1] - mn[1])

    @functools.cached_property
    def xyxy(self):
        b = self.aabb
        return b.x, b.y, b.x + b.w, b.y + b.h

    @functools.cached_property
    def is_approximate_axis_aligned(self) -> bool:
        v1, v2 = self._axes()
        u1, u2 = v1 / np.linalg.norm(v1), v2 / np.linalg.norm(v2)
        return bool(min(abs(u1[0]), abs(u1[1]), abs(u2[0]), abs(u2[1])) < 0.05)

    @functools.cached_property
    def angle(self) -> float:
        v1, _ = self._axes()
        return float(np.fmod(np.arccos((v1 / np.linalg.norm(v1))[0]) + np.pi, np.pi))

    @functools.cached_property
    def area(self) -> float:
        return hull_area(self.pts)

    def poly_distance(self, other) -> float:
        return polygon_distance(_hull(self.pts), _hull(other.pts))

    @functools.cached_property
    def centroid(self) -> np.ndarray:
        return np.average(self.pts, axis=0)

    def distance(self, other: "Quadrilateral", rho: float = 0.5) -> float:
        """Reading-order distance between two lines (generic.py:543-596): distance of the line starts (or ends / middles) when
        the quadrilateral spanned by the corresponding edges is thin relative to the font size, chosen per assigned direction."""
        fs = max(self.font_size, other.font_size)

        def d(p, q):
            return float(np.hypot(float(p[0]) - float(q[0]), float(p[1]) - float(q[1])))

        if self.assigned_direction == "h":
            d1 = hull_area([self.pts[0], self.pts[3], other.pts[0], other.pts[3]]) / fs
            d2 = hull_area([self.pts[2], self.pts[1], other.pts[2], other.pts[1]]) / fs
            d3 = hull_area([self.structure[0], self.structure[1], other.structure[0], other.structure[1]]) / fs
            pattern = "left"
            if d2 < fs * rho and d2 < d1:
                pattern = "right"
            if d3 < fs * rho and d3 < d1 and d3 < d2:
                pattern = "middle"
            if pattern == "left":
                return d(self.pts[0], other.pts[0])
            if pattern == "right":
                return d(self.pts[1], other.pts[1])
            return d(self.structure[0], other.structure[0])
        d1 = hull_area([self.pts[0], self.pts[1], other.pts[0], other.pts[1]]) / fs
        d2 = hull_area([self.pts[2], self.pts[3], other.pts[2], other.pts[3]]) / fs
        if d2 < fs * rho and d2 < d1:
            return d(self.pts[2], other.pts[2])
        return d(self.pts[0], other.pts[0])

    def clip(self, width, height):
        self.pts[:, 0] = np.clip(np.round(self.pts[:, 0]), 0, width)
        self.pts[:, 1] = np.clip(np.round(self.pts[:, 1]), 0, height)

    def get_transformed_region(self, img, direction, textheight) -> np.ndarray:
        """Perspective crop to a `textheight`-tall strip (generic.py:445-481): crop the AABB, homography from the 4 corners
        (cv2.findHomography RANSAC 5.0), warpPerspective; vertical lines are rotated 90 degrees counter-clockwise."""
        (x1, y1, x2, y2), M, (w, h) = warp_setup(self, img.shape[0], img.shape[1], direction, textheight)
        region = cv2.warpPerspective(img[y1:y2, x1:x2], M, (w, h))
        if direction == "v":
            region = cv2.rotate(region, cv2.ROTATE_90_COUNTERCLOCKWISE)
        return region


def warp_setup(q, im_h: int, im_w: int, direction: str, textheight):
    """Host half of Quadrilateral.get_transformed_region (generic.py:445-468): the clipped AABB of the quad, the homography onto
    the (w, h) strip and that size.  `q` is any object with `.structure` and `.pts` (ours or the reference's Quadrilateral)."""
    l1a, l1b, l2a, l2b = [np.asarray(a).astype(np.float32) for a in q.structure]
    ratio = np.linalg.norm(l1b - l1a) / np.linalg.norm(l2b - l2a)
    src = q.pts.astype(np.int64).copy()
    x1, y1 = np.clip(src[:, 0].min(), 0, im_w), np.clip(src[:, 1].min(), 0, im_h)
    x2, y2 = np.clip(src[:, 0].max(), 0, im_w), np.clip(src[:, 1].max(), 0, im_h)
    src[:, 0] -= x1
    src[:, 1] -= y1
    q.assigned_direction = direction
    if direction == "h":
        h, w = max(int(textheight), 2), max(int(round(textheight / ratio)), 2)
    else:
        w, h = max(int(textheight), 2), max(int(round(textheight * ratio)), 2)
    dst = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float32)
    M, _ = cv2.findHomography(src, dst, cv2.RANSAC, 5.0)
    return (int(x1), int(y1), int(x2), int(y2)), M, (w, h)


def warp_record(q, im_h: int, im_w: int, direction: str, textheight=48):
    """One line record of `mitb_op_warp_lines_u8` (include/mitb.h): float64[16] = inverse homography (cv2.invert, the very call
    cv2.warpPerspective makes on M), crop origin and size, strip size before the rotation, rotation flag; plus the width of the
    strip as it lands in the OCR canvas.  A degenerate quad (empty crop, or no homography) yields an all-zero strip, where the
    reference would raise inside cv2."""
    (x1, y1, x2, y2), M, (w, h) = warp_setup(q, im_h, im_w, direction, textheight)
    rec = np.zeros(16, dtype=np.float64)
    rot = 1 if direction == "v" else 0
    if M is not None and x2 > x1 and y2 > y1:
        rec[:9] = cv2.invert(np.asarray(M, dtype=np.float64))[1].reshape(-1)
        rec[9:16] = (x1, y1, x2 - x1, y2 - y1, w, h, rot)
    else:
        rec[13:16] = (w, h, rot)
    return rec, (h if rot else w)


def can_merge_region(a: Quadrilateral, b: Quadrilateral, ratio=1.9, discard_connection_gap=2, char_gap_tolerance=0.6,
                     char_gap_tolerance2=1.5, font_size_ratio_tol=1.5, aspect_ratio_tol=2) -> bool:
    """Whether two text lines belong to one region (generic.py:653-698)."""
    (x1, y1, w1, h1), (x2, y2, w2, h2) = a.aabb, b.aabb
    char_size = min(a.font_size, b.font_size)
    # exact early-out: the AABB gap is a lower bound of the polygon distance
    gx = max(0, max(x1, x2) - min(x1 + w1, x2 + w2))
    gy = max(0, max(y1, y2) - min(y1 + h1, y2 + h2))
    if (gx * gx + gy * gy) ** 0.5 > discard_connection_gap * char_size:
        return False
    dist = polygon_distance(a.pts, b.pts)
    if dist > discard_connection_gap * char_size:
        return False
    if max(a.font_size, b.font_size) / char_size > font_size_ratio_tol:
        return False
    if a.aspect_ratio > aspect_ratio_tol and b.aspect_ratio < 1.0 / aspect_ratio_tol:
        return False
    if b.aspect_ratio > aspect_ratio_tol and a.aspect_ratio < 1.0 / aspect_ratio_tol:
        return False
    if a.is_approximate_axis_aligned and b.is_approximate_axis_aligned:
        if dist >= char_size * char_gap_tolerance:
            return False
        if abs(x1 + w1 // 2 - (x2 + w2 // 2)) < char_gap_tolerance2:
            return True
        if (w1 > h1 * ratio and h2 > w2 * ratio) or (w2 > h2 * ratio and h1 > w1 * ratio):
            return False
        lim = char_size * char_gap_tolerance2
        if w1 > h1 * ratio or w2 > h2 * ratio:
            return bool(abs(x1 - x2) < lim or abs(x1 + w1 - (x2 + w2)) < lim)
        if h1 > w1 * ratio or h2 > w2 * ratio:
            return bool(abs(y1 - y2) < lim or abs(y1 + h1 - (y2 + h2)) < lim)
        return False
    if abs(a.angle - b.angle) < 15 * np.pi / 180:
        fs = min(a.font_size, b.font_size)
        if a.poly_distance(b) > fs * char_gap_tolerance2:
            return False
        return bool(abs(a.font_size - b.font_size) / fs <= 0.25)
    return False


def pair_features(bboxes) -> np.ndarray:
    """Per-line record of `mitb_op_textline_pairs` (include/mitb.h): float64 [n,16] = corners, AABB, font_size, aspect_ratio, angle, flags
    (bit 0: approximately axis aligned, bit 1: strictly convex - the device evaluates polygon distances on the quad itself, which
    equals the convex hull the reference uses only then)."""
    f = np.zeros((len(bboxes), 16), dtype=np.float64)
    for i, b in enumerate(bboxes):
        p = np.asarray(b.pts, dtype=np.float64)
        f[i, :8] = p.reshape(-1)
        a = b.aabb
        f[i, 8:12] = (a.x, a.y, a.w, a.h)
        e = np.roll(p, -1, axis=0) - p
        en = np.roll(e, -1, axis=0)
        turn = e[:, 0] * en[:, 1] - e[:, 1] * en[:, 0]
        convex = bool((turn > 0).all() or (turn < 0).all())
        f[i, 12:16] = (b.font_size, b.aspect_ratio, b.angle, (1 if b.is_approximate_axis_aligned else 0) | (2 if convex else 0))
    return f


def can_merge_matrix(bboxes, engine, ratio=1.9, discard_connection_gap=2, char_gap_tolerance=0.6, char_gap_tolerance2=1.5,
                     font_size_ratio_tol=1.5, aspect_ratio_tol=2) -> np.ndarray:
    """`can_merge_region` for all pairs at once on the device; pairs the kernel leaves undecided (a non-convex quad)
    are evaluated here.  Returns a symmetric bool matrix."""
    n = len(bboxes)
    if n < 2:
        return np.zeros((n, n), dtype=bool)
    params = (ratio, discard_connection_gap, char_gap_tolerance, char_gap_tolerance2, font_size_ratio_tol, aspect_ratio_tol)
    adj = engine.textline_pairs(pair_features(bboxes), params)
    for u, v in np.argwhere(np.triu(adj == 2, 1)):
        r = can_merge_region(bboxes[int(u)], bboxes[int(v)], *params)
        adj[u, v] = adj[v, u] = 1 if r else 0
    return adj == 1


def generate_text_direction(bboxes: List[Quadrilateral], engine=None):
    """Yield (quad, majority direction of its merge-group), groups ordered reading-wise (ocr/common.py:12-39).  With `engine` the
    pair predicate runs on the device (`mitb_op_textline_pairs`)."""
    if not bboxes:
        return
    import networkx as nx
    G = nx.Graph()
    for i, box in enumerate(bboxes):
        G.add_node(i, box=box)
    if engine is not None:
        for u, v in np.argwhere(np.triu(can_merge_matrix(bboxes, engine, aspect_ratio_tol=1), 1)):
            G.add_edge(int(u), int(v))
        yield from _ordered_groups(G, bboxes)
        return
    # The reference tests all n(n-1)/2 pairs in Python; every pair whose AABB gap already exceeds the connection distance is
    # rejected by can_merge_region's first test, so a vectorised (slightly conservative) version of that test prunes the pair list
    # first - same edges, a fraction of the interpreter time on pages with many lines.
    xy = np.array([b.xyxy for b in bboxes], dtype=np.float64)
    fs = np.array([b.font_size for b in bboxes], dtype=np.float64)
    gx = np.maximum(0.0, np.maximum(xy[:, None, 0], xy[None, :, 0]) - np.minimum(xy[:, None, 2], xy[None, :, 2]))
    gy = np.maximum(0.0, np.maximum(xy[:, None, 1], xy[None, :, 1]) - np.minimum(xy[:, None, 3], xy[None, :, 3]))
    near = np.hypot(gx, gy) <= 2.0 * np.minimum(fs[:, None], fs[None, :]) * (1.0 + 1e-9) + 1e-9      # discard_connection_gap = 2
    for u, v in np.argwhere(np.triu(near, 1)):
        if can_merge_region(bboxes[int(u)], bboxes[int(v)], aspect_ratio_tol=1):
            G.add_edge(int(u), int(v))
    yield from _ordered_groups(G, bboxes)


def _ordered_groups(G, bboxes):
    import networkx as nx
    for comp in nx.algorithms.components.connected_components(G):
        nodes = list(comp)
        major = Counter(bboxes[i].direction for i in nodes).most_common(1)[0][0]
        if major == "h":
            nodes = sorted(nodes, key=lambda i: bboxes[i].aabb.y + bboxes[i].aabb.h // 2)
        elif major == "v":
            nodes = sorted(nodes, key=lambda i: -(bboxes[i].aabb.x + bboxes[i].aabb.w))
        for i in nodes:
            yield bboxes[i], major

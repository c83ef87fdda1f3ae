"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's mask refinement,
`manga_translator/mask_refinement/__init__.py:9-50` (dispatch) and `text_mask_utils.py:64-195` (refine_mask, complete_mask), in the
reference's own statement order, with cv2 for the calls the reference makes to cv2.

Two third-party dependencies of that code are absent from this image and from /root/reference, and are restated here:

* **shapely** (unpinned in requirements.txt): `Polygon.area`, `Polygon.intersection(rect).area`, `Polygon.distance(Point)` for the
  text-line quads against a connected component's bounding rectangle -> shoelace area, Sutherland-Hodgman clipping against the
  axis-aligned rectangle, point-to-polygon distance.  Pinned by hand-derived cases in tests/test_mask_refine.py.
* **pydensecrf** (git dependency `lucasb-eyer/pydensecrf`, unpinned; it wraps Philipp Kraehenbuehl's densecrf): `DenseCRF2D`,
  `setUnaryEnergy`, `addPairwiseGaussian(sxy=1, compat=3)`, `addPairwiseBilateral(sxy=23, srgb=7, compat=20)` (DIAG_KERNEL,
  NO_NORMALIZATION), `inference(5)`, and `utils.unary_from_softmax`.  Restated from the published algorithm (Adams, Baek, Davis:
  "Fast high-dimensional filtering using the permutohedral lattice", 2010; densecrf `permutohedral.cpp`, `pairwise.cpp`,
  `densecrf.cpp`): elevate -> round to the nearest remainder-0 point -> rank -> barycentric weights -> splat, blur along the
  d+1 axes with (1/2, 1, 1/2), slice with alpha = 1 / (1 + 2^-d); Potts compatibility; mean field Q <- softmax(-U + sum_k w_k K_k Q).
  **Parity unpinned**: no run of the real library is possible here, so tests can only check this restatement against closed-form
  properties and the CUDA path against this restatement.

Pinned part: tests/test_mask_refine.py::test_oracle_dispatch_equals_reference_code_with_restated_dependencies runs the reference's own,
unmodified mask_refinement package (from /root/reference) with shapely / pydensecrf bound to the restatements below and requires
`dispatch` here to return the same mask bit for bit - so everything outside those two libraries is pinned on the reference's code.
"""
import math

import cv2
import numpy as np

# ------------------------------------------------------------------------------------------------ shapely stand-ins


def poly_area(pts) -> float:
    p = np.asarray(pts, dtype=np.float64)
    x, y = p[:, 0], p[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def clip_poly_rect(pts, x0, y0, x1, y1):
    """Sutherland-Hodgman: subject polygon clipped by the axis-aligned rectangle [x0,x1] x [y0,y1]."""
    out = [tuple(map(float, p)) for p in pts]
    for axis, bound, keep_less in ((0, x0, False), (0, x1, True), (1, y0, False), (1, y1, True)):
        if not out:
            break
        inp, out = out, []
        for i in range(len(inp)):
            a, b = inp[i - 1], inp[i]
            ina = a[axis] <= bound if keep_less else a[axis] >= bound
            inb = b[axis] <= bound if keep_less else b[axis] >= bound
            if ina != inb:
                t = (bound - a[axis]) / (b[axis] - a[axis])
                out.append((a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1])))
            if inb:
                out.append(b)
    return out


def intersection_area_rect(pts, x0, y0, x1, y1) -> float:
    c = clip_poly_rect(pts, x0, y0, x1, y1)
    return poly_area(c) if len(c) >= 3 else 0.0


def point_in_poly(pts, px, py) -> bool:
    p = np.asarray(pts, dtype=np.float64)
    inside = False
    for i in range(len(p)):
        (xa, ya), (xb, yb) = p[i - 1], p[i]
        if (ya > py) != (yb > py) and px < (xb - xa) * (py - ya) / (yb - ya) + xa:
            inside = not inside
    return inside


def point_poly_distance(pts, px, py) -> float:
    """shapely Polygon.distance(Point): 0 inside (or on the boundary), else the distance to the nearest edge."""
    if point_in_poly(pts, px, py):
        return 0.0
    p = np.asarray(pts, dtype=np.float64)
    best = math.inf
    for i in range(len(p)):
        a, b = p[i - 1], p[i]
        ab = b - a
        den = float(ab @ ab)
        t = 0.0 if den == 0 else min(1.0, max(0.0, float((np.array([px, py]) - a) @ ab) / den))
        q = a + t * ab
        best = min(best, math.hypot(px - q[0], py - q[1]))
    return best


# ------------------------------------------------------------------------------------------------ densecrf restatement


class Permutohedral:
    """densecrf permutohedral.cpp: Permutohedral::init / seqCompute (float32 arithmetic, sequential accumulation order)."""

    def __init__(self, feature: np.ndarray):
        f = np.ascontiguousarray(feature, dtype=np.float32)
        d, N = f.shape
        self.d, self.N = d, N
        f32 = np.float32
        inv_std_dev = f32(math.sqrt(2.0 / 3.0) * (d + 1))
        scale = np.array([f32(1.0 / math.sqrt(float((i + 2) * (i + 1))) * float(inv_std_dev)) for i in range(d)], dtype=f32)
        # elevate: y = E p
        elevated = np.zeros((d + 1, N), f32)
        sm = np.zeros(N, f32)
        for j in range(d, 0, -1):
            cf = f[j - 1] * scale[j - 1]
            elevated[j] = sm - f32(j) * cf
            sm = sm + cf
        elevated[0] = sm
        # nearest remainder-0 lattice point
        down, up = f32(1.0 / (d + 1)), f32(d + 1)
        rem0 = np.zeros((d + 1, N), f32)
        ssum = np.zeros(N, np.int32)
        for i in range(d + 1):
            v = down * elevated[i]
            upv, dnv = np.ceil(v) * up, np.floor(v) * up
            rd = np.where(upv - elevated[i] < elevated[i] - dnv, upv, dnv).astype(np.int16).astype(f32)
            rem0[i] = rd
            ssum = (ssum.astype(f32) + rd * down).astype(np.int32)               # `int sum; sum += rd2*down_factor;`
        # rank of each coordinate's remainder
        rank = np.zeros((d + 1, N), np.int32)
        diff = elevated - rem0
        for i in range(d):
            for j in range(i + 1, d + 1):
                lt = diff[i] < diff[j]
                rank[i] += lt
                rank[j] += ~lt
        rank += ssum[None]
        lo, hi = rank < 0, rank > d
        rank = np.where(lo, rank + d + 1, np.where(hi, rank - (d + 1), rank))
        rem0 = np.where(lo, rem0 + f32(d + 1), np.where(hi, rem0 - f32(d + 1), rem0))
        # barycentric coordinates
        bary = np.zeros((d + 2, N), f32)
        cols = np.arange(N)
        for i in range(d + 1):
            v = (elevated[i] - rem0[i]) * down
            bary[d - rank[i], cols] += v
            bary[d - rank[i] + 1, cols] -= v
        bary[0] = (1.0 + bary[d + 1].astype(np.float64) + bary[0].astype(np.float64)).astype(f32)
        # simplex vertices -> hash table
        canonical = np.zeros((d + 1, d + 1), np.int32)
        for i in range(d + 1):
            canonical[i, :d - i + 1] = i
            canonical[i, d - i + 1:] = i - (d + 1)
        rem_i = rem0.astype(np.int32)
        keys = np.zeros((N, d + 1, d), np.int32)
        for r in range(d + 1):
            keys[:, r, :] = (rem_i[:d] + canonical[r][rank[:d]]).T
        uniq, inv = np.unique(keys.reshape(-1, d), axis=0, return_inverse=True)
        self.offset = inv.reshape(N, d + 1)
        self.bary = np.ascontiguousarray(bary[:d + 1].T)
        self.M = M = len(uniq)
        table = {tuple(k): i for i, k in enumerate(uniq.tolist())}
        self.n1 = np.full((d + 1, M), -1, np.int64)
        self.n2 = np.full((d + 1, M), -1, np.int64)
        for j in range(d + 1):
            a, b = uniq - 1, uniq + 1
            if j < d:
                a = a.copy(); b = b.copy()
                a[:, j] = uniq[:, j] + d
                b[:, j] = uniq[:, j] - d
            self.n1[j] = [table.get(tuple(k), -1) for k in a.tolist()]
            self.n2[j] = [table.get(tuple(k), -1) for k in b.tolist()]

    def compute(self, inp: np.ndarray) -> np.ndarray:
        """inp float32 [value_size, N] -> filtered [value_size, N]"""
        f32 = np.float32
        d, N, M = self.d, self.N, self.M
        vs = inp.shape[0]
        values = np.zeros((M + 2, vs), f32)
        x = np.ascontiguousarray(inp.T, dtype=f32)
        o = (self.offset + 1).reshape(-1)
        w = self.bary.reshape(-1)
        np.add.at(values, o, w[:, None] * np.repeat(x, d + 1, axis=0))
        for j in range(d + 1):
            n1, n2 = self.n1[j] + 1, self.n2[j] + 1
            new = np.zeros_like(values)
            new[1:M + 1] = values[1:M + 1] + f32(0.5) * (values[n1] + values[n2])
            values = new
        alpha = f32(1.0 / (1.0 + 2.0 ** (-d)))
        out = np.zeros((N, vs), f32)
        for j in range(d + 1):
            out += self.bary[:, j:j + 1] * values[self.offset[:, j] + 1] * alpha
        return np.ascontiguousarray(out.T)


def dense_crf_2d(rgb: np.ndarray, unary: np.ndarray, n_iter: int = 5, sxy_g: float = 1.0, compat_g: float = 3.0, sxy_b: float = 23.0,
                 srgb: float = 7.0, compat_b: float = 20.0) -> np.ndarray:
    """DenseCRF2D(W, H, L) + setUnaryEnergy + addPairwiseGaussian + addPairwiseBilateral + inference(n_iter): returns Q [L, H*W]."""
    H, W = rgb.shape[:2]
    f32 = np.float32
    ys, xs = np.mgrid[0:H, 0:W]
    xs, ys = xs.reshape(-1).astype(f32), ys.reshape(-1).astype(f32)
    lat_g = Permutohedral(np.stack([xs / f32(sxy_g), ys / f32(sxy_g)]))
    c = rgb.reshape(-1, 3).astype(f32)
    lat_b = Permutohedral(np.stack([xs / f32(sxy_b), ys / f32(sxy_b), c[:, 0] / f32(srgb), c[:, 1] / f32(srgb), c[:, 2] / f32(srgb)]))
    unary = np.asarray(unary, dtype=f32)

    def exp_and_normalize(v):
        e = np.exp(v - v.max(axis=0, keepdims=True))
        return (e / e.sum(axis=0, keepdims=True)).astype(f32)

    Q = exp_and_normalize(-unary)
    for _ in range(n_iter):
        tmp = -unary
        tmp = tmp - (-f32(compat_g) * lat_g.compute(Q))
        tmp = tmp - (-f32(compat_b) * lat_b.compute(Q))
        Q = exp_and_normalize(tmp)
    return Q


def refine_mask(rgbimg: np.ndarray, rawmask: np.ndarray) -> np.ndarray:
    """text_mask_utils.py:71-94"""
    m = rawmask.reshape(rawmask.shape[0], rawmask.shape[1])
    sm = np.stack([cv2.bitwise_not(m), m]).astype(np.float32) / 255.0
    unary = (-np.log(np.clip(sm.reshape(2, -1), 1e-5, 1.0))).astype(np.float32)           # pydensecrf.utils.unary_from_softmax
    Q = dense_crf_2d(np.ascontiguousarray(rgbimg), unary, 5)
    res = np.argmax(Q, axis=0).reshape(rgbimg.shape[0], rgbimg.shape[1])
    return np.array(res * 255, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------ complete_mask / dispatch


def extend_rect(x, y, w, h, max_x, max_y, extend_size):
    x1 = max(x - extend_size, 0)
    y1 = max(y - extend_size, 0)
    w1 = min(w + extend_size * 2, max_x - x1 - 1)
    h1 = min(h + extend_size * 2, max_y - y1 - 1)
    return x1, y1, w1, h1


def assign_components(stats: np.ndarray, textlines, keep_threshold=1e-2):
    """The per-label decision of complete_mask (text_mask_utils.py:110-160): label -> text-line index or -1."""
    M = len(textlines)
    polys = [np.asarray(t.pts, dtype=np.float64) for t in textlines]
    areas = [poly_area(p) for p in polys]
    out = np.full(len(stats), -1, np.int64)
    for label in range(1, len(stats)):
        x1, y1, w1, h1, area1 = [int(v) for v in stats[label]]
        if area1 <= 9:
            continue
        ratio = np.zeros(M, np.float32)
        dist = np.zeros(M, np.float32)
        for t in range(M):
            ov = intersection_area_rect(polys[t], x1, y1, x1 + w1, y1 + h1)
            ratio[t] = ov / min(area1, areas[t])
            dist[t] = point_poly_distance(polys[t], x1 + w1 / 2.0, y1 + h1 / 2.0)
        avg = int(np.argmax(ratio))
        if area1 >= areas[avg]:
            continue
        if ratio[avg] <= keep_threshold:
            avg = int(np.argmin(dist))
            unit = max(min([textlines[avg].font_size, w1, h1]), 10)
            if dist[avg] >= 0.5 * unit:
                continue
        out[label] = avg
    return out


def complete_mask(img, mask, textlines, keep_threshold=1e-2, dilation_offset=0, kernel_size=3, refine=refine_mask):
    """text_mask_utils.py:96-190 (mask is modified in place, like the reference does)."""
    bboxes = [t.aabb_xywh for t in textlines]
    for (x, y, w, h) in bboxes:
        cv2.rectangle(mask, (int(x), int(y)), (int(x + w), int(y + h)), (0), 1)
    num_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask)
    M = len(textlines)
    textline_ccs = [np.zeros_like(mask) for _ in range(M)]
    iinfo = np.iinfo(labels.dtype)
    rects = np.full((M, 4), [iinfo.max, iinfo.max, iinfo.min, iinfo.min], dtype=labels.dtype)
    owner = assign_components(stats, textlines, keep_threshold)
    valid = False
    for label in range(1, num_labels):
        avg = owner[label]
        if avg < 0:
            continue
        x1, y1, w1, h1 = [int(v) for v in stats[label, :4]]
        textline_ccs[avg][y1:y1 + h1, x1:x1 + w1][labels[y1:y1 + h1, x1:x1 + w1] == label] = 255
        rects[avg, 0] = min(rects[avg, 0], x1)
        rects[avg, 1] = min(rects[avg, 1], y1)
        rects[avg, 2] = max(rects[avg, 2], x1 + w1)
        rects[avg, 3] = max(rects[avg, 3], y1 + h1)
        valid = True
    if not valid:
        return None
    rects[:, 2] -= rects[:, 0]
    rects[:, 3] -= rects[:, 1]
    final_mask = np.zeros_like(mask)
    img = cv2.bilateralFilter(img, 17, 80, 80)
    for i, cc in enumerate(textline_ccs):
        x1, y1, w1, h1 = [int(v) for v in rects[i]]
        text_size = min(w1, h1, textlines[i].font_size)
        x1, y1, w1, h1 = extend_rect(x1, y1, w1, h1, img.shape[1], img.shape[0], int(text_size * 0.1))
        dilate_size = max((int((text_size + dilation_offset) * 0.3) // 2) * 2 + 1, 3)
        kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (dilate_size, dilate_size))
        cc_region = np.ascontiguousarray(cc[y1:y1 + h1, x1:x1 + w1])
        if cc_region.size == 0:
            continue
        img_region = np.ascontiguousarray(img[y1:y1 + h1, x1:x1 + w1])
        cc[y1:y1 + h1, x1:x1 + w1] = refine(img_region, cc_region)
        x2, y2, w2, h2 = extend_rect(x1, y1, w1, h1, img.shape[1], img.shape[0], -(-dilate_size // 2))
        cc[y2:y2 + h2, x2:x2 + w2] = cv2.dilate(cc[y2:y2 + h2, x2:x2 + w2], kern)
        final_mask[y2:y2 + h2, x2:x2 + w2] = cv2.bitwise_or(final_mask[y2:y2 + h2, x2:x2 + w2], cc[y2:y2 + h2, x2:x2 + w2])
    kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (kernel_size, kernel_size))
    return cv2.dilate(final_mask, kern)


class _Line:
    """What complete_mask reads of a reference Quadrilateral built from `line * scale_factor` (generic.py:356-417, 432-444)."""

    def __init__(self, quad_cls, pts):
        q = quad_cls(pts, "", 0)
        self.pts = q.pts
        self.font_size = float(q.font_size)
        a = q.aabb
        self.aabb_xywh = np.array([a.x, a.y, a.w, a.h] if hasattr(a, "x") else list(a), dtype=np.int32)


def dispatch(text_regions, raw_image: np.ndarray, raw_mask: np.ndarray, quad_cls, method: str = "fit_text", dilation_offset: int = 0,
             ignore_bubble: int = 0, kernel_size: int = 3, refine=refine_mask) -> np.ndarray:
    """mask_refinement/__init__.py:9-31 (the `ignore_bubble` tail, :33-50, is outside the default configuration and not restated).
    `text_regions`: objects with `.lines` (float arrays [n,4,2]); `quad_cls`: the Quadrilateral class to build the scaled lines with."""
    assert method == "fit_text" and not (1 <= ignore_bubble <= 50)
    scale_factor = max(min((raw_mask.shape[0] - raw_image.shape[0] / 3) / raw_mask.shape[0], 1), 0.5)
    size = (int(raw_image.shape[1] * scale_factor), int(raw_image.shape[0] * scale_factor))
    img_resized = cv2.resize(raw_image, size, interpolation=cv2.INTER_LINEAR)
    mask_resized = cv2.resize(raw_mask, size, interpolation=cv2.INTER_LINEAR)
    mask_resized[mask_resized > 0] = 255
    textlines = [_Line(quad_cls, l * scale_factor) for region in text_regions for l in region.lines]
    final_mask = complete_mask(img_resized, mask_resized, textlines, dilation_offset=dilation_offset, kernel_size=kernel_size, refine=refine)
    if final_mask is None:
        return np.zeros((raw_image.shape[0], raw_image.shape[1]), dtype=np.uint8)
    final_mask = cv2.resize(final_mask, (raw_image.shape[1], raw_image.shape[0]), interpolation=cv2.INTER_LINEAR)
    final_mask[final_mask > 0] = 255
    return final_mask

"""Mask refinement on the GPU: drop-in for `manga_translator.mask_refinement.dispatch`
(mask_refinement/__init__.py:9-50; complete_mask / refine_mask in text_mask_utils.py:64-190), the CPU stage the reference runs
between OCR and inpainting.

Same signature and result type as the reference (`async def dispatch(text_regions, raw_image, raw_mask, method, dilation_offset,
ignore_bubble, verbose, kernel_size) -> uint8 [H,W]`).  Page and raw mask go to the device once; resize, rectangle cuts, connected
components, the 17-px bilateral filter, the per-line DenseCRF (all lines of the page batched into one set of launches), the ellipse
dilations and the resize back are kernels of libmitb.so (csrc/maskrefine.cu, csrc/bilateral.cu).  The host keeps what is small and
geometric: which text line owns which connected component (a few hundred rectangle-vs-quad tests per page).

Not covered: `method != 'fit_text'` (the reference's `complete_mask_fill` reads an undefined variable, text_mask_utils.py:59-62) and
the `ignore_bubble` tail (:33-50), which is off by default; both raise here instead of silently doing something else.
There is no CPU fallback: without the CUDA library this module raises.
"""
from __future__ import annotations

import ctypes
from typing import List

import cv2
import numpy as np
import torch

from ._lib import MitbError
from .engine import Engine, _ptr, get_engine
from .host.geometry import Quadrilateral

# DenseCRF parameters of refine_mask (text_mask_utils.py:82-91)
CRF_ITERS, SXY_G, W_G, SXY_B, SRGB, W_B = 5, 1.0, 3.0, 23.0, 7.0, 20.0
U_ON = float(-np.log(np.clip(np.float32(0.0), 1e-5, 1.0).astype(np.float32)))       # unary_from_softmax of probability 0 (clip 1e-5)
CC_CAP = 1 << 18


# ------------------------------------------------------------------------------------------------ host geometry (shapely stand-ins)
def _poly_area(p: np.ndarray) -> float:
    x, y = p[:, 0], p[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def _clip_halfplane(poly: np.ndarray, axis: int, bound: float, keep_less: bool) -> np.ndarray:
    """One Sutherland-Hodgman step, vectorised over the polygon's edges."""
    if len(poly) == 0:
        return poly
    prev = np.roll(poly, 1, axis=0)
    inside_c = poly[:, axis] <= bound if keep_less else poly[:, axis] >= bound
    inside_p = prev[:, axis] <= bound if keep_less else prev[:, axis] >= bound
    out = []
    for a, b, ia, ib in zip(prev, poly, inside_p, inside_c):
        if ia != ib:
            t = (bound - a[axis]) / (b[axis] - a[axis])
            out.append(a + t * (b - a))
        if ib:
            out.append(b)
    return np.asarray(out, dtype=np.float64).reshape(-1, 2)


def _overlap_area(poly: np.ndarray, x0: float, y0: float, x1: float, y1: float) -> float:
    """Area of polygon ∩ axis-aligned rectangle (shapely: polys[i].intersection(cc_poly).area)."""
    c = poly
    for axis, bound, less in ((0, x0, False), (0, x1, True), (1, y0, False), (1, y1, True)):
        c = _clip_halfplane(c, axis, bound, less)
        if len(c) < 3:
            return 0.0
    return _poly_area(c)


def _point_distance(poly: np.ndarray, px: float, py: float) -> float:
    """shapely: Polygon.distance(Point) - 0 inside, else distance to the boundary."""
    x, y = poly[:, 0], poly[:, 1]
    xp, yp = np.roll(x, 1), np.roll(y, 1)
    cross = (yp > py) != (y > py)
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = (x - xp) * (py - yp) / (y - yp) + xp
    if (np.count_nonzero(cross & (px < xi)) & 1) == 1:
        return 0.0
    a = np.stack([xp, yp], 1)
    ab = poly - a
    den = (ab * ab).sum(1)
    t = np.where(den > 0, ((np.array([px, py]) - a) * ab).sum(1) / np.where(den > 0, den, 1.0), 0.0).clip(0.0, 1.0)
    q = a + t[:, None] * ab
    return float(np.sqrt(((q - np.array([px, py])) ** 2).sum(1)).min())


def assign_components(stats: np.ndarray, polys: List[np.ndarray], font_sizes: List[float], keep_threshold: float = 1e-2) -> np.ndarray:
    """complete_mask's decision per component (text_mask_utils.py:110-160): stats int [n][5] = x0, y0, x1, y1 (inclusive), area ->
    owner line per component (-1: dropped).  Vectorised over (component, line) pairs: a component rectangle that lies inside a
    (convex) line quad overlaps it by exactly its own area, a rectangle that misses the quad's bounding box by exactly 0; only
    the partially overlapping pairs are clipped (Sutherland-Hodgman), and distances are evaluated only for components that no
    line overlaps enough - a few per page."""
    M = len(polys)
    owner = np.full(len(stats), -1, np.int32)
    if M == 0 or len(stats) == 0:
        return owner
    st = np.asarray(stats, dtype=np.int64)
    big = np.nonzero(st[:, 4] > 9)[0]
    if len(big) == 0:
        return owner
    x0, y0 = st[big, 0].astype(np.float64), st[big, 1].astype(np.float64)
    w1, h1 = (st[big, 2] - st[big, 0] + 1).astype(np.float64), (st[big, 3] - st[big, 1] + 1).astype(np.float64)
    x1, y1, area1 = x0 + w1, y0 + h1, st[big, 4].astype(np.float64)      # cc_pts = (x, y) .. (x + w, y + h) in the reference
    P = np.stack(polys).astype(np.float64)                                  # [M,4,2]
    areas = np.array([_poly_area(p) for p in polys])
    pmin, pmax = P.min(1), P.max(1)
    touch = (pmin[None, :, 0] < x1[:, None]) & (pmax[None, :, 0] > x0[:, None]) & (pmin[None, :, 1] < y1[:, None]) & (pmax[None, :, 1] > y0[:, None])
    E = np.roll(P, -1, axis=1) - P                                          # [M,4 edges,2]
    En = np.roll(E, -1, axis=1)
    turn = E[:, :, 0] * En[:, :, 1] - E[:, :, 1] * En[:, :, 0]
    convex = (turn >= 0).all(1) | (turn <= 0).all(1)
    kk, tt = np.nonzero(touch)                                              # the (component, line) pairs whose boxes intersect
    corners = np.stack([np.stack([x0, y0], 1), np.stack([x1, y0], 1), np.stack([x1, y1], 1), np.stack([x0, y1], 1)], 1)[kk]     # [p,4,2]
    rel = corners[:, :, None, :] - P[tt][:, None, :, :]                     # [p,4 corners,4 edges,2]
    cross = E[tt][:, None, :, 0] * rel[..., 1] - E[tt][:, None, :, 1] * rel[..., 0]
    inside = ((cross >= 0).all(axis=(1, 2)) | (cross <= 0).all(axis=(1, 2))) & convex[tt]      # every corner on the inner side of every edge
    ratio = np.zeros((len(big), M), np.float32)
    denom = np.minimum(area1[:, None], areas[None, :])
    ratio[kk[inside], tt[inside]] = ((w1 * h1)[kk[inside]] / denom[kk[inside], tt[inside]])
    for k, t in zip(kk[~inside], tt[~inside]):
        ratio[k, t] = _overlap_area(polys[t], x0[k], y0[k], x1[k], y1[k]) / denom[k, t]
    avg = ratio.argmax(1)
    keep = area1 < areas[avg]
    weak = keep & (ratio[np.arange(len(big)), avg] <= keep_threshold)
    for k in np.nonzero(weak)[0]:
        cx, cy = x0[k] + w1[k] / 2.0, y0[k] + h1[k] / 2.0
        dist = np.array([_point_distance(p, cx, cy) for p in polys], dtype=np.float32)
        a = int(np.argmin(dist))
        unit = max(min([font_sizes[a], w1[k], h1[k]]), 10)
        if dist[a] >= 0.5 * unit:
            keep[k] = False
        avg[k] = a
    owner[big[keep]] = avg[keep]
    return owner


def _extend_rect(x, y, w, h, max_x, max_y, extend_size):
    x1 = max(x - extend_size, 0)
    y1 = max(y - extend_size, 0)
    return x1, y1, min(w + extend_size * 2, max_x - x1 - 1), min(h + extend_size * 2, max_y - y1 - 1)


def _pow2_at_least(n: int) -> int:
    return 1 << max(4, int(n - 1).bit_length())


# ------------------------------------------------------------------------------------------------ device steps
class MaskRefiner:
    """Device-side implementation bound to one Engine (one GPU)."""

    def __init__(self, engine: Engine):
        self.eng = engine
        self.lib = engine.lib
        self._se_cache = {}

    def _dev(self, a: np.ndarray) -> torch.Tensor:
        return self.eng.h2d(np.ascontiguousarray(a))

    def resize(self, src: torch.Tensor, dw: int, dh: int, binarize: bool = False) -> torch.Tensor:
        cn = 1 if src.dim() == 2 else int(src.shape[2])
        dst = torch.empty((dh, dw) if cn == 1 else (dh, dw, cn), dtype=torch.uint8, device=src.device)
        self.eng._call(self.lib.mitb_op_resize_linear_u8, _ptr(src), int(src.shape[0]), int(src.shape[1]), cn, _ptr(dst), dh, dw, int(binarize),
                       self.eng._stream())
        return dst

    def components(self, mask: torch.Tensor):
        h, w = int(mask.shape[0]), int(mask.shape[1])
        dev = mask.device
        labels = torch.empty((h * w,), dtype=torch.int32, device=dev)
        stats = torch.empty((CC_CAP, 5), dtype=torch.int32, device=dev)
        ncomp = torch.empty((1,), dtype=torch.int32, device=dev)
        scratch = torch.empty((2 * h * w,), dtype=torch.int32, device=dev)
        self.eng._call(self.lib.mitb_op_cc_label, _ptr(mask), h, w, _ptr(labels), _ptr(stats), _ptr(ncomp), CC_CAP, _ptr(scratch), self.eng._stream())
        n = int(self.eng.d2h(ncomp)[0])
        if n > CC_CAP:
            raise MitbError(f"mask refinement: {n} connected components exceed the capacity {CC_CAP}")
        return labels, (self.eng.d2h(stats[:n]) if n else np.zeros((0, 5), np.int32))

    def _ellipse(self, k: int) -> np.ndarray:
        if k not in self._se_cache:
            self._se_cache[k] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)).astype(np.uint8)
        return self._se_cache[k]

    def refine(self, img: np.ndarray, raw_mask: np.ndarray, lines_pts: List[np.ndarray], dilation_offset: int = 0, kernel_size: int = 3,
               keep_threshold: float = 1e-2) -> np.ndarray:
        """`dispatch` from the resize on (mask_refinement/__init__.py:13-31): returns the final uint8 mask at the page size."""
        eng, lib = self.eng, self.lib
        H, W = raw_image_hw = img.shape[:2]
        scale = max(min((raw_mask.shape[0] - H / 3) / raw_mask.shape[0], 1), 0.5)
        w, h = int(W * scale), int(H * scale)
        img_d = self.resize(self._dev(img), w, h)
        mask_d = self.resize(self._dev(raw_mask), w, h, binarize=True)
        lines = [Quadrilateral(np.asarray(l) * scale, "", 0) for l in lines_pts]
        if not lines:
            return np.zeros(raw_image_hw, np.uint8)
        polys = [np.asarray(q.pts, dtype=np.float64) for q in lines]
        fonts = [float(q.font_size) for q in lines]
        boxes = np.array([[q.aabb.x, q.aabb.y, q.aabb.w, q.aabb.h] for q in lines], dtype=np.float64).astype(np.int32)   # BBox.xywh: int32 truncation
        boxes_d = self._dev(boxes)                      # (device temporaries stay referenced until their launch is enqueued)
        eng._call(lib.mitb_op_cut_rects, _ptr(mask_d), h, w, _ptr(boxes_d), len(boxes), eng._stream())
        labels, stats = self.components(mask_d)
        owner = assign_components(stats, polys, fonts, keep_threshold)
        if not (owner >= 0).any():
            return np.zeros(raw_image_hw, np.uint8)
        n = h * w
        omap = torch.empty((n,), dtype=torch.int32, device=img_d.device)
        owner_d = self._dev(owner)
        eng._call(lib.mitb_op_owner_map, _ptr(labels), _ptr(owner_d), n, _ptr(omap), eng._stream())
        # per line: union rectangle of its components -> CRF region (rect1) and dilation region (rect2)
        crf2, crf5, dil, ses, se_off = [], [], [], [], {}
        pix0 = slot2 = slot5 = 0
        for i in range(len(lines)):
            own = stats[owner == i]
            if len(own) == 0:
                continue
            x1, y1 = int(own[:, 0].min()), int(own[:, 1].min())
            w1, h1 = int(own[:, 2].max()) + 1 - x1, int(own[:, 3].max()) + 1 - y1
            text_size = min(w1, h1, fonts[i])
            x1, y1, w1, h1 = _extend_rect(x1, y1, w1, h1, w, h, int(text_size * 0.1))
            if w1 <= 0 or h1 <= 0:
                continue
            dilate_size = max((int((text_size + dilation_offset) * 0.3) // 2) * 2 + 1, 3)
            x2, y2, w2, h2 = _extend_rect(x1, y1, w1, h1, w, h, -(-dilate_size // 2))
            if dilate_size not in se_off:
                se_off[dilate_size] = sum(s.size for s in ses)
                ses.append(self._ellipse(dilate_size).reshape(-1))
            npx = w1 * h1
            cap2, cap5 = _pow2_at_least(2 * 3 * npx), _pow2_at_least(2 * 6 * npx)
            # the line index travels as blockIdx.y, so skipped lines keep an (empty) entry
            while len(crf2) < i:
                crf2.append([0] * 8); crf5.append([0] * 8); dil.append([0] * 12)
            crf2.append([x1, y1, w1, h1, pix0, slot2, cap2, 0])
            crf5.append([x1, y1, w1, h1, pix0, slot5, cap5, 0])
            dil.append([x1, y1, w1, h1, x2, y2, max(w2, 0), max(h2, 0), pix0, se_off[dilate_size], dilate_size, 0])
            pix0 += npx; slot2 += cap2; slot5 += cap5
        if pix0 == 0:
            return np.zeros(raw_image_hw, np.uint8)
        nl = len(crf2)
        a2, a5, ad = np.array(crf2, np.int32), np.array(crf5, np.int32), np.array(dil, np.int32)
        filt = eng.bilateral17(img_d)
        nbytes = ctypes.c_ulonglong(0)
        lib.mitb_op_crf_workspace(pix0, slot2, slot5, ctypes.byref(nbytes))
        work = torch.empty((int(nbytes.value),), dtype=torch.uint8, device=img_d.device)
        refined = torch.empty((pix0,), dtype=torch.uint8, device=img_d.device)
        err = torch.zeros((1,), dtype=torch.int32, device=img_d.device)
        omap_hw = omap
        a2_d, a5_d, ad_d, se_d, se3_d = self._dev(a2), self._dev(a5), self._dev(ad), self._dev(np.concatenate(ses)), self._dev(self._ellipse(kernel_size))
        eng._call(lib.mitb_op_dense_crf, _ptr(a2_d), _ptr(a5_d), nl, _ptr(filt), _ptr(omap_hw), w, int((a2[:, 2] * a2[:, 3]).max()),
                  int(a2[:, 6].max()), int(a5[:, 6].max()), pix0, slot2, slot5, CRF_ITERS, SXY_G, W_G, SXY_B, SRGB, W_B, U_ON, _ptr(work), _ptr(refined),
                  _ptr(err), eng._stream())
        final = torch.zeros((h, w), dtype=torch.uint8, device=img_d.device)
        eng._call(lib.mitb_op_dilate_lines, _ptr(ad_d), nl, int((ad[:, 6] * ad[:, 7]).max()), _ptr(omap_hw), _ptr(refined), _ptr(se_d), w, _ptr(final),
                  eng._stream())
        final2 = torch.empty_like(final)
        eng._call(lib.mitb_op_dilate_se, _ptr(final), h, w, _ptr(se3_d), kernel_size, _ptr(final2), eng._stream())
        out = self.resize(final2, W, H, binarize=True)
        code = int(eng.d2h(err)[0])
        if code:
            raise MitbError(f"mask refinement: DenseCRF lattice error {code} (1: key outside the packed range, 2: hash table full)")
        return eng.d2h(out, scratch=True).copy()


_refiners = {}


def get_refiner(device="cuda:0") -> MaskRefiner:
    eng = get_engine(device)
    if id(eng) not in _refiners:
        _refiners[id(eng)] = MaskRefiner(eng)
    return _refiners[id(eng)]


async def dispatch(text_regions, raw_image: np.ndarray, raw_mask: np.ndarray, method: str = "fit_text", dilation_offset: int = 0,
                   ignore_bubble: int = 0, verbose: bool = False, kernel_size: int = 3, device: str = "cuda:0") -> np.ndarray:
    """Signature of manga_translator.mask_refinement.dispatch (+ `device`).  `text_regions`: objects with `.lines` ([n,4,2] arrays)."""
    if method != "fit_text":
        raise NotImplementedError("mask refinement: only method='fit_text' (the reference's other branch reads an undefined variable)")
    if 1 <= ignore_bubble <= 50:
        raise NotImplementedError("mask refinement: the ignore_bubble tail (mask_refinement/__init__.py:33-50) is not covered")
    if raw_image.dtype != np.uint8 or raw_mask.dtype != np.uint8 or raw_image.ndim != 3 or raw_mask.ndim != 2:
        raise MitbError("mask refinement expects a uint8 HxWx3 page and a uint8 HxW raw mask")
    lines = [np.asarray(l) for region in text_regions for l in region.lines]
    return get_refiner(device).refine(raw_image, raw_mask, lines, dilation_offset, kernel_size)

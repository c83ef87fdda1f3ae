"""The one exchange step of the multi-GPU path (north star: "NCCL all-gather of boxes/text/masks back to
rank 0 over NVLink").  Pages are independent, so the only communication is bringing every rank's per-page results home.

Each page becomes ONE fixed-size byte record, so that all ranks hold equal-size buffers and a single
``all_gather_into_tensor`` moves everything:

    header   int32[8]              magic, n_boxes, n_lines, H, W, kmax, lmax, 0
    boxes    int32[kmax,4,2]       detector quads (Quadrilateral.pts), first n_boxes valid
    scores   float64[kmax]         detector Quadrilateral.prob
    l_pts    int32[kmax,4,2]       quads of the OCR lines that survived, first n_lines valid
    l_prob   float64[kmax]         OCR Quadrilateral.prob  (exp(mean log-prob), model_48px_ctc.py:126)
    l_col    int32[kmax,6]         fg_r, fg_g, fg_b, bg_r, bg_g, bg_b
    l_len    int32[kmax]           UTF-8 byte length of the text
    l_text   uint8[kmax,lmax]      UTF-8 text
    mask     uint8[H,W]            detector raw_mask
    page     uint8[H,W,3]          inpainted page

The metadata block (everything before ``mask``) is built on the host (it is host data: contours, CTC strings); the mask and
the inpainted page are copied device-to-device when they already live on the GPU.  Rank 0 de-interleaves the round-robin
order: page i is record [i % world, i // world].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from .compat import Quadrilateral

MAGIC = 0x4D495442          # "MITB"
KMAX, LMAX = 64, 256


@dataclass
class Layout:
    H: int
    W: int
    kmax: int = KMAX
    lmax: int = LMAX

    def __post_init__(self):
        k, l = self.kmax, self.lmax
        off = 0
        self.o = {}
        for name, nbytes in (("header", 32), ("boxes", k * 32), ("scores", k * 8), ("l_pts", k * 32), ("l_prob", k * 8),
                             ("l_col", k * 24), ("l_len", k * 4), ("l_text", k * l)):
            self.o[name] = (off, nbytes)
            off += nbytes
        self.meta_bytes = off
        off = (off + 255) & ~255
        self.o["mask"] = (off, self.H * self.W)
        off += self.H * self.W
        off = (off + 255) & ~255
        self.o["page"] = (off, self.H * self.W * 3)
        off += self.H * self.W * 3
        self.record_bytes = (off + 255) & ~255


def _quad_pts(q) -> np.ndarray:
    return np.asarray(q.pts, dtype=np.int64).reshape(4, 2)


def pack_meta(lay: Layout, textlines: Sequence, ocr_lines: Sequence) -> np.ndarray:
    """Host metadata block of one page as uint8[meta_bytes].  Raises if a page exceeds the fixed capacities."""
    k, l = lay.kmax, lay.lmax
    if len(textlines) > k or len(ocr_lines) > k:
        raise ValueError(f"page has {len(textlines)} boxes / {len(ocr_lines)} lines, record capacity is {k}")
    buf = np.zeros(lay.meta_bytes, np.uint8)

    def view(name, dtype, shape):
        o, n = lay.o[name]
        return buf[o:o + n].view(dtype).reshape(shape)

    view("header", np.int32, (8,))[:] = (MAGIC, len(textlines), len(ocr_lines), lay.H, lay.W, k, l, 0)
    b, s = view("boxes", np.int32, (k, 4, 2)), view("scores", np.float64, (k,))
    for i, q in enumerate(textlines):
        b[i] = _quad_pts(q)
        s[i] = float(q.prob)
    lp, pr, col, ln, tx = (view("l_pts", np.int32, (k, 4, 2)), view("l_prob", np.float64, (k,)), view("l_col", np.int32, (k, 6)),
                           view("l_len", np.int32, (k,)), view("l_text", np.uint8, (k, l)))
    for i, q in enumerate(ocr_lines):
        raw = q.text.encode("utf-8")
        if len(raw) > l:
            raise ValueError(f"OCR line of {len(raw)} UTF-8 bytes exceeds the record capacity {l}")
        lp[i] = _quad_pts(q)
        pr[i] = float(q.prob)
        col[i] = (q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b)
        ln[i] = len(raw)
        tx[i, :len(raw)] = np.frombuffer(raw, np.uint8)
    return buf


def pack_page(lay: Layout, record: torch.Tensor, textlines, ocr_lines, raw_mask, inpainted) -> None:
    """Fill one record (uint8[record_bytes] view of the rank's result buffer, host or device).  raw_mask / inpainted may be numpy
    arrays or torch tensors already on the record's device (then the copy never touches the host)."""
    meta = torch.from_numpy(pack_meta(lay, textlines, ocr_lines))
    record[:lay.meta_bytes].copy_(meta, non_blocking=True)
    for name, src in (("mask", raw_mask), ("page", inpainted)):
        o, n = lay.o[name]
        t = src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src))
        if t.numel() != n or t.dtype != torch.uint8:
            raise ValueError(f"{name}: expected {n} uint8 elements, got {tuple(t.shape)} {t.dtype}")
        record[o:o + n].copy_(t.reshape(-1), non_blocking=True)


@dataclass
class GatheredPage:
    textlines: List[Quadrilateral]
    ocr_lines: List[Quadrilateral]
    raw_mask: np.ndarray
    inpainted: np.ndarray


def unpack_page(lay: Layout, record: np.ndarray, copy: bool = True) -> GatheredPage:
    """Inverse of pack_page on a host uint8[record_bytes] array.  copy=False returns raw_mask / inpainted as views of `record`
    (valid until the buffer is overwritten by the next exchange) instead of 12.6 MB copies per page."""
    def view(name, dtype, shape):
        o, n = lay.o[name]
        return record[o:o + n].view(dtype).reshape(shape)

    hdr = view("header", np.int32, (8,))
    if int(hdr[0]) != MAGIC or (int(hdr[3]), int(hdr[4]), int(hdr[5]), int(hdr[6])) != (lay.H, lay.W, lay.kmax, lay.lmax):
        raise ValueError("result record does not match the layout (corrupt exchange buffer?)")
    nb, nl = int(hdr[1]), int(hdr[2])
    k, l = lay.kmax, lay.lmax
    b, s = view("boxes", np.int32, (k, 4, 2)), view("scores", np.float64, (k,))
    textlines = [Quadrilateral(b[i].astype(np.int64), "", float(s[i])) for i in range(nb)]
    lp, pr, col, ln, tx = (view("l_pts", np.int32, (k, 4, 2)), view("l_prob", np.float64, (k,)), view("l_col", np.int32, (k, 6)),
                           view("l_len", np.int32, (k,)), view("l_text", np.uint8, (k, l)))
    lines = []
    for i in range(nl):
        q = Quadrilateral(lp[i].astype(np.int64), bytes(tx[i, :int(ln[i])]).decode("utf-8"), float(pr[i]))
        q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b = (int(v) for v in col[i])
        lines.append(q)
    mask, page = view("mask", np.uint8, (lay.H, lay.W)), view("page", np.uint8, (lay.H, lay.W, 3))
    return GatheredPage(textlines, lines, mask.copy() if copy else mask, page.copy() if copy else page)


def gather_records(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather the per-rank record buffers uint8[pages_per_rank, record_bytes] -> [world, pages_per_rank, record_bytes] on
    every rank: one NCCL collective over NVLink/NVSwitch (gloo in the CPU tests)."""
    import torch.distributed as dist
    if world == 1:
        return local[None]
    local = local.contiguous()
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "gloo":
        dist.all_gather(list(out.unbind(0)), local)
    else:
        dist.all_gather_into_tensor(out, local)
    return out


def unpack_gathered(lay: Layout, gathered: torch.Tensor, n_pages_total: int, pinned: Optional[torch.Tensor] = None,
                    copy: bool = True) -> List[GatheredPage]:
    """Rank 0: device [world, pages_per_rank, record_bytes] -> pages in their original order (page i was processed by rank
    i % world as its (i // world)-th page).  One D2H copy (into `pinned` when given); copy=False leaves the page-sized arrays as
    views of that host buffer."""
    world = gathered.shape[0]
    if gathered.is_cuda:
        host = pinned if pinned is not None else torch.empty(gathered.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream(gathered.device).synchronize()
    else:
        host = gathered
    arr = host.numpy()
    return [unpack_page(lay, arr[i % world, i // world], copy) for i in range(n_pages_total)]

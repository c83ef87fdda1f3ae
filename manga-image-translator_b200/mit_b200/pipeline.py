"""Page-level driver of the hot path: detect -> OCR -> inpaint for a list of pages on ONE GPU, plus the multi-GPU
sharding used by bench.py (round-robin page scatter, results all-gathered to every rank over NCCL; the reference itself
runs pages strictly one at a time, manga_translator/manga_translator.py:1491-1519).

Two entry points per page:
  * ``process_page``      -- the user-facing path: host numpy in, host results out, through the three plugin ``infer`` calls
                             (what `dispatch()` does in the reference, detection/__init__.py:35-40 etc.);
  * ``run_resident``      -- the same device work with inputs already staged in HBM (kernel-throughput measurement).
"""
from __future__ import annotations

import asyncio
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import exchange, plugins, synth
from .compat import InpainterConfig, OcrConfig, chunks
from .engine import get_engine
from .host import mpe
from .host.geometry import warp_record


@dataclass
class PageResult:
    textlines: list
    raw_mask: np.ndarray
    ocr_lines: list
    inpainted: np.ndarray


@dataclass
class StagedPage:
    page_u8: torch.Tensor                   # [H,W,3] uint8 on device
    ocr_chunks: list                        # per chunk of <= 16 lines: (float64 [n,16] warp records on the device, canvas width)
    mask_u8: torch.Tensor                   # [H,W] uint8 on device
    rel_pos: Optional[torch.Tensor]
    direct: Optional[torch.Tensor]
    bytes: int = 0


class HotPath:
    """The three plugins bound to one CUDA device, constructed with injected (seeded) weights or checkpoint files."""

    def __init__(self, device: str, dbnet_sd=None, ocr_sd=None, dictionary=None, lama_sd=None, mpe_sd=None, large=False,
                 detect_size=2048, inpainting_size=2048):
        self.device = device
        self.detect_size, self.inpainting_size = detect_size, inpainting_size
        if dbnet_sd is not None:
            plugins.DBConvNextDetector.set_state_dict(dbnet_sd)
        if ocr_sd is not None:
            plugins.Model48pxCTCOCR.set_state_dict(ocr_sd)
            plugins.Model48pxCTCOCR.set_dictionary(dictionary)
        inp_cls = plugins.LamaLargeInpainter if large else plugins.LamaMPEInpainter
        if lama_sd is not None:
            inp_cls.set_state_dict({"gen_state_dict": lama_sd, "str_state_dict": mpe_sd})
        self.det, self.ocr, self.inp = plugins.DBConvNextDetector(), plugins.Model48pxCTCOCR(), inp_cls()
        self.use_mpe = not large
        for p in (self.det, self.ocr, self.inp):
            asyncio.run(p.load(device))
        self.engine = get_engine(device)

    def close(self):
        for p in (self.det, self.ocr, self.inp):
            asyncio.run(p.unload())

    # ------------------------------------------------------------------ user-facing path (host buffers)
    def process_page(self, page: np.ndarray, quads, mask: np.ndarray, keep_on_device: bool = False) -> PageResult:
        """`keep_on_device`: leave the inpainted page in HBM (torch uint8 CUDA tensor) for the NCCL exchange instead of copying it
        to the host - the multi-GPU driver's ranks hand their results to rank 0 over NVLink, not through their own host."""
        textlines, raw_mask, _ = asyncio.run(self.det.infer(page, self.detect_size, 0.5, 0.7, 2.3))
        lines = asyncio.run(self.ocr.infer(page, quads, OcrConfig(prob=0.0)))
        if keep_on_device:
            out = asyncio.run(self.inp.infer(page, mask, InpainterConfig(), self.inpainting_size, _device_out=True))
        else:
            out = asyncio.run(self.inp.infer(page, mask, InpainterConfig(), self.inpainting_size))
        return PageResult(textlines, raw_mask, lines, out)

    def process_page_chain(self, page: np.ndarray, ocr_prob: float = 0.0, dilation_offset: int = 20, kernel_size: int = 3):
        """The reference's own stage order for one page, every stage on the GPU path (manga_translator.py:431-600): detect ->
        OCR of the DETECTED lines -> text-line merge (host) -> mask refinement (8f N1) -> inpaint with the REFINED
        mask.  Returns (text regions, refined mask uint8 [H,W], inpainted page uint8 [H,W,3]).  `process_page` above is the
        benchmark workload instead (fixed synthetic quads and mask, detector output not fed forward)."""
        from . import mask_refinement
        from .host import textline_merge
        textlines, raw_mask, _ = asyncio.run(self.det.infer(page, self.detect_size, 0.5, 0.7, 2.3))
        lines = asyncio.run(self.ocr.infer(page, textlines, OcrConfig(prob=ocr_prob))) if textlines else []
        regions = textline_merge.dispatch(lines, page.shape[1], page.shape[0], engine=self.engine) if lines else []
        if not regions:
            return [], np.zeros(page.shape[:2], np.uint8), page.copy()
        mask = asyncio.run(mask_refinement.dispatch(regions, page, raw_mask, "fit_text", dilation_offset, 0, False, kernel_size,
                                                    device=str(self.engine.device)))
        out = asyncio.run(self.inp.infer(page, mask, InpainterConfig(), self.inpainting_size))
        return regions, mask, out

    def process_pages(self, items, workers: int = 4, keep_on_device: bool = False) -> List[PageResult]:
        """A batch of (page, quads, mask) through the same three ``infer`` calls per page, with `workers` host threads so one
        page's host work (H2D/D2H, contours, crops, CTC collapse) overlaps other pages' kernels.  GPU submissions are
        serialised by the engine lock and execute in stream order; results come back in input order."""
        if workers <= 1 or len(items) <= 1:
            return [self.process_page(*it, keep_on_device=keep_on_device) for it in items]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(lambda it: self.process_page(*it, keep_on_device=keep_on_device), items))

    # ------------------------------------------------------------------ device-resident path
    def stage(self, page: np.ndarray, quads, mask: np.ndarray) -> StagedPage:
        eng = self.engine
        dev = eng.device
        # OCR lines: only the 4-point homographies are host work; the crops themselves are cut out of the resident page by
        # mitb_op_warp_lines_u8 inside run_resident (one launch per chunk of 16 lines, sorted by width like the reference)
        recs = [warp_record(q, page.shape[0], page.shape[1], q.direction, 48) for q in quads]
        perm = sorted(range(len(recs)), key=lambda i: recs[i][1])
        chunks_dev = []
        for indices in chunks(perm, 16):
            widths = [recs[i][1] for i in indices]
            rec = np.stack([recs[i][0] for i in indices])
            chunks_dev.append((torch.from_numpy(rec).to(dev), max(widths) + 7 + 128))
        rel = direct = None
        if self.use_mpe:
            r, d = mpe.mpe_tables_256(((mask.astype(np.float32) / 255.0) >= 0.5).astype(np.float32))
            rel, direct = torch.from_numpy(r[None]).to(dev), torch.from_numpy(d[None]).to(dev)
        sp = StagedPage(torch.from_numpy(page).to(dev), chunks_dev, torch.from_numpy(mask).to(dev), rel, direct)
        sp.bytes = sum(t.numel() * t.element_size() for t in [sp.page_u8, sp.mask_u8] + [c[0] for c in chunks_dev] +
                       ([rel, direct] if rel is not None else []))
        return sp

    def run_resident(self, sp: StagedPage):
        eng = self.engine
        filt = eng.bilateral17(sp.page_u8)
        db, dmask = eng.dbnet_forward(filt[None])
        ocr = []
        for rec, wp in sp.ocr_chunks:
            pred, logprob, colors = eng.ocr_forward(eng.warp_lines(sp.page_u8, rec, wp))
            ocr.append(eng.ctc_collapse(pred, logprob, colors))
        out = eng.lama_infer_u8(sp.page_u8, sp.mask_u8, sp.rel_pos, sp.direct, composite=True)
        return db, dmask, ocr, out


def shard_indices(n_pages_total: int, rank: int, world: int) -> List[int]:
    """Round-robin page scatter: page i is processed by rank i mod world."""
    return list(range(rank, n_pages_total, world))


class ResultExchange:
    """Per-rank fixed-size result buffer + the single all-gather that brings every page's boxes / text / mask / inpainted page to
    rank 0 (exchange.py has the record layout).  One instance per rank, reused every step."""

    def __init__(self, device, pages_per_rank: int, H: int, W: int, kmax: int = exchange.KMAX, lmax: int = exchange.LMAX):
        self.lay = exchange.Layout(H, W, kmax, lmax)
        self.device = torch.device(device)
        self.buf = torch.zeros((pages_per_rank, self.lay.record_bytes), dtype=torch.uint8, device=self.device)
        self._pinned = None

    def pack(self, results: List[PageResult]):
        assert len(results) == self.buf.shape[0]
        for i, r in enumerate(results):
            exchange.pack_page(self.lay, self.buf[i], r.textlines, r.ocr_lines, r.raw_mask, r.inpainted)

    def exchange(self, world: int, rank: int, n_pages_total: int, copy: bool = False):
        """Collective.  Returns the pages in original order on rank 0, None elsewhere.  With copy=False (default) the masks and pages
        are views of this object's pinned host buffer, valid until the next exchange()."""
        g = exchange.gather_records(self.buf, world)
        if rank != 0:
            return None
        if g.is_cuda and (self._pinned is None or self._pinned.shape != g.shape):
            self._pinned = torch.empty(g.shape, dtype=torch.uint8, pin_memory=True)
        return exchange.unpack_gathered(self.lay, g, n_pages_total, self._pinned, copy)

    @property
    def gathered_bytes(self) -> int:
        return self.buf.numel()


def gather_results(local_u8: torch.Tensor, world: int) -> Optional[torch.Tensor]:
    """All-gather equal-size per-rank result buffers over NCCL (NVLink); returns [world, ...] on every rank."""
    return exchange.gather_records(local_u8, world)

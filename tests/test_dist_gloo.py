"""N>1 host logic on CPU (gloo, world size 2, spawned processes): round-robin page sharding and THE exchange step of the
product code -- pipeline.ResultExchange (fixed-size per-page records: boxes / scores / OCR text / colours / raw mask /
inpainted page, one all-gather, rank 0 de-interleaves) -- compared with the single-process result.  The NCCL/NVLink variant
of exactly this code runs in bench.py under torchrun."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mit_b200 import exchange
from mit_b200.compat import Quadrilateral
from mit_b200.pipeline import PageResult, ResultExchange, gather_results, shard_indices

H, W = 64, 48


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_result(i: int) -> PageResult:
    """Deterministic stand-in for page i's results (what the three plugins return)."""
    rng = np.random.default_rng(1000 + i)
    nb, nl = int(rng.integers(0, 6)), int(rng.integers(0, 5))

    def quad():
        x0, y0 = int(rng.integers(0, W - 20)), int(rng.integers(0, H - 20))
        return np.array([[x0, y0], [x0 + 15, y0], [x0 + 15, y0 + 9], [x0, y0 + 9]], np.int64)

    boxes = [Quadrilateral(quad(), "", float(rng.random())) for _ in range(nb)]
    lines = []
    for k in range(nl):
        q = Quadrilateral(quad(), "ページ%d 行%d ✓" % (i, k), float(rng.random()))
        q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b = (int(v) for v in rng.integers(0, 256, 6))
        lines.append(q)
    return PageResult(boxes, rng.integers(0, 256, (H, W), dtype=np.uint8), lines, rng.integers(0, 256, (H, W, 3), dtype=np.uint8))


def _same(a, b):
    assert len(a.textlines) == len(b.textlines) and len(a.ocr_lines) == len(b.ocr_lines)
    for p, q in zip(a.textlines, b.textlines):
        assert np.array_equal(p.pts, q.pts) and p.prob == q.prob
    for p, q in zip(a.ocr_lines, b.ocr_lines):
        assert np.array_equal(p.pts, q.pts) and p.text == q.text and p.prob == q.prob
        assert (p.fg_r, p.fg_g, p.fg_b, p.bg_r, p.bg_g, p.bg_b) == (q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b)
    assert np.array_equal(a.raw_mask, b.raw_mask) and np.array_equal(a.inpainted, b.inpainted)


def _worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_pages, rank, world)
    ex = ResultExchange("cpu", len(mine), H, W)
    ex.pack([_fake_result(i) for i in mine])
    pages = ex.exchange(world, rank, n_pages)
    legacy = gather_results(torch.full((2, 3), float(rank)), world)          # the plain equal-size gather still works
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)         # the max-over-ranks timing reduction of bench.py
    ok = None
    if rank == 0:
        try:
            assert len(pages) == n_pages
            for i, p in enumerate(pages):
                _same(p, _fake_result(i))             # identical to what a single process computes for page i
            assert legacy[:, 0, 0].tolist() == [float(r) for r in range(world)]
            ok = "ok"
        except Exception as e:  # noqa: BLE001
            ok = repr(e)
        q.put((ok, t.item(), mine))
    else:
        assert pages is None
    dist.destroy_process_group()


def test_round_robin_shard_and_result_exchange_world2():
    assert shard_indices(8, 0, 2) == [0, 2, 4, 6] and shard_indices(8, 1, 2) == [1, 3, 5, 7]
    assert sorted(sum((shard_indices(256, r, 8) for r in range(8)), [])) == list(range(256))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax, mine = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == "ok" and tmax == 2.0 and mine == [0, 2, 4, 6]


def test_record_roundtrip_and_capacity_errors():
    lay = exchange.Layout(H, W, kmax=4, lmax=16)
    r = _fake_result(3)
    rec = torch.zeros(lay.record_bytes, dtype=torch.uint8)
    exchange.pack_page(lay, rec, r.textlines[:4], [], r.raw_mask, torch.from_numpy(r.inpainted))
    back = exchange.unpack_page(lay, rec.numpy())
    assert len(back.textlines) == min(4, len(r.textlines)) and np.array_equal(back.inpainted, r.inpainted)
    import pytest
    too_long = Quadrilateral(np.array([[0, 0], [9, 0], [9, 9], [0, 9]]), "x" * 40, 0.5)
    with pytest.raises(ValueError):
        exchange.pack_meta(lay, [], [too_long])
    with pytest.raises(ValueError):
        exchange.pack_meta(lay, [too_long] * 5, [])
    bad = rec.numpy().copy()
    bad[0] ^= 0xFF
    with pytest.raises(ValueError):
        exchange.unpack_page(lay, bad)

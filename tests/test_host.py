"""Host-side logic (no GPU): geometry / detector post-processing / MPE tables / CTC collapse / rearrangement / the C ABI
surface.  Where the reference helper is importable here (build container) it is the checker; otherwise the oracle is."""
import asyncio
import ctypes
import os
import re
import warnings

import cv2
import numpy as np
import pytest

from mit_b200 import synth
from mit_b200.host import det_post, geometry, mpe, rearrange
from oracle import nets, refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def test_abi_header_and_library_agree():
    """Every function declared in include/mitb.h is exported by libmitb.so and bound in mit_b200._lib (no compute calls)."""
    from mit_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mitb.h")).read()
    declared = set(re.findall(r"\b(mitb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mitb_ctx", "mitb_tensor"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mitb_ocr_timesteps(647) == 160 and lib.mitb_ocr_timesteps(512) == 127
    assert b"sm_100a" in lib.mitb_version()


def test_no_cpu_fallback():
    """Without a CUDA device context creation fails loudly (and the plugins refuse non-CUDA devices)."""
    import torch
    from mit_b200 import MitbError, plugins
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mit_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.mitb_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.mitb_last_error(None)
    det = plugins.DBConvNextDetector()
    with pytest.raises(MitbError):
        asyncio.run(det.load("cpu"))
    with pytest.raises(Exception):
        asyncio.run(det.infer(np.zeros((8, 8, 3), np.uint8), 2048, 0.5, 0.7, 2.3))   # infer before load


def test_mpe_tables_match_oracle():
    rng = np.random.default_rng(1)
    for (h, w) in ((256, 256), (120, 312), (64, 48)):
        m = np.zeros((h, w), np.float32)
        for _ in range(4):
            y, x = rng.integers(0, h - 8), rng.integers(0, w - 8)
            m[y:y + rng.integers(4, h // 2), x:x + rng.integers(4, w // 2)] = 1
        a, b = mpe.mpe_tables(m), nets.mpe_tables(m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for m in (np.zeros((40, 40), np.float32), np.ones((40, 40), np.float32)):
        a, b = mpe.mpe_tables(m), nets.mpe_tables(m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_ctc_collapse_matches_oracle():
    from mit_b200.plugins import ctc_collapse
    rng = np.random.default_rng(2)
    idx = rng.integers(0, 4, (5, 60))
    steps = ctc_collapse(idx)
    ref = nets.ctc_greedy(idx, np.zeros(idx.shape, np.float32), np.zeros(idx.shape + (6,), np.float32))
    assert [[int(idx[b, t]) for t in s] for b, s in enumerate(steps)] == [[c[0] for c in l] for l in ref]
    assert [len(s) for s in ctc_collapse(np.zeros((2, 7), np.int64))] == [0, 0]


def test_polygon_helpers():
    sq = np.array([[0, 0], [4, 0], [4, 3], [0, 3]])
    assert geometry.polygon_area(sq) == 12 and geometry.polygon_perimeter(sq) == 14
    assert geometry.hull_area(np.array([[0, 0], [4, 0], [2, 1], [4, 3], [0, 3]])) == 12
    far = sq + np.array([10, 0])
    assert abs(geometry.polygon_distance(sq, far) - 6) < 1e-9
    assert geometry.polygon_distance(sq, sq + 1) == 0 and geometry.polygon_distance(sq, np.array([[1, 1], [2, 1], [2, 2], [1, 2]])) == 0
    diag = np.array([[7, 7], [9, 7], [9, 9], [7, 9]])
    assert abs(geometry.polygon_distance(sq, diag) - 5) < 1e-9        # corner to corner (4,3)->(7,7)


def test_unclip_and_boxes_on_synthetic_prob_map():
    prob = np.zeros((200, 300), np.float32)
    prob[50:80, 40:200] = 0.9
    prob[120:124, 10:14] = 0.9     # too small after unclip filter? (short side 4 -> kept only if >= 3)
    boxes, scores = det_post.boxes_from_prob(prob, 0.5, 0.7, 2.3, 300, 200)
    polys = det_post.polys_from_boxes(boxes, scores, 1.0, 1.0)
    big = [p for p in polys if (p[:, 0].max() - p[:, 0].min()) > 100]
    assert len(big) == 1
    p = big[0]
    # rectangle 160x30 (contour spans 159x29): distance = A*r/L
    d = (159 * 29) * 2.3 / (2 * (159 + 29))
    assert abs((p[:, 0].max() - p[:, 0].min()) - (159 + 2 * d)) <= 2 and abs((p[:, 1].max() - p[:, 1].min()) - (29 + 2 * d)) <= 2
    assert p.sum(axis=1).argmin() == 0                              # starts at the top-left corner


def test_synthetic_page_is_deterministic():
    p1, b1, m1 = synth.make_page(3, 512, 384, 6)
    p2, b2, m2 = synth.make_page(3, 512, 384, 6)
    assert np.array_equal(p1, p2) and np.array_equal(m1, m2) and len(b1) == 6
    assert all(np.array_equal(a, b) for a, b in zip(b1, b2))
    q = synth.make_quads(b1)
    assert [x.direction for x in q] == ["h"] * 3 + ["v"] * 3


@needs_ref
def test_quadrilateral_matches_reference():
    warnings.filterwarnings("ignore")
    U = refload.load()["utils"]
    rng = np.random.default_rng(4)
    page, boxes, _ = synth.make_page(1, 1024, 768, 10)
    for b in boxes + [np.array([[100, 100], [400, 130], [390, 190], [95, 160]]), np.array([[50, 50], [90, 60], [70, 400], [30, 390]])]:
        b = b[rng.permutation(4)]
        mine, ref = geometry.Quadrilateral(b, "", 1.0), U.Quadrilateral(b, "", 1.0)
        assert np.array_equal(mine.pts, ref.pts) and mine.direction == ref.direction
        assert abs(mine.aspect_ratio - ref.aspect_ratio) < 1e-6 and abs(mine.font_size - ref.font_size) < 1e-6
        assert tuple(mine.aabb) == (ref.aabb.x, ref.aabb.y, ref.aabb.w, ref.aabb.h)
        assert mine.is_approximate_axis_aligned == ref.is_approximate_axis_aligned and abs(mine.angle - ref.angle) < 1e-6
        for d in ("h", "v"):
            assert np.array_equal(mine.get_transformed_region(page, d, 48), ref.get_transformed_region(page, d, 48))


@needs_ref
def test_rearrange_matches_reference():
    warnings.filterwarnings("ignore")
    U = refload.load()["utils"]

    def fwd(batch, device=None):
        batch = np.asarray(batch).astype(np.float32)
        s = batch.shape[1]
        db = np.stack([batch[..., 0] / 255.0, batch[..., 1] / 255.0], 1).astype(np.float32)
        mask = np.stack([cv2.resize(b[..., 2], (s // 2, s // 2)) / 255.0 for b in batch])[:, None].astype(np.float32)
        return db, mask
    rng = np.random.default_rng(0)
    for shape in ((3000, 500, 3), (500, 3300, 3), (1024, 768, 3)):
        img = cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (0, 0), 5)
        r = U.det_rearrange_forward(img, fwd, 1024, 4)
        o = rearrange.rearrange_forward(img, fwd, 1024, 4)
        if r[0] is None:
            assert o[0] is None
        else:
            assert np.array_equal(r[0], o[0]) and np.array_equal(r[1], o[1])


@needs_ref
def test_detector_helpers_match_reference():
    warnings.filterwarnings("ignore")
    refload.load()
    import importlib
    du = importlib.import_module("manga_translator.detection.default_utils.dbnet_utils")
    ip = importlib.import_module("manga_translator.detection.default_utils.imgproc")
    rep = du.SegDetectorRepresenter(0.5, 0.7, unclip_ratio=2.3)
    rng = np.random.default_rng(5)
    prob = cv2.GaussianBlur(rng.random((120, 160)).astype(np.float32), (0, 0), 4)
    cnts, _ = cv2.findContours(((prob > prob.mean()) * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    for c in cnts[:10]:
        c = c.squeeze(1)
        if len(c) < 3:
            continue
        a, b = det_post.mini_box(c), rep.get_mini_boxes(c)
        assert np.allclose(np.array(a[0]), np.array(b[0])) and a[1] == b[1]
        assert abs(det_post.box_score(prob, c) - rep.box_score_fast(prob, c)) < 1e-12
    img = rng.integers(0, 256, (300, 200, 3), dtype=np.uint8)
    for size in (512, 256, 300):
        a, b = det_post.resize_aspect_ratio(img, size, cv2.INTER_LINEAR), ip.resize_aspect_ratio(img, size, cv2.INTER_LINEAR, mag_ratio=1)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]


def test_bench_roofline_object_from_profile():
    """bench.py's roofline block is pure host code: feed it the per-class profile of a recorded run."""
    import json as _json
    import os as _os
    import bench
    prof = _json.loads(open(_os.path.join(bench.ROOT, "profiles", "r01_layers_tma_v13.txt")).read().strip().splitlines()[-1])
    peaks = bench.load_peaks()
    roof = bench.roofline_from_profile(prof, peaks, 1)
    assert roof["kernel"] == "conv_tc" and roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s"
    assert 0.0 < roof["frac"] < 1.0 / 3.0 + 1e-6                       # bf16x3 cannot exceed a third of the bf16 peak
    assert abs(roof["achieved"] * 1e12 * prof["conv_tc"]["ms"] / 1e3 - prof["conv_tc"]["flops"]) < 1e-3 * prof["conv_tc"]["flops"]
    assert roof["traffic"] is None or roof["traffic"] > roof["algorithmic_bytes_per_launch"] * 0.5
    assert set(roof["classes"]) == set(prof)
    assert bench.roofline_from_profile({}, peaks, 1) is None
    _json.dumps(roof)


def test_host_logic_on_empty_and_degenerate_inputs():
    """Edge cases of the host glue: empty detector map, no text lines, all-blank and all-repeat CTC rows, empty shards."""
    import numpy as np
    from mit_b200 import plugins, synth
    from mit_b200.host import det_post, geometry
    from mit_b200.pipeline import shard_indices
    # nothing above threshold -> no boxes, no polygons
    boxes, scores = det_post.boxes_from_prob(np.zeros((64, 48), np.float32), 0.5, 0.7, 2.3, 48, 64)
    assert len(boxes) == 0 and len(scores) == 0
    assert len(det_post.polys_from_boxes(boxes, scores, 1.0, 1.0)) == 0
    # a single saturated blob still yields exactly one box
    prob = np.zeros((64, 96), np.float32); prob[20:40, 10:80] = 0.99
    boxes, scores = det_post.boxes_from_prob(prob, 0.5, 0.7, 2.3, 96, 64)
    assert len(boxes) == 1 and scores[0] > 0.9
    # direction graph / quads of nothing
    assert geometry.generate_text_direction([]) == [] or list(geometry.generate_text_direction([])) == []
    assert len(synth.make_quads([])) == 0
    # CTC collapse: all blank, all the same symbol, alternating with blanks
    idx = np.array([[0, 0, 0, 0], [5, 5, 5, 5], [5, 0, 5, 0], [1, 2, 2, 3]], np.int64)
    kept = plugins.ctc_collapse(idx)
    assert [k.tolist() for k in kept] == [[], [0], [0, 2], [0, 1, 3]]
    # sharding more ranks than pages leaves some ranks empty, never duplicates or drops a page
    parts = [list(shard_indices(3, r, 8)) for r in range(8)]
    assert sorted(sum(parts, [])) == [0, 1, 2] and sum(1 for p in parts if not p) == 5


def test_bench_lama_ffc_figure_from_recorded_launches():
    """The LaMa FFC block figure of the bench line, computed from a recorded per-layer table (profiles/r01_layers_tma_v13.txt)."""
    import json as _json
    import os as _os
    import bench
    lines = open(_os.path.join(bench.ROOT, "profiles", "r01_layers_tma_v13.txt")).read().strip().splitlines()
    prof = _json.loads(lines[-1])
    launches = []
    for ln in lines[2:-1]:
        f = ln.split()
        if len(f) == 7 and f[0].startswith("conv"):
            kind, m, k, n, cnt, ms = f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), float(f[5])
            launches += [[kind, m, k, n, ms / cnt]] * cnt
    fig = bench.ffc_block_from_launches(launches, prof, 1, bench.load_peaks())
    assert fig is not None and abs(fig["layers_timed"] - 18) < 1e-9          # 9 blocks x 2 FFC layers per LaMa-MPE page
    assert 300 < fig["us_per_layer"] < 3000 and 0 < fig["hbm_frac"] < 1 and 0 < fig["tensor_frac"] < 1.0 / 3.0
    assert fig["binding_term"] == "tensor"                                    # the fused block is tensor bound
    assert bench.ffc_block_from_launches([], prof, 1, bench.load_peaks()) is None
    _json.dumps(fig)


class _FakeEngine:
    """Stands in for mit_b200.engine.Engine so the plugins' HOST logic can run without a GPU (never part of the product path)."""

    def __init__(self, T=20, V=12):
        self.T, self.V = T, V
        self.calls = []

    def h2d(self, t, dtype=None):
        return t

    def d2h(self, t, scratch=False):
        return t

    def ocr_forward(self, region):
        import numpy as np
        n = region.shape[0]
        rng = np.random.default_rng(5)
        pred = rng.integers(0, self.V, (n, self.T)).astype(np.int32)
        pred[:, ::3] = 0
        logprob = (-rng.random((n, self.T)) * 0.2).astype(np.float32)
        colors = rng.random((n, self.T, 6)).astype(np.float32)
        self.calls.append(("ocr", region.shape))
        return pred, logprob, colors

    def lama_infer_u8(self, img, mask, rel, direct, composite=True):
        self.calls.append(("lama", img.shape, mask.shape, None if rel is None else rel.shape, composite))
        return (255 - img).copy()


def test_plugin_host_logic_with_a_fake_engine():
    """OCR post-processing (probability / colour statistics, in-place quad mutation) and the inpainter's resize + composite path."""
    import asyncio
    import numpy as np
    from mit_b200 import plugins, synth
    from mit_b200.compat import InpainterConfig, OcrConfig
    page, boxes, mask = synth.make_page(3, 512, 384, 6)
    quads = synth.make_quads(boxes)
    ocr = plugins.Model48pxCTCOCR.__new__(plugins.Model48pxCTCOCR)
    plugins.Model48pxCTCOCR.__init__(ocr)
    ocr.engine = _FakeEngine()
    ocr.dictionary = ["<S>", "</S>", "<SP>"] + [chr(0x3042 + i) for i in range(9)]
    out = asyncio.run(ocr._infer(page, quads, OcrConfig(), False))
    assert len(out) >= 1 and all(q.text and 0 < q.prob <= 1 for q in out)
    assert all(0 <= c <= 255 for q in out for c in (q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b))
    # reference arithmetic of the statistics, recomputed per element for the first kept line
    eng = _FakeEngine()
    pred, logprob, colors = eng.ocr_forward(np.zeros((6, 48, 8, 3), np.uint8))
    steps = plugins.ctc_collapse(pred)[0]
    want_prob = np.exp(np.mean([float(v) for v in logprob[0, steps]]))
    assert any(abs(q.prob - want_prob) < 1e-12 for q in out)

    inp = plugins.LamaMPEInpainter.__new__(plugins.LamaMPEInpainter)
    plugins.LamaMPEInpainter.__init__(inp)
    inp.engine = _FakeEngine()
    page0, mask0 = page.copy(), mask.copy()
    res = asyncio.run(inp._infer(page, mask, InpainterConfig(), 1024, False))           # no resize: device composite
    assert res.shape == page.shape and inp.engine.calls[-1][-1] is True
    res = asyncio.run(inp._infer(page, mask, InpainterConfig(), 256, False))            # resize: host composite with the {0,1} mask
    m01 = (mask0 >= 127)[:, :, None]
    assert res.shape == page.shape and inp.engine.calls[-1][-1] is False
    assert (res[~np.broadcast_to(m01, res.shape)] == page0[~np.broadcast_to(m01, page0.shape)]).all()   # untouched outside the mask
    assert (page == page0).all() and (mask == mask0).all()                               # borrowed inputs were not written

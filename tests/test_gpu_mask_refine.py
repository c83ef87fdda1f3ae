"""GPU tests of the mask-refinement row: every kernel against cv2 where cv2 defines the result (bit-exact), the
batched DenseCRF against the oracle restatement (oracle/mask_refine_ref.py; parity unpinned against the real pydensecrf), and the
whole `dispatch` against the oracle's statement-order restatement of the reference."""
import asyncio
import types

import cv2
import numpy as np
import pytest
import torch

from mit_b200 import mask_refinement as MR
from mit_b200 import synth
from mit_b200.engine import _ptr, get_engine
from mit_b200.host import geometry
from oracle import mask_refine_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    return MR.get_refiner("cuda:0")


def _page(seed=3, h=768, w=576, n=8):
    page, boxes, _ = synth.make_page(seed, h, w, n)
    raw = cv2.dilate(((page[..., 0] < 100) * 255).astype(np.uint8), np.ones((3, 3), np.uint8))
    return page, boxes, raw


def test_resize_linear_u8_equals_cv2(ref):
    rng = np.random.default_rng(0)
    for (H, W) in ((768, 576), (701, 333), (2048, 1536)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        m = ((rng.random((H, W)) > 0.7) * 255).astype(np.uint8)
        sf = max(min((H - H / 3) / H, 1), 0.5)
        w, h = int(W * sf), int(H * sf)
        for src in (img, m):
            d = torch.from_numpy(src).cuda()
            small = ref.resize(d, w, h)
            want = cv2.resize(src, (w, h), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(small.cpu().numpy(), want)
            back = ref.resize(small, W, H)
            assert np.array_equal(back.cpu().numpy(), cv2.resize(want, (W, H), interpolation=cv2.INTER_LINEAR))
        b = ref.resize(torch.from_numpy(m).cuda(), w, h, binarize=True).cpu().numpy()
        want = cv2.resize(m, (w, h), interpolation=cv2.INTER_LINEAR)
        want[want > 0] = 255
        assert np.array_equal(b, want)


def test_cut_rects_and_components_equal_cv2(ref):
    page, boxes, raw = _page()
    eng = ref.eng
    mask = raw.copy()
    rects = np.array([[b[:, 0].min() - 3, b[:, 1].min() + 2, np.ptp(b[:, 0]) + 5, np.ptp(b[:, 1])] for b in boxes] + [[-5, -5, 40, 30], [560, 750, 100, 100]],
                     dtype=np.int32)
    want = raw.copy()
    for (x, y, w, h) in rects:
        cv2.rectangle(want, (int(x), int(y)), (int(x + w), int(y + h)), (0), 1)
    d = torch.from_numpy(mask).cuda()
    rects_d = torch.from_numpy(rects).cuda()
    eng._call(eng.lib.mitb_op_cut_rects, _ptr(d), mask.shape[0], mask.shape[1], _ptr(rects_d), len(rects), eng._stream())
    assert np.array_equal(d.cpu().numpy(), want)
    # components: same partition, same stats
    rng = np.random.default_rng(1)
    noise = ((rng.random(raw.shape) > 0.55) * 255).astype(np.uint8)            # many small components, diagonal contacts
    spiral = np.zeros((64, 64), np.uint8)
    for k in range(0, 30, 4):
        spiral[k, k:64 - k] = 255; spiral[k:64 - k, 63 - k] = 255; spiral[63 - k, k + 2:64 - k] = 255; spiral[k + 4:64 - k, k + 2] = 255
    for m in (want, noise, spiral, np.zeros((5, 7), np.uint8), np.full((3, 3), 255, np.uint8)):
        labels, stats = ref.components(torch.from_numpy(m).cuda())
        labels = labels.cpu().numpy().reshape(m.shape)
        num, lab_cv, st_cv, _ = cv2.connectedComponentsWithStats(m)
        assert len(stats) == num - 1
        assert ((labels >= 0) == (lab_cv > 0)).all()
        if num > 1:
            pairs = np.unique(np.stack([labels[lab_cv > 0], lab_cv[lab_cv > 0]], 1), axis=0)
            assert len(pairs) == num - 1 and len(np.unique(pairs[:, 0])) == num - 1            # a bijection between the labellings
            for mine, theirs in pairs:
                x0, y0, x1, y1, area = stats[mine]
                assert (x0, y0, x1 - x0 + 1, y1 - y0 + 1, area) == tuple(st_cv[theirs])


def test_dilate_se_equals_cv2(ref):
    rng = np.random.default_rng(2)
    m = ((rng.random((301, 257)) > 0.97) * 255).astype(np.uint8)
    eng = ref.eng
    for k in (3, 5, 17, 37):
        se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
        d, o = torch.from_numpy(m).cuda(), torch.empty((301, 257), dtype=torch.uint8, device="cuda")
        se_d = torch.from_numpy(se.astype(np.uint8)).cuda()
        eng._call(eng.lib.mitb_op_dilate_se, _ptr(d), 301, 257, _ptr(se_d), k, _ptr(o), eng._stream())
        assert np.array_equal(o.cpu().numpy(), cv2.dilate(m, se))


def _crf_device(ref, img, mask_on):
    """One region = the whole (small) image through mitb_op_dense_crf; returns the refined uint8 mask."""
    import ctypes
    eng, lib = ref.eng, ref.lib
    h, w = mask_on.shape
    npx = h * w
    cap2, cap5 = MR._pow2_at_least(6 * npx), MR._pow2_at_least(12 * npx)
    a2 = np.array([[0, 0, w, h, 0, 0, cap2, 0]], np.int32)
    a5 = np.array([[0, 0, w, h, 0, 0, cap5, 0]], np.int32)
    omap = torch.from_numpy(np.where(mask_on, 0, -1).astype(np.int32).reshape(-1)).cuda()
    nbytes = ctypes.c_ulonglong(0)
    lib.mitb_op_crf_workspace(npx, cap2, cap5, ctypes.byref(nbytes))
    work = torch.empty((int(nbytes.value),), dtype=torch.uint8, device="cuda")
    refined = torch.empty((npx,), dtype=torch.uint8, device="cuda")
    err = torch.zeros((1,), dtype=torch.int32, device="cuda")
    a2_d, a5_d, img_d = torch.from_numpy(a2).cuda(), torch.from_numpy(a5).cuda(), torch.from_numpy(img).cuda()      # keep the temporaries alive
    eng._call(lib.mitb_op_dense_crf, _ptr(a2_d), _ptr(a5_d), 1, _ptr(img_d), _ptr(omap), w,
              npx, cap2, cap5, npx, cap2, cap5, MR.CRF_ITERS, MR.SXY_G, MR.W_G, MR.SXY_B, MR.SRGB, MR.W_B, MR.U_ON, _ptr(work), _ptr(refined), _ptr(err),
              eng._stream())
    assert int(err.cpu()[0]) == 0
    return refined.cpu().numpy().reshape(h, w)


def test_dense_crf_matches_oracle_restatement(ref):
    """The batched CUDA DenseCRF against the numpy restatement of densecrf on regions of a synthetic page and on random-colour noise
    (every pixel its own lattice cell: stresses the hash table).  Only the splat's summation order differs (atomics), so labels may
    flip where Q0 and Q1 tie to ~1e-6: allow a handful of pixels."""
    page, boxes, raw = _page()
    rng = np.random.default_rng(4)
    cases = []
    for b in boxes[:3]:
        x0, y0, x1, y1 = b[:, 0].min() - 4, b[:, 1].min() - 4, b[:, 0].max() + 4, b[:, 1].max() + 4
        cases.append((np.ascontiguousarray(page[y0:y1, x0:x1]), raw[y0:y1, x0:x1] > 0))
    cases.append((rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), rng.random((37, 53)) > 0.5))
    for img, on in cases:
        got = _crf_device(ref, img, on)
        want = R.refine_mask(img, (on * 255).astype(np.uint8))
        bad = int((got != want).sum())
        print(f"crf region {img.shape[:2]}: {bad} of {got.size} labels differ from the oracle; text pixels {int((want > 0).sum())}")
        assert bad <= max(2, got.size // 5000)


def test_dispatch_matches_oracle(ref):
    """End to end: mit_b200.mask_refinement.dispatch against the statement-order restatement of the reference's dispatch /
    complete_mask.  cv2 runs with IPP off so that its bilateral filter is OpenCV's own arithmetic (the one the CUDA filter pins)."""
    page, boxes, raw = _page(seed=5, h=1024, w=768, n=12)
    regions = [types.SimpleNamespace(lines=[b.astype(np.float64) for b in boxes[i:i + 3]]) for i in range(0, len(boxes), 3)]
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        want = R.dispatch(regions, page, raw.copy(), geometry.Quadrilateral, dilation_offset=20, kernel_size=3)
    finally:
        cv2.ipp.setUseIPP(ipp)
    got = asyncio.run(MR.dispatch(regions, page, raw.copy(), "fit_text", 20, 0, False, 3))
    assert got.shape == want.shape and got.dtype == np.uint8 and set(np.unique(got)) <= {0, 255}
    inter, union = ((got > 0) & (want > 0)).sum(), ((got > 0) | (want > 0)).sum()
    print(f"dispatch: IoU {inter / union:.6f}, {int((got != want).sum())} of {got.size} pixels differ, mask covers {100.0 * (want > 0).mean():.2f} %")
    assert inter / union >= 0.999
    assert asyncio.run(MR.dispatch([], page, raw.copy())).sum() == 0
    with pytest.raises(NotImplementedError):
        asyncio.run(MR.dispatch(regions, page, raw.copy(), ignore_bubble=5))

"""TEST INFRASTRUCTURE (oracle): numpy restatement of `cv2.warpPerspective(src, M, (w, h))` for uint8 images with the default
INTER_LINEAR / BORDER_CONSTANT(0) - the call `Quadrilateral.get_transformed_region` makes (utils/generic.py:471,478) - and of the
line-record semantics of `mitb_op_warp_lines_u8` (include/mitb.h).

The algorithm lives in a third-party dependency (OpenCV; `opencv-python` unpinned in the reference's requirements.txt, 4.13 installed
here).  It is restated from the published source, modules/imgproc/src/imgwarp.cpp:
  * `WarpPerspectiveInvoker::operator()`: destination walked in blocks (BLOCK_SZ = 32: bh0 = min(16, h), bw0 = min(1024 / bh0, w));
    per pixel, in doubles, X0 = M0*xb + M1*y + M2 (block start xb), W = W0 + M6*x1, W = W ? INTER_TAB_SIZE / W : 0,
    fX = max(INT_MIN, min(INT_MAX, (X0 + M0*x1) * W)), X = cvRound(fX); sx = X >> 5, alpha = (Y & 31) * 32 + (X & 31);
  * `initInterTab2D(INTER_LINEAR, fixpt=true)`: 15-bit fixed-point weights; the entry for alpha = 0 saturates to 32767 and the
    sum-fix loop (which scans k1, k2 in {ksize/2, ksize/2 + 1} = {1, 2} for ksize = 2) adds the missing 1 to weight [1][1];
  * `remapBilinear<FixedPtCast<int, uchar, 15>>`: (sum + (1 << 14)) >> 15, out-of-image samples are the border value.
Pinned: tests/test_host.py::test_warp_oracle_equals_cv2 compares it with the installed cv2 on random quads (bit-exact)."""
import cv2
import numpy as np


def bilinear_itab() -> np.ndarray:
    """BilinearTab_i of initInterTab2D as int32 [1024, 4] (index alpha = ay * 32 + ax; weights for (y,x), (y,x+1), (y+1,x), (y+1,x+1))."""
    tab1 = np.zeros((32, 2), np.float32)
    for i in range(32):
        x = np.float32(i) / np.float32(32)
        tab1[i] = (np.float32(1) - x, x)
    flat = np.zeros(32 * 32 * 4 + 8, np.int32)          # the sum-fix loop of the reference reads past the 2x2 entry (zeros at that time)
    for i in range(32):
        for j in range(32):
            base = (i * 32 + j) * 4
            isum = 0
            for k1 in range(2):
                for k2 in range(2):
                    v = np.float32(tab1[i, k1] * tab1[j, k2])
                    iv = int(np.clip(np.rint(np.float32(v * np.float32(32768))), -32768, 32767))
                    flat[base + k1 * 2 + k2] = iv
                    isum += iv
            if isum != 32768:
                diff = isum - 32768
                Mk = mk = (1, 1)
                for k1 in (1, 2):
                    for k2 in (1, 2):
                        val = flat[base + k1 * 2 + k2]
                        if val < flat[base + mk[0] * 2 + mk[1]]:
                            mk = (k1, k2)
                        elif val > flat[base + Mk[0] * 2 + Mk[1]]:
                            Mk = (k1, k2)
                if diff < 0:
                    flat[base + Mk[0] * 2 + Mk[1]] -= diff
                else:
                    flat[base + mk[0] * 2 + mk[1]] -= diff
    return flat[:32 * 32 * 4].reshape(32 * 32, 4)


_ITAB = None


def warp_inverse(src: np.ndarray, Minv, w: int, h: int) -> np.ndarray:
    """The remap half: `Minv` (9 doubles) maps destination pixels to source coordinates."""
    global _ITAB
    if _ITAB is None:
        _ITAB = bilinear_itab()
    Mi = np.asarray(Minv, dtype=np.float64).reshape(-1)
    rows, cols = src.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    bh0 = min(16, h)
    bw0 = min(1024 // bh0, w)
    xb = (xs // bw0) * bw0
    x1 = (xs - xb).astype(np.float64)
    xb = xb.astype(np.float64)
    yf = ys.astype(np.float64)
    X0 = Mi[0] * xb + Mi[1] * yf + Mi[2]
    Y0 = Mi[3] * xb + Mi[4] * yf + Mi[5]
    W0 = Mi[6] * xb + Mi[7] * yf + Mi[8]
    W = W0 + Mi[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Wv = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.maximum(-2147483648.0, np.minimum(2147483647.0, (X0 + Mi[0] * x1) * Wv))
    fY = np.maximum(-2147483648.0, np.minimum(2147483647.0, (Y0 + Mi[3] * x1) * Wv))
    X, Y = np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    wt = _ITAB[(Y & 31) * 32 + (X & 31)]
    s64 = src.astype(np.int64).reshape(rows, cols, -1)

    def fetch(yy, xx):
        ok = (yy >= 0) & (yy < rows) & (xx >= 0) & (xx < cols)
        v = s64[np.clip(yy, 0, rows - 1), np.clip(xx, 0, cols - 1)]
        return np.where(ok[..., None], v, 0)

    acc = (fetch(sy, sx) * wt[..., 0:1] + fetch(sy, sx + 1) * wt[..., 1:2] + fetch(sy + 1, sx) * wt[..., 2:3] +
           fetch(sy + 1, sx + 1) * wt[..., 3:4])
    out = np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return out.reshape((h, w) + src.shape[2:])


def warp_perspective(src: np.ndarray, M, w: int, h: int) -> np.ndarray:
    """cv2.warpPerspective(src, M, (w, h)): inverts M exactly as OpenCV does (cv::invert, DECOMP_LU) and remaps."""
    return warp_inverse(src, cv2.invert(np.asarray(M, dtype=np.float64))[1], w, h)


def warp_line_record(page: np.ndarray, rec: np.ndarray, canvas_w: int, canvas_h: int = 48) -> np.ndarray:
    """One line of mitb_op_warp_lines_u8: record float64[16] -> uint8 [canvas_h, canvas_w, 3] (zero padded)."""
    x1, y1, cw, ch, w, h, rot = [int(v) for v in rec[9:16]]
    out = np.zeros((canvas_h, canvas_w, 3), np.uint8)
    if cw <= 0 or ch <= 0:
        return out
    region = warp_inverse(page[y1:y1 + ch, x1:x1 + cw], rec[:9], w, h)
    if rot:
        region = np.ascontiguousarray(np.rot90(region, 1))         # == cv2.rotate(ROTATE_90_COUNTERCLOCKWISE)
    hh, ww = min(region.shape[0], canvas_h), min(region.shape[1], canvas_w)
    out[:hh, :ww] = region[:hh, :ww]
    return out

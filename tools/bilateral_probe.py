"""Exact-equality probe of the CUDA bilateral filter against this box's cv2.bilateralFilter (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "manga-image-translator_b200")):
    sys.path.insert(0, p)
import cv2  # noqa: E402
import numpy as np  # noqa: E402

from mit_b200 import synth  # noqa: E402
from mit_b200.engine import get_engine  # noqa: E402

eng = get_engine("cuda:0")
rng = np.random.default_rng(3)
imgs = {"random150": rng.integers(0, 256, (150, 203, 3), dtype=np.uint8)}
imgs["smooth150"] = cv2.GaussianBlur(imgs["random150"], (0, 0), 3)
imgs["smooth600"] = cv2.GaussianBlur(rng.integers(0, 256, (600, 811, 3), dtype=np.uint8), (0, 0), 5)
imgs["page2048"] = synth.make_page(0)[0]
imgs["photo1024"] = cv2.GaussianBlur(rng.integers(0, 256, (1024, 768, 3), dtype=np.uint8), (0, 0), 9)
print("cv2", cv2.__version__, "threads", cv2.getNumThreads(), "avx512", cv2.checkHardwareSupport(getattr(cv2, "CPU_AVX_512F", 13)), "avx2", cv2.checkHardwareSupport(getattr(cv2, "CPU_AVX2", 11)),
      "fma3", cv2.checkHardwareSupport(getattr(cv2, "CPU_FMA3", 12)))
for name, im in imgs.items():
    ref = cv2.bilateralFilter(im, 17, 80, 80)
    out = eng.bilateral17(im).cpu().numpy()
    d = out.astype(int) - ref.astype(int)
    bad = np.argwhere(d != 0)
    print(f"{name:10s} {im.shape} mismatching bytes {len(bad)} of {d.size}; max |d| {np.abs(d).max()}; first {bad[:6].tolist()} cols%16 {sorted(set((bad[:, 1] % 16).tolist()))[:16]}")
    cv2.setNumThreads(1)
    ref1 = cv2.bilateralFilter(im, 17, 80, 80)
    cv2.setNumThreads(-1)
    print("           single-thread cv2 equals multi-thread cv2:", bool(np.array_equal(ref, ref1)), " vs ours:", int((out != ref1).sum()))

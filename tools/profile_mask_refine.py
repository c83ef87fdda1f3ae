"""Where the time of one mask-refinement call goes (development tool): wall time per section with a device synchronise after each,
on full-size synthetic pages.  `python tools/profile_mask_refine.py [n_pages]`"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "manga-image-translator_b200")):
    sys.path.insert(0, p)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mit_b200 import mask_refinement as MR  # noqa: E402
from mit_b200 import synth  # noqa: E402

ref = MR.get_refiner("cuda:0")
eng = ref.eng
acc = {}
names = ["mitb_op_resize_linear_u8", "mitb_op_cut_rects", "mitb_op_cc_label", "mitb_op_owner_map", "mitb_op_bilateral17", "mitb_op_dense_crf",
         "mitb_op_dilate_lines", "mitb_op_dilate_se"]
orig_call = eng._call


def timed_call(fn, *a):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = orig_call(fn, *a)
    torch.cuda.synchronize()
    nm = getattr(fn, "__name__", None) or getattr(fn, "name", str(fn))
    acc[nm] = acc.get(nm, 0.0) + time.perf_counter() - t0
    return r


n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
items = []
for i in range(n + 1):
    page, boxes, _ = synth.make_page(i)
    raw = cv2.dilate(((page[..., 0] < 100) * 255).astype(np.uint8), np.ones((3, 3), np.uint8))
    items.append((page, raw, [b.astype(np.float64) for b in boxes]))
ref.refine(*items[0], 20, 3)
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in items[1:]:
    ref.refine(*it, 20, 3)
torch.cuda.synchronize()
plain = (time.perf_counter() - t0) / n
eng._call = timed_call
t0 = time.perf_counter()
for it in items[1:]:
    ref.refine(*it, 20, 3)
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / n
print(f"refine: {1e3 * plain:.1f} ms/page (plain), {1e3 * total:.1f} ms/page with a synchronise around every library call")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:32s} {1e3 * v / n:8.2f} ms/page")
print(f"  {'host + copies (remainder)':32s} {1e3 * (total - sum(acc.values()) / n):8.2f} ms/page")

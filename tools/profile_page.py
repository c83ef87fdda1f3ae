"""One device-resident page (detect + OCR + inpaint) under a profiler: warm-up pass, then a pass bracketed by
cudaProfilerStart/Stop (use `ncu --profile-from-start off ...`).  Development tool, not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "manga-image-translator_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from mit_b200 import synth  # noqa: E402
from mit_b200.pipeline import HotPath  # noqa: E402

torch.set_grad_enabled(False)
W = bench.build_weights()
hp = HotPath("cuda:0", W["dbnet"], W["ocr"], W["dictionary"], W["lama"], W["mpe"])
page, boxes, mask = synth.make_page(0)
sp = hp.stage(page, synth.make_quads(boxes), mask)
hp.run_resident(sp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
hp.run_resident(sp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches", hp.engine.launches)

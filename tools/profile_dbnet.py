"""One device-resident DBNet forward at 2048x1536 under a profiler: warm-up pass, then a pass bracketed by
cudaProfilerStart/Stop (use `ncu --profile-from-start off ...`).  Development tool, not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "manga-image-translator_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mit_b200 import synth  # noqa: E402
from mit_b200.engine import get_engine  # noqa: E402
from oracle import weights  # noqa: E402

torch.set_grad_enabled(False)
eng = get_engine("cuda:0")
eng.load_dbnet(weights.dbnet_weights())
page = torch.from_numpy(synth.make_page(0)[0]).cuda()[None]
eng.dbnet_forward(page)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.dbnet_forward(page)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches", eng.launches)

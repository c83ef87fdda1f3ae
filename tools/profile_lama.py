"""One device-resident LaMa-MPE forward at 2048x1536 under a profiler: warm-up pass, then a pass bracketed by
cudaProfilerStart/Stop (use `ncu --profile-from-start off ...`).  Development tool, not part of the product."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "manga-image-translator_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from mit_b200 import synth  # noqa: E402
from mit_b200.engine import get_engine  # noqa: E402
from mit_b200.host import mpe  # noqa: E402
from oracle import weights  # noqa: E402

torch.set_grad_enabled(False)
eng = get_engine("cuda:0")
eng.load_lama(weights.lama_weights(9), weights.mpe_weights())
page, boxes, mask = synth.make_page(0)
r, d = mpe.mpe_tables_256(((mask.astype("float32") / 255.0) >= 0.5).astype("float32"))
pg, mk = torch.from_numpy(page).cuda(), torch.from_numpy(mask).cuda()
rel, direct = torch.from_numpy(r[None]).cuda(), torch.from_numpy(d[None]).cuda()
eng.lama_infer_u8(pg, mk, rel, direct)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.lama_infer_u8(pg, mk, rel, direct)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches", eng.launches)

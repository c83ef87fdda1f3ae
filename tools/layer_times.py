"""Per-layer conv timing of one device-resident page via the library's own CUDA-event profiler
(MITB_PROFILE_LAUNCHES=1): prints the slowest conv launches with their GEMM shape and achieved TFLOP/s."""
import json
import os
import sys

os.environ["MITB_PROFILE_LAUNCHES"] = "1"
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
for _ in range(2):
    hp.run_resident(sp)
torch.cuda.synchronize()
hp.engine.profile(True)
hp.run_resident(sp)
torch.cuda.synchronize()
rep = hp.engine.profile_report()
L = rep.pop("_launches", [])
tot = sum(x[4] for x in L)
print("conv launches", len(L), "total ms", round(tot, 2))
agg = {}
for kind, m, k, n, ms in L:
    a = agg.setdefault((kind, m, k, n), [0, 0.0])
    a[0] += 1
    a[1] += ms
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(f"{'kind':16s} {'M':>8s} {'K':>6s} {'N':>6s} {'cnt':>4s} {'ms':>8s} {'TF/s':>7s}")
for (kind, m, k, n), (cnt, ms) in rows[:45]:
    print(f"{kind:16s} {m:8d} {k:6d} {n:6d} {cnt:4d} {ms:8.3f} {2.0 * m * k * n * cnt / ms / 1e9:7.1f}")
print(json.dumps({k: v for k, v in rep.items()}, indent=None)[:1500])

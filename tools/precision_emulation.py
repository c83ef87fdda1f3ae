"""Operand-precision study behind the bf16x3 choice: the CPU oracle with every dense conv / linear
operand rounded the way a tensor-core scheme would round it (products and accumulation stay fp32, like the MMA), against the
plain fp32 oracle.  Development tool: `python tools/precision_emulation.py dbnet 1024 768`, `... lama 1024 768 [blocks]`,
`... ocr 46000 647 16`.  Prints max / mean absolute output error per scheme (and OCR argmax flips)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import cases, nets, weights  # noqa: E402

torch.set_grad_enabled(False)


def r_fp16(t):
    return t.half().float()


def r_bf16(t):
    return t.bfloat16().float()


def r_tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


ident = lambda t: t  # noqa: E731
# scheme -> list of (round_a, round_w) operand pairs whose products are summed (each term = one MMA pass)
def split(r):
    return (lambda t: r(t)), (lambda t: r(t - r(t)))


bh, bm = split(r_bf16)
fh, fm = split(r_fp16)
SCHEMES = {
    "bf16 x1  (1 MMA)": [(bh, bh)],
    "tf32 x1  (2 MMA-equivalents)": [(r_tf32, r_tf32)],
    "fp16 x1  (1 MMA)": [(fh, fh)],
    "fp16 x2  (2 MMAs: A hi+lo, W hi)": [(fh, fh), (fm, fh)],
    "bf16 x3  (3 MMAs, shipped)": [(bh, bh), (bh, bm), (bm, bh)],
    "fp16 x3  (3 MMAs)": [(fh, fh), (fh, fm), (fm, fh)],
}
orig = dict(conv2d=F.conv2d, linear=F.linear, convT=F.conv_transpose2d)


class Patch:
    def __init__(self, terms):
        self.terms = terms

    def __enter__(self):
        terms = self.terms

        def conv2d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if groups != 1:
                return orig["conv2d"](x, w, b, stride, padding, dilation, groups)
            y = sum(orig["conv2d"](ra(x), rw(w), None, stride, padding, dilation, groups) for ra, rw in terms)
            return y if b is None else y + b.view(1, -1, 1, 1)

        def linear(x, w, b=None):
            y = sum(orig["linear"](ra(x), rw(w)) for ra, rw in terms)
            return y if b is None else y + b

        def convT(x, w, b=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
            y = sum(orig["convT"](ra(x), rw(w), None, stride, padding, output_padding, groups, dilation) for ra, rw in terms)
            return y if b is None else y + b.view(1, -1, 1, 1)
        F.conv2d, F.linear, F.conv_transpose2d = conv2d, linear, convT

    def __exit__(self, *a):
        F.conv2d, F.linear, F.conv_transpose2d = orig["conv2d"], orig["linear"], orig["convT"]


def main():
    which = sys.argv[1]
    if which == "dbnet":
        H, W = int(sys.argv[2]), int(sys.argv[3])
        sd = weights.dbnet_weights()
        _, x = cases.dbnet_case(H, W)
        db0, m0 = nets.dbnet_forward(sd, x)
        db0 = db0.sigmoid()
        for name, terms in SCHEMES.items():
            with Patch(terms):
                db, m = nets.dbnet_forward(sd, x)
            db = db.sigmoid()
            print(f"dbnet {H}x{W} | {name:34s} | db max {(db - db0).abs().max().item():.2e} mean {(db - db0).abs().mean().item():.2e} | "
                  f"mask max {(m - m0).abs().max().item():.2e}", flush=True)
    elif which == "lama":
        H, W = int(sys.argv[2]), int(sys.argv[3])
        nb = int(sys.argv[4]) if len(sys.argv) > 4 else 9
        sd, msd = weights.lama_weights(nb), weights.mpe_weights()
        img, m = cases.lama_case(H, W)
        rel, direct = nets.mpe_tables(m[0, 0].numpy())
        rel, direct = torch.from_numpy(rel)[None], torch.from_numpy(direct)[None]
        o0 = nets.lama_forward(sd, msd if nb == 9 else None, img, m, rel, direct)
        for name, terms in SCHEMES.items():
            with Patch(terms):
                o = nets.lama_forward(sd, msd if nb == 9 else None, img, m, rel, direct)
            print(f"lama({nb}) {H}x{W} | {name:34s} | out max {(o - o0).abs().max().item():.2e} mean {(o - o0).abs().mean().item():.2e}", flush=True)
    elif which == "ocr":
        V, wp, n = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
        sd = weights.ocr_weights(V)
        _, x = cases.ocr_case(n, wp)
        lg0, c0 = nets.ocr_forward(sd, x)
        top2 = lg0.topk(2, -1).values
        margin = top2[..., 0] - top2[..., 1]
        for name, terms in SCHEMES.items():
            with Patch(terms):
                lg, c = nets.ocr_forward(sd, x)
            flips = lg.argmax(-1) != lg0.argmax(-1)
            safe = margin > 1e-3
            print(f"ocr {n}x48x{wp} V={V} | {name:34s} | logits max {(lg - lg0).abs().max().item():.2e} | argmax flips {int(flips.sum())} of {flips.numel()} "
                  f"({int((flips & safe).sum())} where the margin > 1e-3)", flush=True)


if __name__ == "__main__":
    main()

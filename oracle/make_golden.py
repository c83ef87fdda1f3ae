"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference modules
(imported from /root/reference through oracle/refload.py) on the seeded cases of oracle/cases.py.

Run in the build container only:  python -m oracle.make_golden
The GPU box has no /root/reference; it consumes the committed fixtures.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

from . import cases, refload, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    warnings.filterwarnings("ignore")
    m = refload.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # --- detector network: DBNetConvNext + the caller's sigmoid (dbnet_convnext.py:499-509)
    sd = weights.dbnet_weights()
    net = m["det"].DBNetConvNext().eval()
    net.load_state_dict(sd)
    _, x = cases.dbnet_case()
    db, mask = net(x)
    np.savez_compressed(os.path.join(OUT, "dbnet_256.npz"), db_sigmoid=db.sigmoid().numpy(), mask=mask.numpy(),
                        db_logit0=db[:, 0].numpy())

    # --- OCR: OCR.forward + decode_ctc_top1 (model_48px_ctc.py:438-494)
    V = cases.OCR_VOCAB_SMALL
    sd = weights.ocr_weights(V)
    ocr = m["ocr"].OCR(weights.synthetic_dictionary(V), 768).eval()
    ocr.load_state_dict(sd, strict=False)
    _, x = cases.ocr_case()
    logits, colors = ocr(x)
    lp = logits.log_softmax(2)
    val, idx = lp.max(2)
    dec = ocr.decode(x, [0] * x.shape[0], 0)
    flat = np.array([[b, int(c[0]), c[1]] + list(c[2:]) for b, line in enumerate(dec) for c in line], np.float64)
    top2 = logits.topk(2, dim=-1).values
    np.savez_compressed(os.path.join(OUT, "ocr_200.npz"), idx=idx.numpy().astype(np.int32), logprob=val.numpy(),
                        colors=colors.clamp(0, 1).numpy(), decoded=flat,
                        margin=(top2[..., 0] - top2[..., 1]).numpy())

    # --- LaMa-MPE and LaMa-large: LamaFourier.__call__ incl. CPU MPE tables (inpainting_lama_mpe.py:713-815)
    img, mask = cases.lama_case()
    for name, nb, use_mpe in (("lama_mpe", 9, True), ("lama_large", 18, False)):
        lf = m["lama"].LamaFourier(build_discriminator=False, use_mpe=use_mpe, large_arch=nb == 18)
        lf.generator.load_state_dict(weights.lama_weights(nb))
        if use_mpe:
            lf.mpe.load_state_dict(weights.mpe_weights())
        lf.eval()
        out = lf(img.clone(), mask)
        extra = {}
        if use_mpe:
            rel, _, direct = lf.load_masked_position_encoding(mask[0, 0].numpy())
            extra = dict(rel_pos=rel.astype(np.int16), direct=direct.astype(np.int8))
        np.savez_compressed(os.path.join(OUT, f"{name}_128x96.npz"), out=out.numpy(), **extra)

    # --- one isolated FFC block + FourierUnit (the north-star kernel) on a non power-of-two spectrum size
    blk = m["lama"].FFCResnetBlock(512, padding_type="reflect", norm_layer=torch.nn.BatchNorm2d,
                                   activation_layer=torch.nn.ReLU, ratio_gin=0.75, ratio_gout=0.75,
                                   enable_lfu=False).eval()
    sd = weights.lama_weights(1)
    blk.load_state_dict({k[len("model.5."):]: v for k, v in sd.items() if k.startswith("model.5.")})
    rng = np.random.default_rng(14)
    xl = torch.from_numpy(rng.standard_normal((1, 128, 20, 14)).astype(np.float32))
    xg = torch.from_numpy(rng.standard_normal((1, 384, 20, 14)).astype(np.float32))
    yl, yg = blk((xl, xg))
    fu = blk.conv1.ffc.convg2g.fu
    s = torch.from_numpy(rng.standard_normal((1, 192, 20, 14)).astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "ffc_block_20x14.npz"), yl=yl.numpy(), yg=yg.numpy(), fu=fu(s).numpy())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())

"""TEST INFRASTRUCTURE -- seeded small cases shared by oracle/make_golden.py and the tests.

Inputs and weights are regenerated from seeds (numpy PCG64), only outputs live in tests/golden/.
"""
from __future__ import annotations

import numpy as np
import torch

from . import weights

OCR_VOCAB_SMALL = 512


def dbnet_case(h=256, w=256, n=1, seed=11):
    rng = np.random.default_rng(seed)
    # smooth-ish page: low-frequency field + strokes, u8 range mapped like det_batch_forward_default
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    x = img.astype(np.float32) / 127.5 - 1.0
    return img, torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def ocr_case(n=2, wp=200, seed=12):
    rng = np.random.default_rng(seed)
    img = np.full((n, 48, wp, 3), 0, np.uint8)
    for i in range(n):
        wi = wp - 135 if i == 0 else max(16, (wp - 135) * (i + 1) // (n + 1))
        line = np.clip(235 + 10 * rng.standard_normal((48, wi, 1)), 0, 255).repeat(3, 2)
        for _ in range(wi // 6):
            x0, y0 = rng.integers(0, wi - 4), rng.integers(4, 40)
            line[y0:y0 + rng.integers(2, 8), x0:x0 + rng.integers(1, 4)] = rng.integers(0, 60)
        img[i, :, :wi] = line.astype(np.uint8)
    x = (torch.from_numpy(img).float() - 127.5) / 127.5
    return img, x.permute(0, 3, 1, 2).contiguous()


def lama_case(h=128, w=96, seed=13):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 1, (1, 3, h, w)).astype(np.float32)
    mask = np.zeros((1, 1, h, w), np.float32)
    mask[:, :, h // 4: h // 4 + h // 3, w // 5: w // 5 + w // 2] = 1
    mask[:, :, (3 * h) // 4: (3 * h) // 4 + h // 8, (2 * w) // 3: (2 * w) // 3 + w // 5] = 1
    img = img * (1 - mask)
    return torch.from_numpy(img), torch.from_numpy(mask)


def all_weights():
    return dict(dbnet=weights.dbnet_weights(), ocr=weights.ocr_weights(OCR_VOCAB_SMALL),
                lama=weights.lama_weights(9), lama_large=weights.lama_weights(18), mpe=weights.mpe_weights())

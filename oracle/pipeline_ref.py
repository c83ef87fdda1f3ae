"""TEST INFRASTRUCTURE -- CPU restatement of the three plugins' ``_infer`` (reference CPU path, fp32):

  DBConvNextDetector._infer   detection/dbnet_convnext.py:541-588   (cv2.bilateralFilter on the host, torch CPU forward)
  Model48pxCTCOCR._infer      ocr/model_48px_ctc.py:62-160
  LamaMPEInpainter._infer     inpainting/inpainting_lama_mpe.py:56-118 (CPU branch: plain fp32, :93-95)

Network arithmetic comes from oracle/nets.py (pinned against the reference modules); the geometry / contour helpers are
the shared host restatements in mit_b200.host (third-party shapely/pyclipper are absent here).  Used by the plugin-level
parity tests and as the timed CPU arm of bench.py (``cpu_baseline`` / ``--impl reference``), never by the product path.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch

from . import nets

from mit_b200.host import det_post, rearrange
from mit_b200.host.geometry import Quadrilateral, generate_text_direction


def detector_infer(sd, image: np.ndarray, detect_size: int, text_threshold: float, box_threshold: float, unclip_ratio: float):
    def batch_forward(batch):
        return nets.dbnet_batch_forward(sd, np.asarray(batch))

    db, mask = rearrange.rearrange_forward(image, batch_forward, detect_size, 4)
    if db is None:
        img_resized, ratio, _, pad_w, pad_h = det_post.resize_aspect_ratio(cv2.bilateralFilter(image, 17, 80, 80), detect_size,
                                                                          cv2.INTER_LINEAR, mag_ratio=1)
        rh, rw = img_resized.shape[:2]
        ratio_h = ratio_w = 1 / ratio
        db, mask = batch_forward([img_resized])
    else:
        rh, rw = image.shape[:2]
        ratio_w = ratio_h = 1
        pad_h = pad_w = 0
    mask = mask[0, 0]
    boxes, scores = det_post.boxes_from_prob(db[0, 0], text_threshold, box_threshold, unclip_ratio, rw, rh)
    polys = det_post.polys_from_boxes(boxes, scores, ratio_w, ratio_h)
    textlines = [Quadrilateral(p.astype(int), "", s) for p, s in zip(polys, scores)]
    textlines = [q for q in textlines if q.area > 16]
    up = cv2.resize(mask, (mask.shape[1] * 2, mask.shape[0] * 2), interpolation=cv2.INTER_LINEAR)
    if pad_h > 0:
        up = up[:-pad_h, :]
    elif pad_w > 0:
        up = up[:, :-pad_w]
    return textlines, np.clip(up * 255, 0, 255).astype(np.uint8), db, mask


def ocr_infer(sd, dictionary, image: np.ndarray, textlines, prob_threshold=None):
    """Returns the surviving quads (mutated in place like the reference) in output order."""
    threshold = 0.5 if prob_threshold is None else prob_threshold
    quads = list(generate_text_direction(textlines))
    regions = [q.get_transformed_region(image, d, 48) for q, d in quads]
    perm = sorted(range(len(regions)), key=lambda i: regions[i].shape[1])
    out = []
    for s in range(0, len(perm), 16):
        indices = perm[s:s + 16]
        widths = [regions[i].shape[1] for i in indices]
        max_width = (4 * (max(widths) + 7) // 4) + 128
        canvas = np.zeros((len(indices), 48, max_width, 3), dtype=np.uint8)
        for i, idx in enumerate(indices):
            canvas[i, :, :widths[i]] = regions[idx]
        x = (torch.from_numpy(canvas).float() - 127.5) / 127.5
        idx_t, lp, col = nets.ocr_top1(sd, x.permute(0, 3, 1, 2).contiguous())
        decoded = nets.ctc_greedy(idx_t.numpy(), lp.numpy(), col.numpy())
        for i, line in enumerate(decoded):
            if not line:
                continue
            chars, lps, cols = [], [], [[] for _ in range(6)]
            for (chid, logprob, *c6) in line:
                ch = dictionary[chid]
                ch = " " if ch == "<SP>" else ch
                chars.append(ch)
                lps.append(logprob)
                if ch != " ":
                    for k in range(6):
                        cols[k].append(int(c6[k] * 255))
            prob = np.exp(sum(lps) / len(lps))
            if prob < threshold:
                continue
            q = quads[indices[i]][0]
            q.text = "".join(chars)
            q.prob = prob
            vals = [int(sum(c) / len(c)) if c else 0 for c in cols]
            q.fg_r, q.fg_g, q.fg_b, q.bg_r, q.bg_g, q.bg_b = vals
            out.append(q)
    return out


def lama_infer(sd, mpe_sd, image: np.ndarray, mask: np.ndarray, inpainting_size: int):
    img_original, mask_original = np.copy(image), np.copy(mask)
    mask_original[mask_original < 127] = 0
    mask_original[mask_original >= 127] = 1
    mask_original = mask_original[:, :, None]
    height, width, _ = image.shape
    if max(image.shape[:2]) > inpainting_size:
        r = float(inpainting_size) / max(image.shape[0], image.shape[1])
        size = (round(image.shape[1] * r), round(image.shape[0] * r))
        image = cv2.resize(image, size, interpolation=cv2.INTER_LINEAR_EXACT)
        mask = cv2.resize(mask, size, interpolation=cv2.INTER_LINEAR_EXACT)
    h, w, _ = image.shape
    new_h = h if h % 8 == 0 else h + (8 - h % 8)
    new_w = w if w % 8 == 0 else w + (8 - w % 8)
    if (new_h, new_w) != (h, w):
        image = cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
        mask = cv2.resize(mask, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    img_t = torch.from_numpy(image).permute(2, 0, 1).unsqueeze(0).float() / 255.0
    mask_t = torch.from_numpy(mask).unsqueeze(0).unsqueeze(0).float() / 255.0
    mask_t[mask_t < 0.5] = 0
    mask_t[mask_t >= 0.5] = 1
    img_t = img_t * (1 - mask_t)
    rel = direct = None
    if mpe_sd is not None:
        rel, direct = nets.mpe_tables(mask_t[0, 0].numpy())
        rel, direct = torch.from_numpy(rel)[None], torch.from_numpy(direct)[None]
    out = nets.lama_forward(sd, mpe_sd, img_t, mask_t, rel, direct)
    res = (out.squeeze(0).permute(1, 2, 0).numpy() * 255.0).astype(np.uint8)
    if (new_h, new_w) != (height, width):
        res = cv2.resize(res, (width, height), interpolation=cv2.INTER_LINEAR)
    return res * mask_original + img_original * (1 - mask_original), out

"""TEST INFRASTRUCTURE -- CPU fp32 oracle for the detect -> OCR -> inpaint hot path.

A functional (state-dict in, tensor out) restatement, in plain torch CPU fp32
ops, of the three reference networks and the small host algorithms around them.
Nothing here is product code: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it, and only
as the checker or as the timed CPU arm.  Every function cites the reference
lines (relative to /root/reference/manga_translator) it restates.

Pinning: the reference's own tests hold NO vectors for this path (SURVEY.md
§4, §8c).  The oracle is pinned instead against the reference modules executed
in the build container (tests/test_oracle_vs_reference.py, via oracle/refload.py)
and against committed outputs of those modules (tests/golden/*.npz, generated
by oracle/make_golden.py).

Layout: all tensors NCHW fp32 like the reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------- #
#  DBNet-ConvNeXt  (detection/dbnet_convnext.py)
# --------------------------------------------------------------------------- #

_LN_EPS = 1e-6   # timm.layers.LayerNorm / LayerNorm2d default (third-party, not in tree)


def _ln_channels_first(x, w, b):
    """timm LayerNorm2d: permute -> layer_norm over C -> permute (dbnet_convnext.py:17, used :152-166, :272)."""
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, _LN_EPS)
    return y.permute(0, 3, 1, 2)


def convnext_block(sd: SD, p: str, x):
    """ConvNeXtBlock.forward (dbnet_convnext.py:112-127).

    conv_dw is depthwise 7x7 when out>=in, otherwise a dense 7x7 (:100); the
    shortcut is a 1x1 conv when channels change (:106-107, Downsample :31-38).
    """
    w = sd[p + "conv_dw.weight"]
    groups = x.shape[1] if w.shape[1] == 1 else 1
    y = F.conv2d(x, w, sd[p + "conv_dw.bias"], padding=3, groups=groups)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (y.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], _LN_EPS)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)  # exact erf GELU (timm Mlp default act nn.GELU)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    y = y.permute(0, 3, 1, 2) * sd[p + "gamma"].reshape(1, -1, 1, 1)
    if p + "shortcut.conv.weight" in sd:
        x = F.conv2d(x, sd[p + "shortcut.conv.weight"], sd[p + "shortcut.conv.bias"])
    return y + x


def convnext_stage(sd: SD, p: str, x, depth: int):
    """ConvNeXtStage.forward (dbnet_convnext.py:190-193): optional LN2d + 2x2 s2 conv, then blocks."""
    if p + "downsample.1.weight" in sd:
        x = _ln_channels_first(x, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"])
        x = F.conv2d(x, sd[p + "downsample.1.weight"], sd[p + "downsample.1.bias"], stride=2)
    for k in range(depth):
        x = convnext_block(sd, f"{p}blocks.{k}.", x)
    return x


def upconv_skip(sd: SD, p: str, x):
    """UpconvSkip.forward (dbnet_convnext.py:377-380)."""
    x = convnext_block(sd, p + "conv.", x)
    return F.conv_transpose2d(x, sd[p + "upconv.weight"], sd[p + "upconv.bias"], stride=2)


def db_head(sd: SD, p: str, x):
    """DBHead.forward eval branch (dbnet_convnext.py:399-407)."""
    def branch(q, final_sigmoid):
        y = F.conv2d(x, sd[q + "0.weight"], sd.get(q + "0.bias"), padding=1)
        y = F.silu(y)
        y = F.conv_transpose2d(y, sd[q + "2.weight"], sd[q + "2.bias"], stride=2, padding=1)
        y = F.silu(y)
        y = F.conv_transpose2d(y, sd[q + "4.weight"], sd[q + "4.bias"], stride=2, padding=1)
        return torch.sigmoid(y) if final_sigmoid else y
    return torch.cat([branch(p + "binarize.", False), branch(p + "thresh.", True)], dim=1)


def dbnet_forward(sd: SD, x, taps: Optional[dict] = None):
    """DBNetConvNext.forward (dbnet_convnext.py:474-491) -> (db logits [N,2,H,W], mask [N,1,H/2,W/2])."""
    depths = (3, 3, 27, 3)
    s = F.conv2d(x, sd["backbone.stem.0.weight"], sd["backbone.stem.0.bias"], stride=4)
    s = _ln_channels_first(s, sd["backbone.stem.1.weight"], sd["backbone.stem.1.bias"])
    h4 = convnext_stage(sd, "backbone.stages.0.", s, depths[0])
    h8 = convnext_stage(sd, "backbone.stages.1.", h4, depths[1])
    h16 = convnext_stage(sd, "backbone.stages.2.", h8, depths[2])
    h32 = convnext_stage(sd, "backbone.stages.3.", h16, depths[3])
    h64 = convnext_stage(sd, "down_conv1.", h32, 2)
    h128 = convnext_stage(sd, "down_conv2.", h64, 2)
    up128 = upconv_skip(sd, "upconv1.", h128)
    up64 = upconv_skip(sd, "upconv2.", torch.cat([up128, h64], 1))
    up32 = upconv_skip(sd, "upconv3.", torch.cat([up64, h32], 1))
    up16 = upconv_skip(sd, "upconv4.", torch.cat([up32, h16], 1))
    up8 = upconv_skip(sd, "upconv5.", torch.cat([up16, h8], 1))
    up4 = upconv_skip(sd, "upconv6.", torch.cat([up8, h4], 1))
    if taps is not None:
        taps.update(stem=s, h4=h4, h8=h8, h16=h16, h32=h32, h64=h64, h128=h128, up8=up8, up4=up4)
    db = db_head(sd, "conv_db.", up8)
    m = F.silu(F.conv2d(up4, sd["conv_mask.0.weight"], sd["conv_mask.0.bias"], padding=1))
    m = F.silu(F.conv2d(m, sd["conv_mask.2.weight"], sd["conv_mask.2.bias"], padding=1))
    m = torch.sigmoid(F.conv2d(m, sd["conv_mask.4.weight"], sd["conv_mask.4.bias"]))
    return db, m


def dbnet_batch_forward(sd: SD, batch_u8_nhwc: np.ndarray):
    """det_batch_forward_default (dbnet_convnext.py:499-509): divide-then-subtract normalisation,
    forward, sigmoid on BOTH db channels (channel 1 is therefore sigmoid(sigmoid(.)))."""
    x = batch_u8_nhwc.astype(np.float32) / 127.5 - 1.0
    x = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
    db, mask = dbnet_forward(sd, x)
    return db.sigmoid().numpy(), mask.numpy()


# --------------------------------------------------------------------------- #
#  48px ResNet + Transformer CTC recogniser (ocr/model_48px_ctc.py)
# --------------------------------------------------------------------------- #

_BN_EPS = 1e-5


def _bn(sd: SD, p: str, x):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, _BN_EPS)


def _ocr_block(sd: SD, p: str, x):
    """BasicBlock.forward, pre-activation (model_48px_ctc.py:389-403)."""
    y = F.conv2d(F.relu(_bn(sd, p + "bn1.", x)), sd[p + "conv1.weight"], padding=1)
    y = F.conv2d(F.relu(_bn(sd, p + "bn2.", y)), sd[p + "conv2.weight"], padding=1)
    if p + "downsample.1.weight" in sd:
        x = F.conv2d(_bn(sd, p + "downsample.0.", x), sd[p + "downsample.1.weight"])
    return y + x


def ocr_backbone(sd: SD, x):
    """ResNet.forward (model_48px_ctc.py:336-370) -> [N,320,1,T]."""
    p = "backbone.ConvNet."
    x = F.conv2d(x, sd[p + "conv0_1.weight"], padding=1)
    x = F.relu(_bn(sd, p + "bn0_1.", x))
    x = F.conv2d(x, sd[p + "conv0_2.weight"], padding=1)
    x = F.avg_pool2d(x, 2, 2)
    for k in range(4):
        x = _ocr_block(sd, f"{p}layer1.{k}.", x)
    x = F.conv2d(F.relu(_bn(sd, p + "bn1.", x)), sd[p + "conv1.weight"], padding=1)
    x = F.avg_pool2d(x, 2, 2)
    for k in range(6):
        x = _ocr_block(sd, f"{p}layer2.{k}.", x)
    x = F.conv2d(F.relu(_bn(sd, p + "bn2.", x)), sd[p + "conv2.weight"], padding=1)
    x = F.avg_pool2d(x, kernel_size=2, stride=(2, 1), padding=(0, 1))  # zero columns are averaged in
    for k in range(8):
        x = _ocr_block(sd, f"{p}layer3.{k}.", x)
    x = F.conv2d(F.relu(_bn(sd, p + "bn3.", x)), sd[p + "conv3.weight"], padding=1)
    for k in range(6):
        x = _ocr_block(sd, f"{p}layer4.{k}.", x)
    x = F.conv2d(F.relu(_bn(sd, p + "bn4_1.", x)), sd[p + "conv4_1.weight"], stride=(2, 1), padding=1)
    x = F.conv2d(F.relu(_bn(sd, p + "bn4_2.", x)), sd[p + "conv4_2.weight"], padding=0)
    return _bn(sd, p + "bn4_3.", x)


def sinusoid_pe(T: int, d: int = 320):
    """PositionalEncoding table (model_48px_ctc.py:163-178), recomputed, never loaded (:45-47)."""
    pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe = torch.zeros(T, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def ocr_encoder_layer(sd: SD, p: str, x, heads: int = 8):
    """CustomTransformerEncoderLayer.forward, norm_first (model_48px_ctc.py:253-274): positional
    encoding is added to q and k only, there is no padding mask, eps 1e-5."""
    N, T, D = x.shape
    hd = D // heads
    z = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    zp = z + sinusoid_pe(T, D).to(z.device)
    W, B = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
    q = F.linear(zp, W[:D], B[:D]).view(N, T, heads, hd).transpose(1, 2)
    k = F.linear(zp, W[D:2 * D], B[D:2 * D]).view(N, T, heads, hd).transpose(1, 2)
    v = F.linear(z, W[2 * D:], B[2 * D:]).view(N, T, heads, hd).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1) @ v
    a = a.transpose(1, 2).reshape(N, T, D)
    x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    z = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    z = F.linear(F.gelu(F.linear(z, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return x + z


def ocr_forward(sd: SD, img):
    """OCR.forward (model_48px_ctc.py:438-445) -> (char logits [N,T,V], colour values [N,T,6])."""
    f = ocr_backbone(sd, img).squeeze(2).permute(0, 2, 1)
    for i in range(3):
        f = ocr_encoder_layer(sd, f"encoders.layers.{i}.", f)
    z = F.gelu(F.layer_norm(f, (f.shape[-1],), sd["char_pred_norm.0.weight"], sd["char_pred_norm.0.bias"], 1e-5))
    logits = F.linear(z, sd["char_pred.weight"], sd["char_pred.bias"])
    colors = F.linear(f, sd["color_pred1.0.weight"], sd["color_pred1.0.bias"])
    return logits, colors


def ocr_top1(sd: SD, img):
    """First half of decode_ctc_top1 (model_48px_ctc.py:460-463): per timestep argmax, its log-prob,
    colours clamped to [0,1]."""
    logits, colors = ocr_forward(sd, img)
    lp = logits.log_softmax(2)
    val, idx = lp.max(2)
    return idx.to(torch.int32), val, colors.clamp(0, 1)


def ctc_greedy(idx: np.ndarray, logprob: np.ndarray, colors: np.ndarray, blank: int = 0):
    """Second half of decode_ctc_top1 (model_48px_ctc.py:464-493): collapse repeats and blanks over ALL
    timesteps (padding included); returns per line a list of (chid, logprob, 6 colours)."""
    out = []
    for b in range(idx.shape[0]):
        line, last = [], blank
        for t in range(idx.shape[1]):
            c = int(idx[b, t])
            if c != last and c != blank:
                line.append((c, float(logprob[b, t])) + tuple(float(v) for v in colors[b, t]))
            last = c
        out.append(line)
    return out


# --------------------------------------------------------------------------- #
#  LaMa FFC generator (inpainting/inpainting_lama_mpe.py)
# --------------------------------------------------------------------------- #

def _conv_reflect(x, w, stride=1, pad=1, bias=None):
    """nn.Conv2d(padding_mode='reflect') as used by FFC (inpainting_lama_mpe.py:333-340)."""
    if pad:
        x = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    return F.conv2d(x, w, bias, stride=stride)


def fourier_unit(sd: SD, p: str, x):
    """FourierUnit.forward (inpainting_lama_mpe.py:214-257): ortho rfft2, (c,re/im) channel
    interleave, 1x1 conv + BN + ReLU, ortho irfft2 back to the input size."""
    n, c, h, w = x.shape
    x = x.float()  # :225-226 -- the FFT runs in fp32 even under autocast (no-op on the CPU fp32 path)
    f = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")
    f = torch.stack((f.real, f.imag), dim=2).reshape(n, 2 * c, h, w // 2 + 1)
    f = F.relu(_bn(sd, p + "bn.", F.conv2d(f, sd[p + "conv_layer.weight"])))
    f = f.reshape(n, -1, 2, h, w // 2 + 1).float()  # :247-248
    f = torch.complex(f[:, :, 0].contiguous(), f[:, :, 1].contiguous())
    return torch.fft.irfftn(f, s=(h, w), dim=(-2, -1), norm="ortho")


def spectral_transform(sd: SD, p: str, x):
    """SpectralTransform.forward with enable_lfu=False, stride 1 (inpainting_lama_mpe.py:286-307)."""
    s = F.relu(_bn(sd, p + "conv1.1.", F.conv2d(x, sd[p + "conv1.0.weight"])))
    return F.conv2d(s + fourier_unit(sd, p + "fu.", s), sd[p + "conv2.weight"])


def ffc_bn_act(sd: SD, p: str, x_l, x_g, stride=1, pad=1):
    """FFC.forward + FFC_BN_ACT.forward, ungated (inpainting_lama_mpe.py:349-369, 394-399).
    x_g is None when the layer has no global input; returns (y_l, y_g or None)."""
    f = p + "ffc."
    y_l = _conv_reflect(x_l, sd[f + "convl2l.weight"], stride, pad)
    if x_g is not None:
        y_l = y_l + _conv_reflect(x_g, sd[f + "convg2l.weight"], stride, pad)
    y_l = F.relu(_bn(sd, p + "bn_l.", y_l))
    y_g = None
    if f + "convl2g.weight" in sd:
        y_g = _conv_reflect(x_l, sd[f + "convl2g.weight"], stride, pad)
        if x_g is not None:
            y_g = y_g + spectral_transform(sd, f + "convg2g.", x_g)
        y_g = F.relu(_bn(sd, p + "bn_g.", y_g))
    return y_l, y_g


def mpe_embed(mpe_sd: SD, rel_pos: torch.Tensor, direct: torch.Tensor):
    """MPE.forward (inpainting_lama_mpe.py:625-632): rel_pos int [B,H,W] -> table lookup * alpha5;
    direct {0,1} [B,H,W,4] @ W[4,64] * alpha6; both returned as [B,64,H,W]."""
    e = mpe_sd["rel_pos_emb.weight"][rel_pos.long()].permute(0, 3, 1, 2) * mpe_sd["alpha5"]
    d = (direct.to(torch.float32) @ mpe_sd["direct_emb.weight"]).permute(0, 3, 1, 2) * mpe_sd["alpha6"]
    return e, d


def lama_n_blocks(sd: SD) -> int:
    n = 0
    while f"model.{5 + n}.conv1.ffc.convl2l.weight" in sd:
        n += 1
    return n


def lama_generator(sd: SD, img, mask, rel_pos_emb=None, direct_emb=None, taps: Optional[dict] = None):
    """FFCResNetGenerator.forward (inpainting_lama_mpe.py:603-613) for the LamaFourier configuration
    (:645-659): 4->64 7x7 reflect stem, 3 stride-2 downs, n FFC res-blocks, 3 ConvT ups, 7x7 out, sigmoid."""
    nb = lama_n_blocks(sd)
    x = torch.cat([img * (1 - mask), mask], 1)
    x = F.pad(x, (3, 3, 3, 3), mode="reflect")
    x_l, _ = ffc_bn_act(sd, "model.1.", x, None, 1, 0)
    if rel_pos_emb is not None:
        x_l = x_l + rel_pos_emb
        x_l = x_l + direct_emb
    x_l, _ = ffc_bn_act(sd, "model.2.", x_l, None, 2, 1)
    x_l, _ = ffc_bn_act(sd, "model.3.", x_l, None, 2, 1)
    x_l, x_g = ffc_bn_act(sd, "model.4.", x_l, None, 2, 1)
    if taps is not None:
        taps.update(down_l=x_l, down_g=x_g)
    for b in range(nb):
        p = f"model.{5 + b}."
        y_l, y_g = ffc_bn_act(sd, p + "conv1.", x_l, x_g)
        y_l, y_g = ffc_bn_act(sd, p + "conv2.", y_l, y_g)
        x_l, x_g = x_l + y_l, x_g + y_g
    x = torch.cat([x_l, x_g], 1)
    if taps is not None:
        taps.update(bottleneck=x)
    k = 5 + nb + 1  # ConcatTupleLayer occupies one index
    for _ in range(3):
        x = F.conv_transpose2d(x, sd[f"model.{k}.weight"], sd[f"model.{k}.bias"], stride=2, padding=1,
                               output_padding=1)
        x = F.relu(_bn(sd, f"model.{k + 1}.", x))
        k += 3
    x = F.pad(x, (3, 3, 3, 3), mode="reflect")
    x = F.conv2d(x, sd[f"model.{k + 1}.weight"], sd[f"model.{k + 1}.bias"])
    return torch.sigmoid(x)


def lama_forward(sd: SD, mpe_sd: Optional[SD], img, mask, rel_pos=None, direct=None):
    """LamaFourier.__call__ in inpaint_only mode (inpainting_lama_mpe.py:713-726).  ``img`` must already be
    pre-masked (img *= 1-mask, :92).  rel_pos/direct are the integer tables of
    ``mpe_tables`` (None for lama_large, which has no MPE)."""
    e = d = None
    if mpe_sd is not None:
        e, d = mpe_embed(mpe_sd, rel_pos, direct)
    pred = lama_generator(sd, img, mask, e, d)
    return pred * mask + (1 - mask) * img


def mpe_tables(mask01: np.ndarray):
    """LamaFourier.load_masked_position_encoding (inpainting_lama_mpe.py:751-815), restated with an
    explicit binary dilation instead of cv2.filter2D on float images.

    mask01: float/uint8 [H,W] with 1 inside the hole.  Returns (rel_pos int32 [H,W], direct int32 [H,W,4]).
    Semantics: work on a 256x256 INTER_AREA-downsampled mask (any coverage >0 counts as hole); ``known`` =
    non-hole pixels; repeatedly dilate ``known`` by a 3x3 box (cv2 BORDER_REFLECT_101 borders); a pixel first
    covered at iteration i gets pos=i; four 2x2-corner kernels flag from which diagonal side the front arrived.
    """
    import cv2
    m = (np.asarray(mask01, dtype=np.float32) * 255).astype(np.uint8)
    H, W = m.shape
    small = cv2.resize(m, (256, 256), interpolation=cv2.INTER_AREA)
    known = small == 0
    pos = np.zeros((256, 256), np.int32)
    direct = np.zeros((256, 256, 4), np.int32)
    # kernel supports as (dy, dx) offsets; filter2D correlates: out(y,x)=sum k(i,j) src(y+i-1,x+j-1)
    box = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    corners = [
        [(-1, -1), (-1, 0), (0, -1), (0, 0)],   # d_filter1: rows 0-1, cols 0-1
        [(0, -1), (0, 0), (1, -1), (1, 0)],     # d_filter2: rows 1-2, cols 0-1
        [(-1, 0), (-1, 1), (0, 0), (0, 1)],     # d_filter3: rows 0-1, cols 1-2
        [(0, 0), (0, 1), (1, 0), (1, 1)],       # d_filter4: rows 1-2, cols 1-2
    ]

    def dilate(k, offs):
        p = np.pad(k, 1, mode="reflect")
        out = np.zeros_like(k)
        for dy, dx in offs:
            out |= p[1 + dy:257 + dy, 1 + dx:257 + dx]
        return out

    i = 0
    if known.any():
        while not known.all():
            i += 1
            grown = dilate(known, box)
            pos[grown & ~known] = i
            for c, offs in enumerate(corners):
                direct[dilate(known, offs) & ~known, c] = 1
            known = grown
    rel = np.clip((pos / 128.0 * 128).astype(np.int32), 0, 127)
    if (H, W) != (256, 256):
        rel = cv2.resize(rel, (W, H), interpolation=cv2.INTER_NEAREST)
        direct = cv2.resize(direct, (W, H), interpolation=cv2.INTER_NEAREST)
        hole = (m.astype(np.float64) / 255) != 0
        rel = rel.copy()
        direct = direct.copy()
        rel[~hole] = 0
        direct[~hole, :] = 0
    return rel.astype(np.int32), direct.astype(np.int32)

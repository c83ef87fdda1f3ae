"""TEST INFRASTRUCTURE -- seeded "hardened" random state_dicts for the three hot-path networks.

No checkpoint exists offline, so parity is checked with random weights that share one
state_dict between the reference modules, the oracle and the CUDA path.  Key names / shapes follow
SURVEY.md Appendix C (verified against the reference modules in tests/test_oracle_vs_reference.py).

"Hardened" = drawn so that a bug cannot hide: BatchNorm running stats / affine away from (0,1,1,0), ConvNeXt
layer-scale gamma O(1) instead of 1e-6, MPE alphas 0.5 instead of 0, fan-in scaled weights so activations stay
O(1) through ~80 layers, vocabulary head scaled for healthy top-1 margins.  numpy PCG64 streams keep the values
identical on every machine.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]


# ----------------------------------------------------------------------------- specs
def _bn(spec: Spec, p: str, c: int):
    spec += [(p + "weight", (c,), "bn_w"), (p + "bias", (c,), "bn_b"),
             (p + "running_mean", (c,), "bn_m"), (p + "running_var", (c,), "bn_v")]


def _cnblock(spec: Spec, p: str, cin: int, cout: int, gamma_kind="gamma"):
    dense = cout < cin
    spec += [(p + "gamma", (cout,), gamma_kind),
             (p + "conv_dw.weight", (cout, cin if dense else 1, 7, 7), "conv"),
             (p + "conv_dw.bias", (cout,), "bias"),
             (p + "norm.weight", (cout,), "ln_w"), (p + "norm.bias", (cout,), "ln_b"),
             (p + "mlp.fc1.weight", (4 * cout, cout), "lin_act"), (p + "mlp.fc1.bias", (4 * cout,), "bias"),
             (p + "mlp.fc2.weight", (cout, 4 * cout), "lin"), (p + "mlp.fc2.bias", (cout,), "bias")]
    if cin != cout:
        spec += [(p + "shortcut.conv.weight", (cout, cin, 1, 1), "lin"), (p + "shortcut.conv.bias", (cout,), "bias")]


def dbnet_spec() -> Spec:
    s: Spec = []
    dims, depths = (128, 256, 512, 1024), (3, 3, 27, 3)
    s += [("backbone.stem.0.weight", (128, 3, 4, 4), "lin"), ("backbone.stem.0.bias", (128,), "bias"),
          ("backbone.stem.1.weight", (128,), "ln_w"), ("backbone.stem.1.bias", (128,), "ln_b")]
    prev = 128
    for i in range(4):
        p = f"backbone.stages.{i}."
        if i > 0:
            s += [(p + "downsample.0.weight", (prev,), "ln_w"), (p + "downsample.0.bias", (prev,), "ln_b"),
                  (p + "downsample.1.weight", (dims[i], prev, 2, 2), "lin"), (p + "downsample.1.bias", (dims[i],), "bias")]
        for k in range(depths[i]):
            _cnblock(s, f"{p}blocks.{k}.", dims[i], dims[i])
        prev = dims[i]
    for name in ("down_conv1.", "down_conv2."):
        s += [(name + "downsample.0.weight", (1024,), "ln_w"), (name + "downsample.0.bias", (1024,), "ln_b"),
              (name + "downsample.1.weight", (1024, 1024, 2, 2), "lin"), (name + "downsample.1.bias", (1024,), "bias")]
        for k in range(2):
            _cnblock(s, f"{name}blocks.{k}.", 1024, 1024)
    for n, (cin, cout) in enumerate([(1024, 128), (1152, 128), (1152, 128), (640, 128), (384, 128), (256, 64)], 1):
        p = f"upconv{n}."
        _cnblock(s, p + "conv.", cin, cout)
        s += [(p + "upconv.weight", (cout, cout, 2, 2), "convT2"), (p + "upconv.bias", (cout,), "bias")]
    for br, has_bias in (("binarize", True), ("thresh", False)):
        p = f"conv_db.{br}."
        s += [(p + "0.weight", (32, 128, 3, 3), "conv_act")]
        if has_bias:
            s += [(p + "0.bias", (32,), "bias")]
        s += [(p + "2.weight", (32, 32, 4, 4), "convT4_act"), (p + "2.bias", (32,), "bias"),
              (p + "4.weight", (32, 1, 4, 4), "convT4_out"), (p + "4.bias", (1,), "bias")]
    s += [("conv_mask.0.weight", (64, 64, 3, 3), "conv_act"), ("conv_mask.0.bias", (64,), "bias"),
          ("conv_mask.2.weight", (32, 64, 3, 3), "conv_act"), ("conv_mask.2.bias", (32,), "bias"),
          ("conv_mask.4.weight", (1, 32, 1, 1), "out"), ("conv_mask.4.bias", (1,), "bias")]
    return s


def ocr_spec(vocab: int) -> Spec:
    s: Spec = []
    p = "backbone.ConvNet."
    s += [(p + "conv0_1.weight", (40, 3, 3, 3), "conv")]
    _bn(s, p + "bn0_1.", 40)
    s += [(p + "conv0_2.weight", (40, 40, 3, 3), "conv_act")]
    inpl = 40
    for L, (blocks, c) in enumerate([(4, 80), (6, 160), (8, 320), (6, 320)], 1):
        for k in range(blocks):
            q = f"{p}layer{L}.{k}."
            cin = inpl if k == 0 else c
            _bn(s, q + "bn1.", cin)
            s += [(q + "conv1.weight", (c, cin, 3, 3), "conv_act")]
            _bn(s, q + "bn2.", c)
            s += [(q + "conv2.weight", (c, c, 3, 3), "conv_res")]
            if k == 0 and cin != c:
                _bn(s, q + "downsample.0.", cin)
                s += [(q + "downsample.1.weight", (c, cin, 1, 1), "lin")]
        inpl = c
        if L < 4:
            _bn(s, f"{p}bn{L}.", c)
            s += [(f"{p}conv{L}.weight", (c, c, 3, 3), "conv_act")]
    _bn(s, p + "bn4_1.", 320)
    s += [(p + "conv4_1.weight", (320, 320, 3, 3), "conv_act")]
    _bn(s, p + "bn4_2.", 320)
    s += [(p + "conv4_2.weight", (320, 320, 3, 3), "conv_act")]
    _bn(s, p + "bn4_3.", 320)
    for i in range(3):
        q = f"encoders.layers.{i}."
        s += [(q + "self_attn.in_proj_weight", (960, 320), "attn_in"), (q + "self_attn.in_proj_bias", (960,), "bias"),
              (q + "self_attn.out_proj.weight", (320, 320), "lin_res"), (q + "self_attn.out_proj.bias", (320,), "bias"),
              (q + "linear1.weight", (1280, 320), "lin_act"), (q + "linear1.bias", (1280,), "bias"),
              (q + "linear2.weight", (320, 1280), "lin_res"), (q + "linear2.bias", (320,), "bias"),
              (q + "norm1.weight", (320,), "ln_w"), (q + "norm1.bias", (320,), "ln_b"),
              (q + "norm2.weight", (320,), "ln_w"), (q + "norm2.bias", (320,), "ln_b")]
    s += [("char_pred_norm.0.weight", (320,), "ln_w"), ("char_pred_norm.0.bias", (320,), "ln_b"),
          ("char_pred.weight", (vocab, 320), "vocab"), ("char_pred.bias", (vocab,), "bias"),
          ("color_pred1.0.weight", (6, 320), "color"), ("color_pred1.0.bias", (6,), "color_b")]
    return s


def lama_spec(n_blocks: int = 9) -> Spec:
    s: Spec = []
    s += [("model.1.ffc.convl2l.weight", (64, 4, 7, 7), "conv_act")]
    _bn(s, "model.1.bn_l.", 64)
    s += [("model.2.ffc.convl2l.weight", (128, 64, 3, 3), "conv_act")]
    _bn(s, "model.2.bn_l.", 128)
    s += [("model.3.ffc.convl2l.weight", (256, 128, 3, 3), "conv_act")]
    _bn(s, "model.3.bn_l.", 256)
    s += [("model.4.ffc.convl2l.weight", (128, 256, 3, 3), "conv_act"),
          ("model.4.ffc.convl2g.weight", (384, 256, 3, 3), "conv_act")]
    _bn(s, "model.4.bn_l.", 128)
    _bn(s, "model.4.bn_g.", 384)
    for b in range(n_blocks):
        for cv in ("conv1.", "conv2."):
            p = f"model.{5 + b}.{cv}"
            f = p + "ffc."
            s += [(f + "convl2l.weight", (128, 128, 3, 3), "ffc_half"), (f + "convl2g.weight", (384, 128, 3, 3), "ffc_half"),
                  (f + "convg2l.weight", (128, 384, 3, 3), "ffc_half"),
                  (f + "convg2g.conv1.0.weight", (192, 384, 1, 1), "conv_act")]
            _bn(s, f + "convg2g.conv1.1.", 192)
            s += [(f + "convg2g.fu.conv_layer.weight", (384, 384, 1, 1), "conv_act")]
            _bn(s, f + "convg2g.fu.bn.", 384)
            s += [(f + "convg2g.conv2.weight", (384, 192, 1, 1), "ffc_half")]
            _bn(s, p + "bn_l.", 128)
            _bn(s, p + "bn_g.", 384)
    k = 5 + n_blocks + 1
    for cin, cout in ((512, 256), (256, 128), (128, 64)):
        s += [(f"model.{k}.weight", (cin, cout, 3, 3), "convT3_act"), (f"model.{k}.bias", (cout,), "bias")]
        _bn(s, f"model.{k + 1}.", cout)
        k += 3
    s += [(f"model.{k + 1}.weight", (3, 64, 7, 7), "out"), (f"model.{k + 1}.bias", (3,), "bias")]
    return s


def mpe_spec() -> Spec:
    return [("rel_pos_emb.weight", (128, 64), "sin_table"), ("direct_emb.weight", (4, 64), "normal"),
            ("alpha5", (), "alpha"), ("alpha6", (), "alpha")]


# ----------------------------------------------------------------------------- generation
def _fan_in(shape, kind):
    if kind.startswith("convT"):
        # ConvTranspose2d weight [Cin, Cout, kh, kw]: each output pixel sees Cin * (kh*kw / stride^2) taps
        return shape[0] * shape[2] * shape[3] / 4.0
    n = 1
    for d in shape[1:]:
        n *= d
    return n


def _sin_table(n_pos=128, dim=64):
    # MaskedSinusoidalPositionalEmbedding._init_weight (inpainting_lama_mpe.py:446-461)
    pe = np.array([[pos / np.power(10000, 2 * (j // 2) / dim) for j in range(dim)] for pos in range(n_pos)])
    out = np.zeros((n_pos, dim), np.float32)
    out[:, :dim // 2] = np.sin(pe[:, 0::2]).astype(np.float32)
    out[:, dim // 2:] = np.cos(pe[:, 1::2]).astype(np.float32)
    return out


def generate(spec: Spec, seed: int) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in spec:
        if kind in ("conv", "lin", "convT2", "normal_fan"):
            a = rng.standard_normal(shape) * math.sqrt(1.0 / _fan_in(shape, kind))
        elif kind in ("conv_act", "lin_act", "convT4_act", "convT3_act"):
            a = rng.standard_normal(shape) * math.sqrt(2.0 / _fan_in(shape, kind))       # He: followed by ReLU/SiLU/GELU
        elif kind in ("conv_res", "lin_res"):
            a = rng.standard_normal(shape) * math.sqrt(0.25 / _fan_in(shape, kind))      # residual branch, keep growth tame
        elif kind == "ffc_half":
            a = rng.standard_normal(shape) * math.sqrt(0.15 / _fan_in(shape, kind))      # summed branches inside x + f(x)
        elif kind == "attn_in":
            a = rng.standard_normal(shape) * math.sqrt(2.0 / _fan_in(shape, kind))
        elif kind in ("out", "convT4_out"):
            a = rng.standard_normal(shape) * math.sqrt(1.0 / _fan_in(shape, kind))       # keep the final sigmoid out of saturation
        elif kind == "vocab":
            a = rng.standard_normal(shape) * (6.0 / math.sqrt(shape[1]))                 # logit std ~6 -> clear top-1
        elif kind == "color":
            a = rng.standard_normal(shape) * (0.02 / math.sqrt(shape[1]))
        elif kind == "color_b":
            a = 0.5 + 0.1 * rng.standard_normal(shape)
        elif kind == "bias":
            a = 0.1 * rng.standard_normal(shape)
        elif kind in ("ln_w", "bn_w"):
            a = rng.uniform(0.5, 1.5, shape)
        elif kind in ("ln_b", "bn_b", "bn_m"):
            a = 0.1 * rng.standard_normal(shape)
        elif kind == "bn_v":
            a = rng.uniform(0.5, 1.5, shape)
        elif kind == "gamma":
            a = rng.uniform(0.3, 1.0, shape)
        elif kind == "sin_table":
            a = _sin_table(*shape)
        elif kind == "normal":
            a = rng.standard_normal(shape)
        elif kind == "alpha":
            a = np.full(shape, 0.5)
        else:
            raise KeyError(kind)
        sd[name] = torch.from_numpy(np.array(a, dtype=np.float32).reshape(shape).copy())
    return sd


def dbnet_weights(seed: int = 1):
    return generate(dbnet_spec(), 1000 + seed)


def ocr_weights(vocab: int, seed: int = 1):
    return generate(ocr_spec(vocab), 2000 + seed)


def lama_weights(n_blocks: int = 9, seed: int = 1):
    return generate(lama_spec(n_blocks), 3000 + seed)


def mpe_weights(seed: int = 1):
    return generate(mpe_spec(), 4000 + seed)


def synthetic_dictionary(vocab: int):
    """Stand-in for alphabet-all-v5.txt (absent offline): index 0 is the CTC blank, '<SP>' maps to a space."""
    d = ["<S>", "</S>", "<SP>"]
    cp = 0x3041
    while len(d) < vocab:
        d.append(chr(cp))
        cp += 1
        if 0xD800 <= cp <= 0xDFFF:
            cp = 0xE000
    return d[:vocab]

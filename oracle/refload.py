"""TEST INFRASTRUCTURE -- not product code.

Imports the *unmodified* reference hot-path modules from /root/reference so the
oracle restatement in ``oracle/nets.py`` can be pinned against them (SURVEY.md
Appendix B).  Only usable in the build container: the GPU box has no
/root/reference, therefore nothing under ``-m gpu``, ``smoke()`` or ``bench.py``
may import this file.  Callers must check ``available()`` first.

The reference package cannot be imported as a package (its ``__init__`` pulls
colorama/omegaconf/shapely/... which are not installed), so we
  1. register path-only package objects for ``manga_translator`` and its
     ``detection`` / ``ocr`` / ``inpainting`` sub-packages,
  2. stub the missing third-party modules with inert placeholders,
  3. provide a functional ``timm.layers`` shim with the semantics the reference
     relies on (``dbnet_convnext.py:17-18``): LayerNorm/LayerNorm2d eps=1e-6,
     Mlp = fc1 -> exact GELU -> fc2, create_conv2d with symmetric padding.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MITB_REFERENCE_ROOT", "/root/reference")
_PKG = "manga_translator"
_loaded = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, _PKG))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_timm_shim():
    try:
        import timm.layers  # noqa: F401  (real timm wins when present)
        return
    except Exception:
        pass
    import torch.nn as nn
    import torch.nn.functional as F
    from itertools import repeat

    class LayerNorm(nn.LayerNorm):
        def __init__(self, num_channels, eps=1e-6, affine=True):
            super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    class LayerNorm2d(nn.LayerNorm):
        def __init__(self, num_channels, eps=1e-6, affine=True):
            super().__init__(num_channels, eps=eps, elementwise_affine=affine)

        def forward(self, x):
            x = x.permute(0, 2, 3, 1)
            x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
            return x.permute(0, 3, 1, 2)

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None,
                     act_layer=nn.GELU, bias=True, drop=0.0, use_conv=False, **_):
            super().__init__()
            assert not use_conv
            self.fc1 = nn.Linear(in_features, hidden_features or in_features, bias=bias)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features, bias=bias)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    def to_ntuple(n):
        def parse(x):
            if isinstance(x, (tuple, list)):
                return tuple(x)
            return tuple(repeat(x, n))
        return parse

    def create_conv2d(in_chs, out_chs, kernel_size, stride=1, dilation=1, padding='',
                      depthwise=False, bias=True, **_):
        if isinstance(padding, str):
            padding = ((stride - 1) + dilation * (kernel_size - 1)) // 2 if padding != 'valid' else 0
        return nn.Conv2d(in_chs, out_chs, kernel_size, stride=stride, padding=padding, dilation=dilation,
                         groups=in_chs if depthwise else 1, bias=bias)

    def get_act_layer(name):
        if not isinstance(name, str):
            return name
        return {"gelu": nn.GELU, "relu": nn.ReLU, "silu": nn.SiLU}[name]

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    def _unused(*a, **k):
        raise NotImplementedError("timm shim: not on the hot path")

    layers = _stub("timm.layers", trunc_normal_=nn.init.trunc_normal_, AvgPool2dSame=_unused, DropPath=DropPath,
                   Mlp=Mlp, GlobalResponseNormMlp=_unused, LayerNorm2d=LayerNorm2d, LayerNorm=LayerNorm,
                   create_conv2d=create_conv2d, get_act_layer=get_act_layer, make_divisible=_unused,
                   to_ntuple=to_ntuple)
    timm = _stub("timm", layers=layers)
    timm.__path__ = []  # mark as package


def _install_third_party_stubs():
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, item):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    for name, attrs in [
        ("omegaconf", dict(OmegaConf=_Any())),
        ("colorama", dict(init=lambda *a, **k: None, Fore=_Any(), Style=_Any(), Back=_Any())),
        ("py3langid", dict(classify=lambda *a, **k: ("en", 0.0))),
        ("pyclipper", dict()),
        ("skimage", dict(io=_Any())),
        ("dotenv", dict(load_dotenv=lambda *a, **k: None)),
        ("langcodes", dict()),
        ("freetype", dict()),
    ]:
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name, **attrs)
    try:
        importlib.import_module("shapely.geometry")
    except Exception:
        geom = _stub("shapely.geometry", Polygon=_Any, MultiPoint=_Any)
        sh = _stub("shapely", affinity=_Any(), geometry=geom)
        sh.__path__ = []


def _pkg(name: str, path: str):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def load():
    """Returns a dict with the reference modules: {'det','ocr','lama','utils','config'}."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _install_third_party_stubs()
    _install_timm_shim()
    base = os.path.join(REF_ROOT, _PKG)
    if _PKG not in sys.modules or not getattr(sys.modules[_PKG], "__path__", None):
        _pkg(_PKG, base)
    config = importlib.import_module(_PKG + ".config")
    utils = importlib.import_module(_PKG + ".utils")
    for sub in ("detection", "ocr", "inpainting"):
        _pkg(f"{_PKG}.{sub}", os.path.join(base, sub))
    _loaded["config"] = config
    _loaded["utils"] = utils
    _loaded["det"] = importlib.import_module(_PKG + ".detection.dbnet_convnext")
    _loaded["ocr"] = importlib.import_module(_PKG + ".ocr.model_48px_ctc")
    _loaded["lama"] = importlib.import_module(_PKG + ".inpainting.inpainting_lama_mpe")
    return _loaded

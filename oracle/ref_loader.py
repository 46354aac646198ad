"""Import the *real* reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by ``tests/golden/make_golden.py`` (to generate the committed
fixtures) and by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is
absent, i.e. on the GPU box).  The reference needs timm / clip / mmcv / mmaction, none of
which is installed; they are only used for ``DropPath``, ``trunc_normal_``, the registry
decorator and a logger, so ~30 lines of stand-ins are inserted into ``sys.modules``
(recipe of SURVEY.md §8c / Appendix C).  No reference source is copied into this repo.
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

import torch.nn as nn

_SUB = "mmaction/models/backbones/"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the mounted reference tree (build container), else the git-ignored copy of the two backbone files that
# __graft_entry__.build() leaves under baseline/_ref/ so that bench.py's reference arm can run on the GPU box
_CANDIDATES = [os.path.join(os.environ.get("AIM_REFERENCE", "/root/reference"), _SUB), os.path.join(_REPO, "baseline", "_ref", _SUB)]
REF_ROOT = next((c for c in _CANDIDATES if os.path.isfile(c + "vitclip_aim.py")), _CANDIDATES[0])


def available() -> bool:
    return os.path.isfile(REF_ROOT + "vitclip_aim.py")


def source() -> str:
    return "mounted reference tree" if REF_ROOT == _CANDIDATES[0] else "baseline/_ref copy"


class _DropPath(nn.Module):
    """timm 0.5.4 DropPath semantics (external dependency of the reference)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0:
            m.div_(keep)
        return x * m


class _Registry:
    def __init__(self):
        self.d = {}

    def register_module(self, *a, **k):
        return lambda c: self.d.setdefault(c.__name__, c)

    def build(self, cfg):
        cfg = dict(cfg)
        return self.d[cfg.pop("type")](**cfg)


_loaded = {}


def _install_stubs():
    if "mmaction.models.builder" in sys.modules and hasattr(sys.modules["mmaction.models.builder"], "_aimb200_stub"):
        return
    tl = types.ModuleType("timm.models.layers")
    tl.DropPath = _DropPath
    tl.to_2tuple = lambda v: (v, v)
    tl.trunc_normal_ = nn.init.trunc_normal_
    mb = types.ModuleType("mmaction.models.builder")
    mb.BACKBONES = _Registry()
    mb._aimb200_stub = True
    mu = types.ModuleType("mmaction.utils")
    mu.get_root_logger = lambda *a, **k: logging.getLogger("mmaction")
    for n in ("timm", "timm.models", "clip", "mmaction", "mmaction.models", "mmaction.models.backbones", "turtle"):   # vit_imagenet.py:2
        m = types.ModuleType(n)
        m.__path__ = []
        sys.modules[n] = m
    sys.modules["turtle"].forward = None          # vit_imagenet.py:2 `from turtle import forward` (unused there)
    sys.modules.update({"timm.models.layers": tl, "mmaction.models.builder": mb, "mmaction.utils": mu})


def load(fname: str):
    if fname in _loaded:
        return _loaded[fname]
    _install_stubs()
    name = "mmaction.models.backbones." + fname[:-3]
    spec = importlib.util.spec_from_file_location(name, REF_ROOT + fname)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _loaded[fname] = mod
    return mod


def reference_module(cfg, state_dict, drop_path_rate=0.0):
    """Build the reference class for ``cfg.block`` ('aim' -> vitclip_aim.AIM(wind_attn=False),
    'fork' -> vit_clip.ViT_CLIP(shift=False)) and load ``state_dict`` into it."""
    kw = dict(input_resolution=cfg.input_resolution, num_frames=cfg.num_frames, patch_size=cfg.patch_size,
              width=cfg.width, layers=cfg.layers, heads=cfg.heads, drop_path_rate=drop_path_rate,
              adapter_scale=cfg.adapter_scale)
    if cfg.block == "aim":
        m = load("vitclip_aim.py").AIM(num_tadapter=cfg.num_tadapter, **kw)
    else:
        m = load("vit_clip.py").ViT_CLIP(**kw)
    m.init_weights()
    missing = m.load_state_dict(state_dict, strict=True)
    return m


def reference_imagenet(cfg, state_dict, drop_path_rate=0.0):
    """The reference ``vit_imagenet.py::ViT_ImageNet`` (build container only: the file is read from the mounted tree)."""
    m = load("vit_imagenet.py").ViT_ImageNet(img_size=cfg.input_resolution, num_frames=cfg.num_frames, patch_size=cfg.patch_size,
                                             embed_dim=cfg.width, depth=cfg.layers, num_heads=cfg.heads,
                                             adapter_scale=cfg.adapter_scale, num_tadapter=cfg.num_tadapter,
                                             drop_path_rate=drop_path_rate)
    m.init_weights()
    m.load_state_dict(state_dict, strict=True)
    return m


def imagenet_available() -> bool:
    return os.path.isfile(REF_ROOT + "vit_imagenet.py")

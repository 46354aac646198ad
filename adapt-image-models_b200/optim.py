"""AdamW over the backbone's flat trainable buffer: one kernel launch per step.

The reference trains with ``torch.optim.AdamW`` built by mmcv from the config (``optimizer = dict(type='AdamW', lr=...,
betas=(0.9, 0.999), weight_decay=0.05)``, ``configs/recognition/vit/vitclip_base_k400.py``), i.e. a multi-tensor launch
chain over 147 small tensors (~0.14 ms per step on B200).  The backbone keeps those tensors as views of ONE flat fp32
buffer and its backward fills ONE flat gradient buffer (``backbone.py::_flatten_trainable``), so the same update is a single
pass over four arrays (``aimb_adamw_flat``).  Same math as ``torch.optim.AdamW`` (decoupled weight decay, bias correction);
``tests/test_recognizer.py::test_flat_adamw_matches_torch_adamw`` checks it step by step.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from . import lib


def _default_decay(name: str) -> bool:
    """Weight matrices of the adapters decay; biases, LayerNorm parameters and temporal_embedding do not."""
    return "Adapter" in name and name.endswith("weight")


class FlatAdamW:
    """``step()`` / ``zero_grad()`` of an optimizer for the trainable tensors of one aimb200 backbone (plus, optionally, a
    stock ``torch.optim`` optimizer for parameters outside it, e.g. the classification head).  The step counter lives on
    the device, so ``step()`` can be captured into a CUDA graph."""

    def __init__(self, backbone, lr: float = 3e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.05,
                 decay_filter: Optional[Callable[[str], bool]] = None, extra: Optional[torch.optim.Optimizer] = None):
        self.backbone, self.extra = backbone, extra
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.decay_filter = decay_filter or _default_decay
        self._state = None

    def _init_state(self):
        bb = self.backbone
        bb._flatten_trainable(bb._engine_params())
        flat = bb._flat
        mask = torch.zeros(flat.numel(), dtype=torch.uint8, device=flat.device)
        for name in bb.trainable_names():
            o, k = bb._offsets[name]
            if self.decay_filter(name):
                mask[o:o + k] = 1
        self._state = dict(flat=flat, m=torch.zeros_like(flat), v=torch.zeros_like(flat), mask=mask,
                           step=torch.zeros(1, dtype=torch.float32, device=flat.device))

    def zero_grad(self, set_to_none: bool = True):
        for p in self.backbone.parameters():
            if p.requires_grad:
                p.grad = None
        if self.extra is not None:
            self.extra.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self):
        bb = self.backbone
        if self._state is None or self._state["flat"] is not bb._flat:
            self._init_state()
        st = self._state
        g = bb._last_flat_grad
        if g is None:
            raise lib.AimbError("FlatAdamW.step() without a backward of the backbone")
        st["step"].add_(1.0)
        lib.adamw_flat(st["flat"], g, st["m"], st["v"], st["mask"], st["step"], self.lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay)
        if self.extra is not None:
            self.extra.step()

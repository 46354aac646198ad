"""``ViT_CLIP`` — drop-in replacement of ``mmaction/models/backbones/vit_clip.py::ViT_CLIP``.

Same registry name, constructor arguments (superset of vit_clip.py:330-331 and vitclip_aim.py:343,
so every ``configs/recognition/vit/vitclip_*.py`` constructs), ``init_weights()``, ``forward(x[B,3,T,H,W])
-> [B, width, T, 1, 1]`` and ``state_dict`` key/shape layout (CLIP ``visual.*`` dicts and AIM
checkpoints load unchanged).  What differs is *how* forward/backward run: as a fixed sequence of
hand-written sm_100a kernels (``engine.py``) with an adapter-only backward; autograd sees one node.

No CPU / eager fallback: calling ``forward`` with a CPU tensor or without the built library raises.
"""
from __future__ import annotations

import contextlib
import logging
import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import lib
from .engine import Dims, Engine
from .registry import BACKBONES

_log = logging.getLogger("aimb200")
_warned_default_block = False


class LayerNorm(nn.LayerNorm):
    """Parameter container; statistics are always fp32 in the kernels (vit_clip.py:71-77)."""


class QuickGELU(nn.Module):
    def forward(self, x):  # only used if someone calls the sub-module directly; the engine fuses it
        return x * torch.sigmoid(1.702 * x)


class Adapter(nn.Module):
    """vit_clip.py:51-69 — D_fc1 (D->r), exact GELU, D_fc2 (r->D), optional skip."""

    def __init__(self, D_features, mlp_ratio=0.25, skip_connect=True):
        super().__init__()
        self.skip_connect = skip_connect
        hidden = int(D_features * mlp_ratio)
        self.act = nn.GELU()
        self.D_fc1 = nn.Linear(D_features, hidden)
        self.D_fc2 = nn.Linear(hidden, D_features)


class ResidualAttentionBlock(nn.Module):
    """Parameter tree of vit_clip.py:85-113 / vitclip_aim.py:111-135 (the math lives in engine.py)."""

    def __init__(self, d_model, n_head, scale=1.0, num_tadapter=1, num_frames=8, drop_path=0.0, block="aim"):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)          # parameter container only, as in the reference
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.n_head, self.d_model, self.scale, self.num_frames = n_head, d_model, scale, num_frames
        self.num_tadapter = num_tadapter
        self.MLP_Adapter = Adapter(d_model, skip_connect=False)
        self.S_Adapter = Adapter(d_model, skip_connect=(block == "aim"))   # aim: skip (vitclip_aim.py:130); fork: none
        self.T_Adapter = Adapter(d_model, skip_connect=False)
        if num_tadapter == 2:
            self.T_Adapter_in = Adapter(d_model)
        self.drop_path_rate = float(drop_path)


class Transformer(nn.Module):
    def __init__(self, num_frames, width, layers, heads, num_tadapter=1, scale=1.0, drop_path=0.1, block="aim"):
        super().__init__()
        self.width, self.layers = width, layers
        dpr = [x.item() for x in torch.linspace(0, drop_path, layers)]      # vit_clip.py:297
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, scale, num_tadapter, num_frames, dpr[i],
                                                                block) for i in range(layers)])


def _is_trainable_name(name: str) -> bool:
    """Freeze rule of vit_clip.py:413-415 (``cls_head`` lives outside the backbone)."""
    return ("temporal_embedding" in name) or ("ln_post" in name) or ("Adapter" in name)


class _BackboneFn(torch.autograd.Function):
    """One autograd node for the whole backbone; backward = Engine.backward (adapter-only grads)."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        ctx.mod = mod
        out = mod._run_forward(x, training=True)
        ctx.gen = mod._gen                  # which training forward's activations this node owns
        return out

    @staticmethod
    def backward(ctx, dfeat):
        mod = ctx.mod
        if ctx.gen != mod._gen or mod._step_ctx is None:
            # one set of saved activations per module (stable pointers, CUDA-graph friendly): a later grad-enabled
            # forward has overwritten the ones this backward needs
            raise lib.AimbError(
                "aimb200.ViT_CLIP keeps ONE training forward in flight: backward() of forward #%d was called after forward "
                "#%d overwrote its saved activations.  Run backward before the next grad-enabled forward (batch the views "
                "/ clips into one call, as Recognizer3D does), or use torch.no_grad() for forwards that need no gradient."
                % (ctx.gen, mod._gen))
        grads = mod._run_backward(dfeat)
        return (None, None) + tuple(grads)


@BACKBONES.register_module()
class ViT_CLIP(nn.Module):
    def __init__(self, input_resolution: int, num_frames: int, patch_size: int, width: int, layers: int, heads: int,
                 drop_path_rate, num_tadapter=1, adapter_scale=0.5, pretrained=None, shift=False, checkpoint=False,
                 block: Optional[str] = None, compute_dtype: Optional[str] = None):
        super().__init__()
        global _warned_default_block
        if block is None and "AIMB200_BLOCK" not in os.environ and type(self).__name__ == "ViT_CLIP" and not _warned_default_block:
            _warned_default_block = True
            # The literal in-tree class of this name (vit_clip.py:199-288) is the fork block; the upstream AIM math that
            # north_star / README GFLOPs / every vitclip_*.py config (num_tadapter) describe lives in the reference as class
            # `AIM` (vitclip_aim.py).  Both are built here; say which one a bare `type='ViT_CLIP'` gets.
            _log.warning("aimb200.ViT_CLIP: block not given -> 'aim' (upstream AIM block = reference class AIM, "
                         "vitclip_aim.py:196-211).  Checkpoints trained with the in-tree vit_clip.py::ViT_CLIP need "
                         "block='fork' (or AIMB200_BLOCK=fork); type='AIM' always selects the upstream block.")
        block = os.environ.get("AIMB200_BLOCK", block or "aim")
        if block not in ("aim", "fork"):
            raise ValueError("block must be 'aim' (upstream AIM math, vitclip_aim.py:196-211) or 'fork' (vit_clip.py:199-288)")
        if block == "fork" and num_tadapter != 1:
            raise ValueError("num_tadapter==2 is only defined for block='aim' (the in-tree fork has no T_Adapter_in)")
        if shift:
            # The reference's own shift branch cannot run: it reshapes xln[2:] (n - 2 = G*G - 1 tokens) to an h x w grid
            # with h = w = int(sqrt(n - 2)) and einops raises for every resolution (checked against the reference at
            # 224 and 64, tests/test_boundary.py::test_shift_true_raises_in_reference_too).  Every in-tree config has
            # shift=False, so there is no behaviour to reproduce.
            raise NotImplementedError("shift=True (PatchShift, vit_clip.py:15-49,233-254) is outside the built path")
        if width % heads or width // heads != 64:
            raise ValueError("head_dim must be 64 (CLIP ViT-B/16, ViT-L/14)")
        if num_tadapter not in (1, 2):
            raise ValueError("num_tadapter must be 1 or 2")
        compute_dtype = os.environ.get("AIMB200_DTYPE", compute_dtype or "bf16")
        if compute_dtype not in ("bf16", "fp32"):
            raise ValueError("compute_dtype must be 'bf16' or 'fp32'")
        self.block = block
        self.compute_dtype = torch.bfloat16 if compute_dtype == "bf16" else torch.float32
        self.input_resolution, self.patch_size, self.num_frames = input_resolution, patch_size, num_frames
        self.width, self.layers, self.heads = width, layers, heads
        self.num_tadapter, self.adapter_scale = num_tadapter, float(adapter_scale)
        self.drop_path_rate = float(drop_path_rate)
        self.pretrained = pretrained
        self.checkpoint = bool(checkpoint)   # per-block activation recompute in backward (vit_clip.py:318-319)
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.temporal_embedding = nn.Parameter(torch.zeros(1, num_frames, width))
        self.transformer = Transformer(num_frames, width, layers, heads, num_tadapter=num_tadapter, scale=adapter_scale,
                                       drop_path=drop_path_rate, block=block)
        self.ln_post = LayerNorm(width)
        # runtime state (not parameters)
        self._engine: Optional[Engine] = None
        self._frozen_cache: Dict[str, tuple] = {}
        self._train_names: Optional[List[str]] = None
        self._flat: Optional[torch.Tensor] = None
        self._flat_ptrs = None
        self._grad_sync = None
        self._input_norm = None
        self._step_ctx = None
        self._gen = 0
        self._norm_dev = None
        self._last_flat_grad = None

    # ------------------------------------------------------------------ reference API
    def init_weights(self, pretrained=None):
        """vit_clip.py:352-423: trunc-normal(.02) linears, LN 1/0, optional CLIP load, zero D_fc2, freeze."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str):
            self.apply(_init)
            self.load_state_dict(self._load_pretrained(self.pretrained), strict=False)
        elif self.pretrained is None:
            self.apply(_init)
        else:
            raise TypeError('pretrained must be a str or None')
        for n, m in self.transformer.named_modules():
            if 'Adapter' in n and n.endswith('D_fc2') and isinstance(m, nn.Linear):
                nn.init.constant_(m.weight, 0)
                nn.init.constant_(m.bias, 0)
        for name, param in self.named_parameters():
            param.requires_grad = _is_trainable_name(name)
        self.invalidate_cache()

    def _load_pretrained(self, spec: str):
        if os.path.isfile(spec):
            sd = torch.load(spec, map_location="cpu")
            sd = sd.get("state_dict", sd)
            sd = {k[len("backbone."):] if k.startswith("backbone.") else k: v for k, v in sd.items()}
            sd.pop("proj", None)
            return sd
        try:  # the reference path: OpenAI CLIP (vit_clip.py:369-376)
            import clip  # type: ignore
        except ImportError as e:
            raise RuntimeError(f"pretrained={spec!r} needs the OpenAI `clip` package and weights (not available "
                               "offline); pass a path to a state_dict file instead") from e
        model, _ = clip.load("ViT-B/16" if self.layers == 12 else "ViT-L/14", device="cpu")
        sd = model.visual.state_dict()
        sd.pop("proj", None)
        return sd

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed', 'temporal_embedding'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table', 'temporal_position_bias_table'}

    # ------------------------------------------------------------------ extras either side of the path
    def set_input_normalization(self, mean, std):
        """Fuse GPUNormalize (mmaction/utils/module_hooks.py:35-87) into the patch load: uint8 clips in."""
        self._input_norm = (torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32))
        self._norm_dev = None            # device copies are made once per device (no per-forward H2D: graph capturable)

    def attach_grad_sync(self, sync):
        """sync: object with ``bucket_done(flat_grad, lo, hi)`` and ``finish()`` (see parallel.GradSync)."""
        self._grad_sync = sync

    def invalidate_cache(self):
        self._frozen_cache.clear()
        self._flat = None
        self._flat_ptrs = None

    def train(self, mode: bool = True):
        return super().train(mode)

    def _apply(self, fn, *a, **k):      # .to()/.cuda()/.half() replace parameter storage
        self.invalidate_cache()
        if self._engine is not None:
            self._engine.release()
            self._engine = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ weights in compute dtype
    def _ekey(self, name: str) -> str:
        """state_dict name -> the key the engine uses (identity here; ViT_ImageNet maps its timm-style names)."""
        return name

    def _engine_params(self) -> Dict[str, nn.Parameter]:
        return {self._ekey(n): p for n, p in self.named_parameters()}

    def _dims(self, B: int) -> Dims:
        p, res = self.patch_size, self.input_resolution
        K = 3 * p * p
        return Dims(B=B, T=self.num_frames, n=(res // p) ** 2 + 1, D=self.width, heads=self.heads, L=self.layers,
                    r=int(self.width * 0.25), patch=p, res=res, kpad=(K + 63) // 64 * 64, num_tadapter=self.num_tadapter,
                    scale=self.adapter_scale, block=self.block)

    def trainable_names(self) -> List[str]:
        """Flat-buffer order: temporal_embedding, block 0 .. L-1 adapters, ln_post — so that the gradient
        slices completed by backward (ln_post, block L-1, ..., block 0, temporal_embedding) are contiguous."""
        if self._train_names is None:
            names = ["temporal_embedding"]
            ads = ["T_Adapter", "S_Adapter", "MLP_Adapter"] + (["T_Adapter_in"] if self.num_tadapter == 2 else [])
            for i in range(self.layers):
                for a in ads:
                    for leaf in ("D_fc1.weight", "D_fc1.bias", "D_fc2.weight", "D_fc2.bias"):
                        names.append(f"transformer.resblocks.{i}.{a}.{leaf}")
            names += ["ln_post.weight", "ln_post.bias"]
            self._train_names = names
        return self._train_names

    def _flatten_trainable(self, params: Dict[str, nn.Parameter]):
        """Make the trainable fp32 masters views of one flat buffer (one cast per step, bucketable grads)."""
        names = self.trainable_names()
        ptrs = tuple(params[n].data_ptr() for n in names)
        if self._flat is not None and ptrs == self._flat_ptrs:
            return
        total = sum(params[n].numel() for n in names)
        dev = params[names[0]].device
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        offs, o = {}, 0
        for n in names:
            p = params[n]
            if p.dtype != torch.float32:
                raise lib.AimbError("trainable parameters must be fp32 masters (compute dtype is chosen by compute_dtype)")
            k = p.numel()
            flat[o:o + k].copy_(p.data.reshape(-1).float())
            p.data = flat[o:o + k].view(p.shape)
            offs[n] = (o, k)
            o += k
        self._flat, self._offsets = flat, offs
        self._flat_ptrs = tuple(params[n].data_ptr() for n in names)
        # gradient bucket boundaries (in elements): [temb | block 0 .. L-1 | ln_post]
        self._block_lo = [offs[f"transformer.resblocks.{i}.T_Adapter.D_fc1.weight"][0] for i in range(self.layers)]

    def _weights(self, training: bool):
        """Return (W, WT): every weight in the compute dtype; WT = transposes for the dgrad GEMMs."""
        cd = self.compute_dtype
        params = self._engine_params()
        self._flatten_trainable(params)
        W: Dict[str, torch.Tensor] = {}
        WT: Dict[str, torch.Tensor] = {}
        tset = set(self.trainable_names())
        # ---- frozen: cached, refreshed when the source parameter changes (load_state_dict, .to())
        for name, p in params.items():
            if name in tset:
                continue
            ver = (p.data_ptr(), p._version, p.dtype, cd)
            ent = self._frozen_cache.get(name)
            if ent is None or ent[0] != ver:
                src = p.detach()
                if name == "conv1.weight":
                    K = src[0].numel()
                    kpad = (K + 63) // 64 * 64
                    w2 = torch.zeros(src.shape[0], kpad, dtype=cd, device=src.device)
                    w2[:, :K] = src.reshape(src.shape[0], K).to(cd)
                    ent = (ver, w2, None)
                elif name in ("temporal_embedding", "positional_embedding"):       # [1, T, D] / [1, n, D] (timm) -> rows
                    ent = (ver, src.to(cd).reshape(-1, src.shape[-1]).contiguous(), None)
                elif name == "class_embedding":
                    ent = (ver, src.to(cd).reshape(-1).contiguous(), None)
                else:
                    w = src.to(cd).contiguous()
                    wt = None
                    if w.dim() == 2 and (name.endswith("in_proj_weight") or name.endswith("out_proj.weight")
                                         or name.endswith("c_fc.weight") or name.endswith("c_proj.weight")):
                        wt = w.t().contiguous()
                    ent = (ver, w, wt)
                self._frozen_cache[name] = ent
            W[name] = ent[1]
            if ent[2] is not None:
                WT[name] = ent[2]
        # ---- trainable: one cast of the flat master per step
        flat_c = self._flat if cd == torch.float32 else self._flat.to(cd)
        self._step_tmp = [flat_c] if flat_c is not self._flat else []
        for name in self.trainable_names():
            o, k = self._offsets[name]
            shape = params[name].shape
            t = flat_c[o:o + k].view(shape)
            if name == "temporal_embedding":
                t = t.reshape(-1, shape[-1])
            W[name] = t
        if training:
            # dgrad through the (trainable) adapter linears needs W^T: all of them in ONE batched launch
            wnames = [n for n in self.trainable_names() if n.endswith("D_fc1.weight") or n.endswith("D_fc2.weight")]
            tb = getattr(self, "_tr_table", None)
            if tb is None or tb[0] is not self._flat:
                rows = [[self._offsets[n][0], params[n].shape[0], params[n].shape[1]] for n in wnames]
                self._tr_table = (self._flat, torch.tensor(rows, dtype=torch.int64, device=self._flat.device))
            flat_t = torch.empty_like(flat_c)
            self._step_tmp.append(flat_t)
            lib.transpose_batched(flat_c, flat_t, self._tr_table[1], len(wnames))
            for n in wnames:
                o, k = self._offsets[n]
                r, c = params[n].shape
                WT[n] = flat_t[o:o + k].view(c, r)
        if self._input_norm is not None:
            dev = self._flat.device
            if self._norm_dev is None or self._norm_dev[0] != dev:
                self._norm_dev = (dev, self._input_norm[0].to(dev), self._input_norm[1].to(dev))
            W["input_mean"], W["input_std"] = self._norm_dev[1], self._norm_dev[2]
        return W, WT

    # ------------------------------------------------------------------ forward / backward
    def _drop_masks(self, d: Dims, device):
        if not self.training or self.drop_path_rate <= 0.0:
            return None
        rates = torch.linspace(0, self.drop_path_rate, self.layers)
        ck = getattr(self, "_keep_cache", None)
        if ck is None or ck[0] != (str(device), self.drop_path_rate):
            self._keep_cache = ((str(device), self.drop_path_rate), (1.0 - rates).to(device).view(-1, 1, 1))
        keep = self._keep_cache[1]
        # timm DropPath on LND tensors: one Bernoulli draw per token index, shared by all frames;
        # two independent draws per block (temporal branch, MLP-adapter branch)
        m = (torch.rand(self.layers, 2, d.n, device=device) < keep).float() / keep
        out = []
        for i in range(self.layers):
            if float(rates[i]) == 0.0:
                out.append((None, None))
            else:
                out.append((m[i, 0].contiguous(), m[i, 1].contiguous()))
        return out

    def _run_forward(self, x: torch.Tensor, training: bool) -> torch.Tensor:
        if self._engine is None or self._engine.dtype != self.compute_dtype or self._engine.device != x.device:
            self._engine = Engine(self.compute_dtype, x.device)
        d = self._dims(x.shape[0])
        if x.dtype == torch.float16:      # apex O1 / auto_fp16 callers (recognizers/base.py:141): exact in fp32 mode, and the
            x = x.to(self.compute_dtype)  # patch GEMM reads bf16 anyway in bf16 mode
        with torch.cuda.device(x.device):     # kernels and streams follow the tensors, not the caller's current device
            # The per-step weight preparation (bf16 cast of the flat fp32 master, batched adapter-weight transposes, DropPath
            # masks, and in the engine the batched W1 Wo products) does not depend on the clip: it runs on a side stream, next to
            # the patch embedding; the engine waits for it where the first prepared tensor is read (~0.1 ms off the step)
            main = torch.cuda.current_stream(x.device)
            overlap = training and os.environ.get("AIMB200_PREP_STREAM", "1") == "1"
            prep = self._engine.prep_stream() if overlap else main
            if overlap:
                prep.wait_stream(main)
            with torch.cuda.stream(prep):
                W, WT = self._weights(training)
                masks = self._drop_masks(d, x.device) if training else None
                ready = None
                if overlap:
                    for t in self._step_tmp:            # the per-step allocations (bf16 copy of the flat master, its transposes):
                        t.record_stream(main)           # produced on the side stream, consumed (and later freed) on the main one
                    for pair in (masks or []):
                        for t in pair:
                            if t is not None:
                                t.record_stream(main)
                    ready = torch.cuda.Event()
                    ready.record(prep)
            if training:
                self._step_ctx = (W, WT, d)
                self._gen += 1
            return self._engine.forward(x.contiguous(), W, d, training, masks, WT, checkpoint=self.checkpoint,
                                        weights_ready=(ready, prep) if overlap else None)

    def _run_backward(self, dfeat: torch.Tensor):
        W, WT, d = self._step_ctx
        names = self.trainable_names()
        flat_grad = torch.zeros_like(self._flat)      # ONE memset: the gradient kernels accumulate into it
        params = self._engine_params()
        grads = {}
        for n in names:
            o, k = self._offsets[n]
            g = flat_grad[o:o + k].view(params[n].shape)
            grads[n] = g
        sync = self._grad_sync
        L = self.layers
        total = flat_grad.numel()
        done_hi = [total]

        def on_done(i):
            if sync is None:
                return
            # completed so far (from the top of the flat buffer): ln_post (i == L), blocks i..L-1, temb (i == -1)
            lo = 0 if i == -1 else (self._block_lo[i] if i < L else self._offsets["ln_post.weight"][0])
            if sync.want_bucket(i, L):
                sync.bucket_done(flat_grad, lo, done_hi[0])
                done_hi[0] = lo

        with torch.cuda.device(flat_grad.device) if flat_grad.is_cuda else contextlib.nullcontext():
            self._engine.backward(dfeat.reshape(d.B, d.D, d.T).float(), W, WT, grads, on_done, grads_prezeroed=True,
                                  bucket_ends_at=(lambda i: sync.want_bucket(i, L)) if sync is not None else None)
        if sync is not None:
            if done_hi[0] > 0:
                sync.bucket_done(flat_grad, 0, done_hi[0])
            sync.finish()
        self._step_ctx = None
        self._last_flat_grad = flat_grad            # optim.FlatAdamW consumes the flat buffer directly
        out = []
        for p_name, p in self.named_parameters():
            out.append(grads.get(self._ekey(p_name)) if p.requires_grad else None)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, 3, T, H, W] -> [B, width, T, 1, 1] (fp32), as vit_clip.py:433-458."""
        if not x.is_cuda:
            raise lib.AimbError("aimb200.ViT_CLIP runs on sm_100a only: move the module and the clip to a CUDA device "
                                "(there is no CPU fallback)")
        B, C, T, H, Wd = x.shape
        if C != 3 or T != self.num_frames or H != self.input_resolution or Wd != self.input_resolution:
            raise ValueError(f"expected [B,3,{self.num_frames},{self.input_resolution},{self.input_resolution}], got {tuple(x.shape)}")
        if x.dtype == torch.uint8 and self._input_norm is None:
            raise ValueError("uint8 clips need set_input_normalization(mean, std)")
        if B == 0:
            return torch.zeros(0, self.width, T, 1, 1, device=x.device, dtype=torch.float32)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            params = list(self.parameters())
            for name, p in self.named_parameters():
                if p.requires_grad and not _is_trainable_name(name):
                    raise lib.AimbError(
                        f"parameter {name} requires grad, but the hand-written backward produces gradients only for "
                        "the AIM trainable set (temporal_embedding, ln_post, *Adapter*); call init_weights() or freeze it")
            feat = _BackboneFn.apply(self, x, *params)
        else:
            feat = self._run_forward(x, training=False)
        return feat.view(B, self.width, T, 1, 1)


@BACKBONES.register_module()
class AIM(ViT_CLIP):
    """Registry name of the reference's upstream-math class (vitclip_aim.py:340-344, default path wind_attn=False):
    same parameter tree, always block='aim'.  The window-attention / prompt options of that class are research variants
    outside the built path and are rejected rather than silently ignored."""

    def __init__(self, input_resolution: int, num_frames: int, patch_size: int, width: int, layers: int, heads: int,
                 drop_path_rate, num_tadapter=1, adapter_scale=0.5, pretrained=None, prompt=True, wind_attn=False,
                 window_size=(32, 2, 2), not_shift=True, compute_dtype: Optional[str] = None):
        if wind_attn:
            raise NotImplementedError("AIM(wind_attn=True) (3-D shifted-window path, vitclip_aim.py:212-287) is outside the "
                                      "built path; wind_attn=False is the upstream AIM block")
        super().__init__(input_resolution, num_frames, patch_size, width, layers, heads, drop_path_rate,
                         num_tadapter=num_tadapter, adapter_scale=adapter_scale, pretrained=pretrained, block="aim",
                         compute_dtype=compute_dtype)


class _TimmAttention(nn.Module):
    """Parameter container of vit_imagenet.py:54-86 (qkv / proj linears)."""

    def __init__(self, dim, qkv_bias=True):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _TimmMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class _TimmBlock(nn.Module):
    """Parameter tree of vit_imagenet.py:88-108 (the math lives in engine.py: same AIM block, exact-GELU MLP)."""

    def __init__(self, dim, num_tadapter=1, qkv_bias=True, mlp_ratio=4.0, eps=1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = _TimmAttention(dim, qkv_bias)
        self.MLP_Adapter = Adapter(dim, skip_connect=False)
        self.S_Adapter = Adapter(dim)
        self.T_Adapter = Adapter(dim, skip_connect=False)
        if num_tadapter == 2:
            self.T_Adapter_in = Adapter(dim)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _TimmMlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)


_TIMM_TO_ENGINE = (("patch_embed.proj.", "conv1."), ("cls_token", "class_embedding"), ("pos_embed", "positional_embedding"),
                   (".norm1.", ".ln_1."), (".norm2.", ".ln_2."), (".attn.qkv.weight", ".attn.in_proj_weight"),
                   (".attn.qkv.bias", ".attn.in_proj_bias"), (".attn.proj.", ".attn.out_proj."), (".mlp.fc1.", ".mlp.c_fc."),
                   (".mlp.fc2.", ".mlp.c_proj."))


@BACKBONES.register_module()
class ViT_ImageNet(ViT_CLIP):
    """Drop-in for ``mmaction/models/backbones/vit_imagenet.py::ViT_ImageNet``: the AIM block on a timm
    ViT-B/16 — same constructor arguments, ``state_dict`` keys / shapes (``patch_embed.proj``, ``cls_token``, ``pos_embed``,
    ``blocks.{i}.norm1 / attn.qkv / attn.proj / mlp.fc1 / mlp.fc2 / *_Adapter``, ``ln_post``), ``init_weights()`` and
    ``forward([B,3,T,H,W]) -> [B, D, T, 1, 1]``.  Differences from the CLIP variant, all handled inside the same engine and
    kernels: patch embedding with bias, no ``ln_pre``, LayerNorm eps 1e-6, exact GELU in the MLP, and DropPath drawn per
    FRAME (the block works on ``[(b t), n, d]`` tensors, so timm's DropPath masks dim 0 = frames, vit_imagenet.py:110-126).

    Training scope: the hand-written backward produces the AIM trainable set (adapters, ``temporal_embedding``, ``ln_post``).
    The reference class itself freezes nothing (full fine-tuning through autograd); that weight-gradient path is not built,
    so ``freeze_backbone=True`` (the default, the AIM recipe) freezes the pre-trained tensors in ``init_weights()`` and a
    grad-enabled forward with any other trainable parameter raises ``AimbError`` instead of returning wrong gradients."""

    def __init__(self, img_size=224, num_frames=8, patch_size=16, in_chans=3, embed_dim=768, depth=12, adapter_scale=0.5,
                 num_tadapter=1, num_heads=12, mlp_ratio=4., patch_embedding_bias=True, qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=None, pretrained=None,
                 freeze_backbone: bool = True, compute_dtype: Optional[str] = None):
        nn.Module.__init__(self)
        if in_chans != 3 or mlp_ratio != 4.0:
            raise NotImplementedError("ViT_ImageNet: in_chans=3 and mlp_ratio=4 are what the kernels implement")
        if drop_rate or attn_drop_rate:
            raise NotImplementedError("ViT_ImageNet: drop_rate / attn_drop_rate > 0 are not used by any in-tree config and not built")
        if embed_dim % num_heads or embed_dim // num_heads != 64:
            raise ValueError("head_dim must be 64")
        if qk_scale is not None and abs(qk_scale - 0.125) > 1e-12:
            raise NotImplementedError("qk_scale other than head_dim ** -0.5 is not built")
        if num_tadapter not in (1, 2):
            raise ValueError("num_tadapter must be 1 or 2")
        eps = 1e-6                                    # the reference default: partial(nn.LayerNorm, eps=1e-6)
        if norm_layer is not None:
            eps = float(getattr(norm_layer(8), "eps", 1e-6))
        compute_dtype = os.environ.get("AIMB200_DTYPE", compute_dtype or "bf16")
        if compute_dtype not in ("bf16", "fp32"):
            raise ValueError("compute_dtype must be 'bf16' or 'fp32'")
        self.block = "aim"
        self.compute_dtype = torch.bfloat16 if compute_dtype == "bf16" else torch.float32
        self.input_resolution, self.patch_size, self.num_frames = img_size, patch_size, num_frames
        self.width, self.layers, self.heads = embed_dim, depth, num_heads
        self.num_features = self.embed_dim = embed_dim
        self.depth = depth
        self.num_tadapter, self.adapter_scale = num_tadapter, float(adapter_scale)
        self.drop_path_rate = float(drop_path_rate)
        self.pretrained = pretrained
        self.checkpoint = False
        self.freeze_backbone = bool(freeze_backbone)
        self.ln_eps = eps
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim, bias=patch_embedding_bias)
        n_patches = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patches + 1, embed_dim))
        self.temporal_embedding = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        self.blocks = nn.ModuleList([_TimmBlock(embed_dim, num_tadapter, qkv_bias, mlp_ratio, eps) for _ in range(depth)])
        self.ln_post = nn.LayerNorm(embed_dim, eps=eps)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self._engine = None
        self._frozen_cache = {}
        self._train_names = None
        self._flat = None
        self._flat_ptrs = None
        self._grad_sync = None
        self._input_norm = None
        self._step_ctx = None
        self._gen = 0
        self._norm_dev = None
        self._last_flat_grad = None
        self._zero_qkv_bias = None

    def _ekey(self, name: str) -> str:
        if name.startswith("blocks."):
            name = "transformer.resblocks." + name[len("blocks."):]
        for a, b in _TIMM_TO_ENGINE:
            name = name.replace(a, b)
        return name

    def _dims(self, B: int) -> Dims:
        d = super()._dims(B)
        return Dims(**{**d.__dict__, "eps": self.ln_eps, "mlp_act": lib.ACT_GELU, "ln_pre": False})

    def _weights(self, training: bool):
        W, WT = super()._weights(training)
        if "transformer.resblocks.0.attn.in_proj_bias" not in W:          # qkv_bias=False: the kernels take a zero bias
            if self._zero_qkv_bias is None or self._zero_qkv_bias.device != self._flat.device:
                self._zero_qkv_bias = torch.zeros(3 * self.width, dtype=self.compute_dtype, device=self._flat.device)
            for i in range(self.layers):
                W[f"transformer.resblocks.{i}.attn.in_proj_bias"] = self._zero_qkv_bias
        return W, WT

    def _drop_masks(self, d: Dims, device):
        """timm DropPath on [(b t), n, d] tensors: one Bernoulli draw per FRAME (vit_imagenet.py:110-126), expanded to one
        multiplier per activation row (row = frame * n + token) for the epilogues' row_scale[m % len]."""
        if not self.training or self.drop_path_rate <= 0.0:
            return None
        rates = torch.linspace(0, self.drop_path_rate, self.layers)
        keep = (1.0 - rates).to(device).view(-1, 1, 1)
        m = ((torch.rand(self.layers, 2, d.BT, device=device) < keep).float() / keep).repeat_interleave(d.n, dim=2)
        return [(None, None) if float(rates[i]) == 0.0 else (m[i, 0].contiguous(), m[i, 1].contiguous())
                for i in range(self.layers)]

    def init_weights(self, pretrained=None):
        """vit_imagenet.py:182-229: trunc-normal(.02) linears, LN 1/0, optional timm checkpoint (``norm.*`` -> ``ln_post.*``),
        zero ``D_fc2`` of every adapter; then the AIM freeze rule unless ``freeze_backbone=False``."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if pretrained:
            self.pretrained = pretrained
        if isinstance(self.pretrained, str):
            self.apply(_init)
            if not os.path.isfile(self.pretrained):
                raise RuntimeError(f"pretrained={self.pretrained!r}: pass the path of a timm ViT state_dict file "
                                   "(the reference reads checkpoints/jx_vit_base_p16_224-80ecf9dd.pth, vit_imagenet.py:199)")
            sd = torch.load(self.pretrained, map_location="cpu")
            sd = sd.get("state_dict", sd)
            if "norm.weight" in sd:
                sd["ln_post.weight"], sd["ln_post.bias"] = sd["norm.weight"], sd["norm.bias"]
            self.load_state_dict(sd, strict=False)
        elif self.pretrained is None:
            self.apply(_init)
        else:
            raise TypeError('pretrained must be a str or None')
        for n, m in self.blocks.named_modules():
            if 'Adapter' in n and n.endswith('D_fc2') and isinstance(m, nn.Linear):
                nn.init.constant_(m.weight, 0)
                nn.init.constant_(m.bias, 0)
        if self.freeze_backbone:
            for name, param in self.named_parameters():
                param.requires_grad = _is_trainable_name(name)
        self.invalidate_cache()

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'temporal_embedding'}

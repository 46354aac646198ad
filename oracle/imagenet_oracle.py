"""CPU oracle for the ``ViT_ImageNet`` variant of the AIM block.  TEST INFRASTRUCTURE ONLY (same rules as
``oracle/aim_oracle.py``: only ``tests/`` and the golden generator import it; the product never does).

From-the-math restatement (plain torch on CPU, fp32 / fp64) of ``mmaction/models/backbones/vit_imagenet.py``:
  * patch embedding with bias, cls / pos / temporal embeddings, NO ln_pre     vit_imagenet.py:232-251
  * block: temporal / spatial / joint adaptation on [(b t), n, d] tensors      vit_imagenet.py:110-126
  * attention (fused qkv linear, q/k/v = row thirds, scale head_dim**-0.5)     vit_imagenet.py:54-86
  * Mlp with exact (erf) GELU                                                   vit_imagenet.py:36-52
  * LayerNorm eps = 1e-6                                                        vit_imagenet.py:151
  * tail: ln_post, cls row, '(b t) c -> b c t'                                  vit_imagenet.py:256-260
  * DropPath (timm): mask shape (x.shape[0], 1, 1) on [(b t), n, d] tensors => one Bernoulli draw per FRAME.

Pinned against the real reference class executed in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/tiny_imagenet.npz``; ``tests/test_oracle_golden.py``).  Parameter names are the reference's own
(``patch_embed.proj``, ``cls_token``, ``pos_embed``, ``blocks.{i}.norm1 / attn.qkv / attn.proj / mlp.fc1 / mlp.fc2``).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import aim_oracle as O

EPS = 1e-6


def param_shapes(cfg: O.OracleCfg) -> Dict[str, tuple]:
    D, n, T, p, r = cfg.width, cfg.tokens, cfg.num_frames, cfg.patch_size, cfg.hidden
    s: Dict[str, tuple] = {"cls_token": (1, 1, D), "pos_embed": (1, n, D), "temporal_embedding": (1, T, D),
                           "patch_embed.proj.weight": (D, 3, p, p), "patch_embed.proj.bias": (D,)}
    adapters = ["MLP_Adapter", "S_Adapter", "T_Adapter"] + (["T_Adapter_in"] if cfg.num_tadapter == 2 else [])
    for i in range(cfg.layers):
        pre = f"blocks.{i}."
        s[pre + "norm1.weight"], s[pre + "norm1.bias"] = (D,), (D,)
        s[pre + "attn.qkv.weight"], s[pre + "attn.qkv.bias"] = (3 * D, D), (3 * D,)
        s[pre + "attn.proj.weight"], s[pre + "attn.proj.bias"] = (D, D), (D,)
        for a in adapters:
            s[pre + a + ".D_fc1.weight"], s[pre + a + ".D_fc1.bias"] = (r, D), (r,)
            s[pre + a + ".D_fc2.weight"], s[pre + a + ".D_fc2.bias"] = (D, r), (D,)
        s[pre + "norm2.weight"], s[pre + "norm2.bias"] = (D,), (D,)
        s[pre + "mlp.fc1.weight"], s[pre + "mlp.fc1.bias"] = (4 * D, D), (4 * D,)
        s[pre + "mlp.fc2.weight"], s[pre + "mlp.fc2.bias"] = (D, 4 * D), (D,)
    s["ln_post.weight"], s["ln_post.bias"] = (D,), (D,)
    return s


def fixture_state_dict(cfg: O.OracleCfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic weights, one CPU generator in ``param_shapes`` order; every bias / D_fc2 / embedding non-zero."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name in ("cls_token", "pos_embed", "temporal_embedding"):
            t = torch.randn(shape, generator=g) * 0.02
        elif name == "patch_embed.proj.weight":
            t = torch.randn(shape, generator=g) * (shape[1] * shape[2] * shape[3]) ** -0.5
        elif "norm" in name.split(".")[-2] or name.startswith("ln_post"):
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.05 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def _attention(p, pre, x, heads):
    D = x.shape[-1]
    qkv = x @ p[pre + "attn.qkv.weight"].T + p[pre + "attn.qkv.bias"]
    o, _ = O.mha_core(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], heads)
    return o @ p[pre + "attn.proj.weight"].T + p[pre + "attn.proj.bias"]


def block(p, i, x, cfg: O.OracleCfg, B: int, mask_t=None, mask_m=None):
    """vit_imagenet.py:110-126.  x [BT, n, D]; mask_* : per-FRAME DropPath multipliers [BT] (already / keep) or None."""
    pre = f"blocks.{i}."
    T, n = cfg.num_frames, x.shape[1]
    n1w, n1b = p[pre + "norm1.weight"], p[pre + "norm1.bias"]
    xt = O.layer_norm(O._rows_to_temporal(x, B, T), n1w, n1b, EPS)
    if cfg.num_tadapter == 2:
        xt = O.adapter(p, pre + "T_Adapter_in", xt, skip=True)
    xt = O._temporal_to_rows(O.adapter(p, pre + "T_Adapter", _attention(p, pre, xt, cfg.heads), skip=False), B, n)
    if mask_t is not None:
        xt = xt * mask_t.view(-1, 1, 1)
    x = x + xt
    x = x + O.adapter(p, pre + "S_Adapter", _attention(p, pre, O.layer_norm(x, n1w, n1b, EPS), cfg.heads), skip=True)
    xn = O.layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"], EPS)
    mlp = O.gelu_erf(xn @ p[pre + "mlp.fc1.weight"].T + p[pre + "mlp.fc1.bias"]) @ p[pre + "mlp.fc2.weight"].T + p[pre + "mlp.fc2.bias"]
    ad = cfg.adapter_scale * O.adapter(p, pre + "MLP_Adapter", xn, skip=False)
    if mask_m is not None:
        ad = ad * mask_m.view(-1, 1, 1)
    return x + mlp + ad


def backbone(p, x, cfg: O.OracleCfg, drop_masks: Optional[List] = None):
    """x [B,3,T,H,W] -> [B, D, T, 1, 1]  (vit_imagenet.py:232-260)."""
    B, C, T, H, W = x.shape
    ps, G, D = cfg.patch_size, cfg.grid, cfg.width
    fr = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    cols = fr.reshape(B * T, C, G, ps, G, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * T, G * G, C * ps * ps)
    tok = cols @ p["patch_embed.proj.weight"].reshape(D, -1).T + p["patch_embed.proj.bias"]
    z = torch.cat([p["cls_token"].expand(B * T, 1, D), tok], 1) + p["pos_embed"]
    z = (z.reshape(B, T, -1, D) + p["temporal_embedding"].view(1, T, 1, D)).reshape(B * T, -1, D)
    for i in range(cfg.layers):
        ma, mm = (None, None) if drop_masks is None else drop_masks[i]
        z = block(p, i, z, cfg, B, ma, mm)
    cls = O.layer_norm(z[:, 0, :], p["ln_post.weight"], p["ln_post.bias"], EPS)
    return cls.reshape(B, T, -1).permute(0, 2, 1).unsqueeze(-1).unsqueeze(-1)


def loss_and_grads(p, x, labels, cfg: O.OracleCfg, head_w, head_b, drop_masks=None):
    """CE loss of the I3D-head logits and gradients of the AIM trainable set (adapters, temporal_embedding, ln_post)."""
    q = {k: v.clone().requires_grad_(O.is_trainable(k)) for k, v in p.items()}
    hw, hb = head_w.clone().requires_grad_(True), head_b.clone().requires_grad_(True)
    lg = O.head_logits(backbone(q, x, cfg, drop_masks), hw, hb)
    loss = F.cross_entropy(lg, labels)
    loss.backward()
    return loss.detach(), lg.detach(), {k: v.grad for k, v in q.items() if v.requires_grad}

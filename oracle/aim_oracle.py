"""CPU oracle for the AIM ``ViT_CLIP`` backbone hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-the-math restatement (plain torch on CPU, fp32 or fp64) of the
reference algorithm.  It is *not* product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline.  The product path
(``aimb200``) never routes through this module.

Parity pin: the reference's own tests hold no golden vector for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the *reference modules
themselves*, executed in the build container through ``oracle/ref_loader.py``
(``tests/golden/make_golden.py`` writes the fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the oracle against them on every run).

What it restates (reference file:line):
  * stem            vitclip_aim.py:445-459  (== vit_clip.py:433-447)
  * LayerNorm       vitclip_aim.py:97-103   (fp32 statistics, eps 1e-5)
  * attention       vitclip_aim.py:153-193  (q/k/v = row slices of in_proj_weight)
  * Adapter         vitclip_aim.py:78-95
  * QuickGELU/MLP   vitclip_aim.py:106-108, 119-123
  * block 'aim'     vitclip_aim.py:196-211  (wind_attn=False)
  * block 'fork'    vit_clip.py:199-288     (shift=False) incl. cross_attention :164-197
  * tail            vitclip_aim.py:463-468
  * head            mmaction/models/heads/i3d_head.py:53-73 (T-mean, fc_cls; eval: no dropout)
  * DropPath        timm 0.5.4 ``drop_path`` (external, pinned in requirements.txt:1):
                    mask shape (x.shape[0],1,1) on LND tensors => one Bernoulli draw per
                    *token index*, shared by all frames.

Layout used here: activations are ``[BT, n, D]`` (frame-major "NLD"), frame f = b*T + t,
token 0 = cls, token 1+gy*G+gx = patch (gy,gx).  The reference works in LND; the two are
permutations of each other and every op below is row-wise or per-sequence, so results
are identical up to fp rounding.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class OracleCfg:
    input_resolution: int = 224
    num_frames: int = 8
    patch_size: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    num_tadapter: int = 1
    adapter_scale: float = 0.5
    block: str = "aim"  # 'aim' | 'fork'

    @property
    def grid(self) -> int:
        return self.input_resolution // self.patch_size

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def hidden(self) -> int:
        return int(self.width * 0.25)


def param_shapes(cfg: OracleCfg) -> Dict[str, tuple]:
    """state_dict names/shapes of the reference module (probe dump)."""
    D, n, T, p, r = cfg.width, cfg.tokens, cfg.num_frames, cfg.patch_size, cfg.hidden
    s: Dict[str, tuple] = {
        "class_embedding": (D,),
        "positional_embedding": (n, D),
        "temporal_embedding": (1, T, D),
        "conv1.weight": (D, 3, p, p),
        "ln_pre.weight": (D,),
        "ln_pre.bias": (D,),
    }
    adapters = ["MLP_Adapter", "S_Adapter", "T_Adapter"]
    if cfg.num_tadapter == 2 and cfg.block == "aim":
        adapters.append("T_Adapter_in")
    for i in range(cfg.layers):
        pre = f"transformer.resblocks.{i}."
        s[pre + "attn.in_proj_weight"] = (3 * D, D)
        s[pre + "attn.in_proj_bias"] = (3 * D,)
        s[pre + "attn.out_proj.weight"] = (D, D)
        s[pre + "attn.out_proj.bias"] = (D,)
        s[pre + "ln_1.weight"] = (D,)
        s[pre + "ln_1.bias"] = (D,)
        s[pre + "mlp.c_fc.weight"] = (4 * D, D)
        s[pre + "mlp.c_fc.bias"] = (4 * D,)
        s[pre + "mlp.c_proj.weight"] = (D, 4 * D)
        s[pre + "mlp.c_proj.bias"] = (D,)
        s[pre + "ln_2.weight"] = (D,)
        s[pre + "ln_2.bias"] = (D,)
        for a in adapters:
            s[pre + a + ".D_fc1.weight"] = (r, D)
            s[pre + a + ".D_fc1.bias"] = (r,)
            s[pre + a + ".D_fc2.weight"] = (D, r)
            s[pre + a + ".D_fc2.bias"] = (D,)
    s["ln_post.weight"] = (D,)
    s["ln_post.bias"] = (D,)
    return s


def is_trainable(name: str) -> bool:
    """Freeze rule of vit_clip.py:413-415 / vitclip_aim.py:425-427."""
    return ("temporal_embedding" in name) or ("ln_post" in name) or ("Adapter" in name)


def fixture_state_dict(cfg: OracleCfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights, generated tensor by tensor from one CPU
    generator in ``param_shapes`` order (so the fixture does not depend on the
    reference's construction-time RNG order).  Magnitudes follow the reference's
    init (trunc-normal std .02 linears, width**-0.5 embeddings, LN ~ 1/0) but every
    bias, every adapter ``D_fc2`` and ``temporal_embedding`` are made non-zero, else
    the adapter/temporal paths would be dead at init."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    D = cfg.width
    for name, shape in param_shapes(cfg).items():
        if name in ("class_embedding", "positional_embedding"):
            t = torch.randn(shape, generator=g) * D ** -0.5
        elif name == "temporal_embedding":
            t = torch.randn(shape, generator=g) * 0.02
        elif name == "conv1.weight":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        elif name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name in ("ln_pre.weight", "ln_post.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias") and ("ln_" in name.split(".")[-2]):
            t = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:  # linear weights
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def fixture_clip(cfg: OracleCfg, batch: int, seed: int = 2, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    R = cfg.input_resolution
    return torch.randn(batch, 3, cfg.num_frames, R, R, generator=g).to(dtype)


def fixture_head(cfg: OracleCfg, num_classes: int = 400, seed: int = 3, dtype=torch.float32):
    """I3DHead fc_cls: normal(std=0.01) weight (i3d_head.py:49-51); bias made non-zero."""
    g = torch.Generator().manual_seed(seed)
    w = 0.01 * torch.randn(num_classes, cfg.width, generator=g)
    b = 0.01 * torch.randn(num_classes, generator=g)
    return w.to(dtype), b.to(dtype)


# ----------------------------------------------------------------------------- ops

def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(u):
    return u * torch.sigmoid(1.702 * u)


def gelu_erf(u):
    return 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))


def adapter(p, pre, x, skip: bool):
    h = gelu_erf(x @ p[pre + ".D_fc1.weight"].T + p[pre + ".D_fc1.bias"])
    y = h @ p[pre + ".D_fc2.weight"].T + p[pre + ".D_fc2.bias"]
    return x + y if skip else y


def mha_core(q, k, v, heads: int):
    """q [S, Lq, D], k/v [S, Lk, D]: S independent sequences; head h = cols h*dh:(h+1)*dh."""
    S, Lq, D = q.shape
    dh = D // heads
    qh = q.reshape(S, Lq, heads, dh).transpose(1, 2)
    kh = k.reshape(S, -1, heads, dh).transpose(1, 2)
    vh = v.reshape(S, -1, heads, dh).transpose(1, 2)
    aff = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    o = torch.softmax(aff, dim=-1) @ vh
    return o.transpose(1, 2).reshape(S, Lq, D), aff


def attention(p, pre, x, heads: int):
    """x [S, L, D] -> out_proj(softmax(q k^T/sqrt(dh)) v), q/k/v from in_proj rows 0:D/D:2D/2D:3D."""
    D = x.shape[-1]
    W, bW = p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"]
    qkv = x @ W.T + bW
    o, aff = mha_core(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], heads)
    return o @ p[pre + "attn.out_proj.weight"].T + p[pre + "attn.out_proj.bias"], aff, qkv


def _rows_to_temporal(x, B, T):
    """[B*T, n, D] -> [B*n, T, D] (sequences over t for each (b, token))."""
    BT, n, D = x.shape
    return x.reshape(B, T, n, D).permute(0, 2, 1, 3).reshape(B * n, T, D)


def _temporal_to_rows(x, B, n):
    Bn, T, D = x.shape
    return x.reshape(B, n, T, D).permute(0, 2, 1, 3).reshape(B * T, n, D)


def block_aim(p, i, x, cfg: OracleCfg, B: int, mask_t=None, mask_m=None):
    """vitclip_aim.py:196-211.  mask_* : per-token DropPath multipliers [n] (already /keep) or None."""
    pre = f"transformer.resblocks.{i}."
    T, n = cfg.num_frames, x.shape[1]
    ln1w, ln1b = p[pre + "ln_1.weight"], p[pre + "ln_1.bias"]
    xt = layer_norm(_rows_to_temporal(x, B, T), ln1w, ln1b)
    if cfg.num_tadapter == 2:
        xt = adapter(p, pre + "T_Adapter_in", xt, skip=True)
    at, _, _ = attention(p, pre, xt, cfg.heads)
    xt = _temporal_to_rows(adapter(p, pre + "T_Adapter", at, skip=False), B, n)
    if mask_t is not None:
        xt = xt * mask_t.view(1, n, 1)
    x = x + xt
    a_s, _, _ = attention(p, pre, layer_norm(x, ln1w, ln1b), cfg.heads)
    x = x + adapter(p, pre + "S_Adapter", a_s, skip=True)
    xn = layer_norm(x, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"])
    mlp = quick_gelu(xn @ p[pre + "mlp.c_fc.weight"].T + p[pre + "mlp.c_fc.bias"]) @ p[pre + "mlp.c_proj.weight"].T \
        + p[pre + "mlp.c_proj.bias"]
    ad = cfg.adapter_scale * adapter(p, pre + "MLP_Adapter", xn, skip=False)
    if mask_m is not None:
        ad = ad * mask_m.view(1, n, 1)
    return x + mlp + ad


def block_fork(p, i, x, cfg: OracleCfg, B: int, mask_s=None, mask_m=None):
    """vit_clip.py:199-288 with shift=False (cls-only temporal attention, lambda-mixed cross attention)."""
    pre = f"transformer.resblocks.{i}."
    T, n, D = cfg.num_frames, x.shape[1], x.shape[2]
    ln1w, ln1b = p[pre + "ln_1.weight"], p[pre + "ln_1.bias"]
    W, bW = p[pre + "attn.in_proj_weight"], p[pre + "attn.in_proj_bias"]
    Wo, bo = p[pre + "attn.out_proj.weight"], p[pre + "attn.out_proj.bias"]
    # temporal attention over the cls token of the T frames of each clip  (:218-229)
    ct = x[:, 0, :].reshape(B, T, D)
    at, _, _ = attention(p, pre, layer_norm(ct, ln1w, ln1b), cfg.heads)
    xt = adapter(p, pre + "T_Adapter", at, skip=False).reshape(B * T, 1, D)      # kept aside, NOT added to x
    # spatial self attention + weights (:264, :147-151)
    xn1 = layer_norm(x, ln1w, ln1b)
    a_o, aff_o, qkv = attention(p, pre, xn1, cfg.heads)
    w_o = torch.exp(aff_o.sum(1)).reshape(B * T, -1).sum(-1)                      # [BT]
    # cross attention to the single temporal cls key of the frame (:265, :164-197); k,v from raw xt
    q = qkv[..., :D]
    k = xt @ W[D:2 * D].T + bW[D:2 * D]
    v = xt @ W[2 * D:].T + bW[2 * D:]
    o_c, aff_c = mha_core(q, k, v, cfg.heads)
    a_c = o_c @ Wo.T + bo
    w_c = torch.exp(aff_c.sum(1)).reshape(B * T, -1).sum(-1)
    lam = (w_c / (w_c + w_o)).detach().view(B * T, 1, 1)                           # computed under no_grad
    sa = cfg.adapter_scale * adapter(p, pre + "S_Adapter", lam * a_c, skip=False)
    if mask_s is not None:
        sa = sa * mask_s.view(1, n, 1)
    x = x + (1 - lam) * a_o + sa
    xn = layer_norm(x, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"])
    mlp = quick_gelu(xn @ p[pre + "mlp.c_fc.weight"].T + p[pre + "mlp.c_fc.bias"]) @ p[pre + "mlp.c_proj.weight"].T \
        + p[pre + "mlp.c_proj.bias"]
    ad = cfg.adapter_scale * adapter(p, pre + "MLP_Adapter", xn, skip=False)
    if mask_m is not None:
        ad = ad * mask_m.view(1, n, 1)
    return x + mlp + ad


def stem(p, x, cfg: OracleCfg):
    """vitclip_aim.py:445-459: patch-embed conv (k=s=p, no bias) as an explicit im2col GEMM,
    cls prepend, +pos, +temporal, ln_pre.  Returns [BT, n, D]."""
    B, C, T, H, W = x.shape
    ps, G, D = cfg.patch_size, cfg.grid, cfg.width
    fr = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    # im2col: row (f, gy, gx), column c*p*p + ky*p + kx   (bit-exact indexing contract)
    cols = fr.reshape(B * T, C, G, ps, G, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * T, G * G, C * ps * ps)
    tok = cols @ p["conv1.weight"].reshape(D, -1).T
    cls = p["class_embedding"].view(1, 1, D).expand(B * T, 1, D)
    z = torch.cat([cls, tok], 1) + p["positional_embedding"].view(1, -1, D)
    z = (z.reshape(B, T, -1, D) + p["temporal_embedding"].view(1, T, 1, D)).reshape(B * T, -1, D)
    return layer_norm(z, p["ln_pre.weight"], p["ln_pre.bias"])


def backbone(p, x, cfg: OracleCfg, drop_masks: Optional[List] = None, taps: Optional[dict] = None):
    """Full backbone: x [B,3,T,H,W] -> [B, D, T, 1, 1].  drop_masks: per block (mask_a, mask_m) or None."""
    B, T = x.shape[0], cfg.num_frames
    assert x.shape[2] == T
    z = stem(p, x, cfg)
    if taps is not None:
        taps["stem"] = z
    for i in range(cfg.layers):
        ma, mm = (None, None) if drop_masks is None else drop_masks[i]
        if cfg.block == "aim":
            z = block_aim(p, i, z, cfg, B, ma, mm)
        else:
            z = block_fork(p, i, z, cfg, B, ma, mm)
        if taps is not None:
            taps[f"block{i}"] = z
    cls = layer_norm(z[:, 0, :], p["ln_post.weight"], p["ln_post.bias"])      # ln_post is row-wise: cls rows suffice
    return cls.reshape(B, T, -1).permute(0, 2, 1).unsqueeze(-1).unsqueeze(-1)


def head_logits(feat, w, b):
    """I3DHead eval forward (i3d_head.py:53-73): mean over (T,H,W) then fc_cls."""
    return feat.mean(dim=(2, 3, 4)) @ w.T + b


def logits(p, x, cfg: OracleCfg, head_w, head_b, drop_masks=None):
    return head_logits(backbone(p, x, cfg, drop_masks), head_w, head_b)


def loss_and_grads(p, x, labels, cfg: OracleCfg, head_w, head_b, drop_masks=None):
    """CE loss of the head logits and grads wrt the trainable set (+ head).  Autograd on the restatement."""
    q = {k: (v.clone().requires_grad_(is_trainable(k))) for k, v in p.items()}
    hw, hb = head_w.clone().requires_grad_(True), head_b.clone().requires_grad_(True)
    lg = logits(q, x, cfg, hw, hb, drop_masks)
    loss = F.cross_entropy(lg, labels)
    loss.backward()
    grads = {k: v.grad for k, v in q.items() if v.requires_grad}
    grads["cls_head.fc_cls.weight"], grads["cls_head.fc_cls.bias"] = hw.grad, hb.grad
    return loss.detach(), lg.detach(), grads


def normalised_max_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a-ref| / max|ref|  (element-wise relative error is ill-defined on near-zero logits)."""
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))

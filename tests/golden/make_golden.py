"""Generate the golden fixtures under tests/golden/ by running the REAL reference modules
(/root/reference, imported through oracle/ref_loader.py) on the deterministic fixture
weights/clips of oracle/aim_oracle.py.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (committed):
  tiny_<block>[_nt2].npz : fp64 logits, loss and every trainable gradient of a 2-layer,
                           width-256 model (full tensors; small)
  vitb16_8x224_<block>.npz : cfg1 of BASELINE.json (ViT-B/16, 8x224, batch 1): fp32 and fp64
                           logits of the reference, loss, per-block cls-token taps (frame 0)
                           and per-tensor gradient summaries (L2 norm, sum, first 8 values).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aim_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY = dict(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)


def _ref(cfg, p, dt):
    m = R.reference_module(cfg, p).to(dt)
    if dt == torch.float64:  # the reference LayerNorm subclass hard-casts to fp32 (vit_clip.py:74-77)
        for mod in m.modules():
            if isinstance(mod, nn.LayerNorm):
                mod.forward = nn.LayerNorm.forward.__get__(mod)
    return m.eval()


def _taps(m, x):
    """cls token of frame 0 after every block (LND tensor row [0, 0, :])."""
    taps = []
    hooks = [blk.register_forward_hook(lambda _m, _i, o: taps.append(o[0, 0, :].detach().clone()))
             for blk in m.transformer.resblocks]
    feat = m(x)
    for h in hooks:
        h.remove()
    return feat, torch.stack(taps)


def make_tiny(block, nt):
    cfg = O.OracleCfg(**TINY, block=block, num_tadapter=nt)
    dt = torch.float64
    p = O.fixture_state_dict(cfg, dtype=dt)
    x = O.fixture_clip(cfg, 2, dtype=dt)
    hw, hb = O.fixture_head(cfg, 16, dtype=dt)
    labels = torch.tensor([3, 11])
    m = _ref(cfg, p, dt)
    feat, taps = _taps(m, x)
    lg = O.head_logits(feat, hw, hb)
    loss = F.cross_entropy(lg, labels)
    loss.backward()
    out = {"logits": lg.detach().numpy(), "loss": loss.detach().numpy(), "taps": taps.numpy(),
           "feat": feat.detach().numpy(), "labels": labels.numpy()}
    for name, prm in m.named_parameters():
        if prm.requires_grad:
            out["grad/" + name] = prm.grad.numpy().astype(np.float32)
    tag = f"tiny_{block}" + ("_nt2" if nt == 2 else "")
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, "loss", float(loss), "n_grads", sum(k.startswith("grad/") for k in out))


def make_full(block):
    cfg = O.OracleCfg(block=block)
    out = {}
    labels = torch.tensor([7])
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        p = O.fixture_state_dict(cfg, dtype=dt)
        x = O.fixture_clip(cfg, 1, dtype=dt)
        hw, hb = O.fixture_head(cfg, 400, dtype=dt)
        m = _ref(cfg, p, dt)
        feat, taps = _taps(m, x)
        lg = O.head_logits(feat, hw, hb)
        loss = F.cross_entropy(lg, labels)
        loss.backward()
        out[f"logits_{tag}"] = lg.detach().numpy()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"taps_{tag}"] = taps.numpy()
        if dt == torch.float64:
            names, norms, sums, heads = [], [], [], []
            for name, prm in m.named_parameters():
                if prm.requires_grad:
                    g = prm.grad.reshape(-1)
                    names.append(name)
                    norms.append(float(g.norm()))
                    sums.append(float(g.sum()))
                    heads.append(g[:8].numpy().copy())
            out["grad_names"] = np.array(names)
            out["grad_norm"] = np.array(norms)
            out["grad_sum"] = np.array(sums)
            out["grad_head"] = np.stack(heads)
            # two complete gradients for element-wise checks (first and last block's T_Adapter fc1, temporal emb)
            for name, prm in m.named_parameters():
                if name in ("temporal_embedding", "ln_post.weight", "ln_post.bias",
                            "transformer.resblocks.0.T_Adapter.D_fc1.bias",
                            "transformer.resblocks.11.MLP_Adapter.D_fc2.bias"):
                    out["grad/" + name] = prm.grad.numpy()
    out["labels"] = labels.numpy()
    np.savez_compressed(os.path.join(HERE, f"vitb16_8x224_{block}.npz"), **out)
    print("full", block, "loss", float(out["loss_f64"]),
          "f32-vs-f64 logit err", O.normalised_max_err(torch.tensor(out["logits_f32"]), torch.tensor(out["logits_f64"])))


def make_tiny_imagenet(nt=1):
    """vit_imagenet.py::ViT_ImageNet (2 layers, width 256): fp64 logits, loss, features and the gradients of the AIM
    trainable set (the reference freezes nothing; gradients of the adapters do not depend on that)."""
    from oracle import imagenet_oracle as OI
    cfg = O.OracleCfg(**TINY, num_tadapter=nt)
    dt = torch.float64
    p = OI.fixture_state_dict(cfg, dtype=dt)
    x = O.fixture_clip(cfg, 2, dtype=dt)
    hw, hb = O.fixture_head(cfg, 16, dtype=dt)
    labels = torch.tensor([3, 11])
    m = R.reference_imagenet(cfg, p).to(dt).eval()
    feat = m(x)
    lg = O.head_logits(feat, hw, hb)
    loss = F.cross_entropy(lg, labels)
    loss.backward()
    out = {"logits": lg.detach().numpy(), "loss": loss.detach().numpy(), "feat": feat.detach().numpy(), "labels": labels.numpy(),
           "state_dict_keys": np.array(list(m.state_dict().keys()))}
    for name, prm in m.named_parameters():
        if O.is_trainable(name):
            out["grad/" + name] = prm.grad.numpy().astype(np.float32)
    tag = "tiny_imagenet" + ("_nt2" if nt == 2 else "")
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, "loss", float(loss), "n_grads", sum(k.startswith("grad/") for k in out))


if __name__ == "__main__":
    assert R.available(), "needs /root/reference"
    torch.manual_seed(0)
    if "--imagenet-only" in sys.argv:
        make_tiny_imagenet(1)
        make_tiny_imagenet(2)
        sys.exit(0)
    make_tiny("aim", 1)
    make_tiny("aim", 2)
    make_tiny("fork", 1)
    make_full("aim")
    make_full("fork")
    make_tiny_imagenet(1)
    make_tiny_imagenet(2)

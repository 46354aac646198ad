"""CPU: the oracle (oracle/aim_oracle.py) against the golden vectors generated from the REAL reference
(tests/golden/make_golden.py), and — when /root/reference is present — against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import aim_oracle as O
from oracle import ref_loader as R

G = os.path.join(os.path.dirname(__file__), "golden")
TINY = dict(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)


@pytest.mark.parametrize("block,nt", [("aim", 1), ("aim", 2), ("fork", 1)])
def test_oracle_matches_golden_tiny(block, nt):
    gold = np.load(os.path.join(G, f"tiny_{block}" + ("_nt2" if nt == 2 else "") + ".npz"))
    cfg = O.OracleCfg(**TINY, block=block, num_tadapter=nt)
    dt = torch.float64
    p = O.fixture_state_dict(cfg, dtype=dt)
    x = O.fixture_clip(cfg, 2, dtype=dt)
    hw, hb = O.fixture_head(cfg, 16, dtype=dt)
    loss, lg, grads = O.loss_and_grads(p, x, torch.tensor(gold["labels"]), cfg, hw, hb)
    assert O.normalised_max_err(lg, torch.tensor(gold["logits"])) < 1e-12
    assert abs(float(loss) - float(gold["loss"])) < 1e-12
    n = 0
    for k in gold.files:
        if k.startswith("grad/"):
            n += 1
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < 1e-6, k
    assert n == (35 if nt == 2 else 27)
    taps = {}
    O.backbone(p, x, cfg, taps=taps)
    for i in range(cfg.layers):   # reference tap = LND[0, 0, :] = frame 0, cls token
        assert O.normalised_max_err(taps[f"block{i}"][0, 0], torch.tensor(gold["taps"][i])) < 1e-12


@pytest.mark.parametrize("block", ["aim", "fork"])
def test_oracle_matches_golden_vitb16(block):
    """cfg1 of BASELINE.json: ViT-B/16 8x224 single clip."""
    gold = np.load(os.path.join(G, f"vitb16_8x224_{block}.npz"))
    cfg = O.OracleCfg(block=block)
    p = O.fixture_state_dict(cfg)
    x = O.fixture_clip(cfg, 1)
    hw, hb = O.fixture_head(cfg, 400)
    with torch.no_grad():
        lg = O.logits(p, x, cfg, hw, hb)
    assert O.normalised_max_err(lg, torch.tensor(gold["logits_f64"])) < 5e-6
    assert O.normalised_max_err(lg, torch.tensor(gold["logits_f32"])) < 5e-6
    assert int(lg.argmax()) == int(gold["logits_f64"].argmax())


def test_param_tree_matches_reference_dump():
    cfg = O.OracleCfg()
    shapes = O.param_shapes(cfg)
    assert len(shapes) == 296                                  # 147 trainable of 296 tensors
    assert sum(O.is_trainable(k) for k in shapes) == 147
    n_train = sum(int(np.prod(s)) for k, s in shapes.items() if O.is_trainable(k))
    assert abs(n_train / 1e6 - 10.659) < 0.001


@pytest.mark.skipif(not R.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("block", ["aim", "fork"])
def test_oracle_matches_live_reference_with_droppath(block):
    """Training mode with DropPath: same per-token masks fed to both (reference draws them from torch RNG)."""
    cfg = O.OracleCfg(**TINY, block=block)
    dt = torch.float64
    p = O.fixture_state_dict(cfg, dtype=dt)
    x = O.fixture_clip(cfg, 2, dtype=dt)
    m = R.reference_module(cfg, p, drop_path_rate=0.5).to(dt)
    import torch.nn as nn
    for mod in m.modules():
        if isinstance(mod, nn.LayerNorm):
            mod.forward = nn.LayerNorm.forward.__get__(mod)
    m.train()
    torch.manual_seed(11)
    ref = m(x)
    # replay the RNG stream: block 0 has rate 0 (Identity); block 1 draws two masks of n tokens
    torch.manual_seed(11)
    keep = 0.5
    n = cfg.tokens
    m1 = torch.empty(n, 1, 1, dtype=dt).bernoulli_(keep).div_(keep).view(n)
    m2 = torch.empty(n, 1, 1, dtype=dt).bernoulli_(keep).div_(keep).view(n)
    out = O.backbone(p, x, cfg, drop_masks=[(None, None), (m1, m2)])
    assert O.normalised_max_err(out, ref.detach()) < 1e-12


# ------------------------------------------------------------------------------------------------ ViT_ImageNet variant (f4)
@pytest.mark.parametrize("nt", [1, 2])
def test_imagenet_oracle_matches_golden_tiny(nt):
    """oracle/imagenet_oracle.py against vit_imagenet.py::ViT_ImageNet (goldens generated from the real class)."""
    from oracle import imagenet_oracle as OI
    gold = np.load(os.path.join(G, "tiny_imagenet" + ("_nt2" if nt == 2 else "") + ".npz"))
    cfg = O.OracleCfg(**TINY, num_tadapter=nt)
    dt = torch.float64
    p = OI.fixture_state_dict(cfg, dtype=dt)
    assert list(p) and sorted(p) == sorted(str(k) for k in gold["state_dict_keys"])
    x = O.fixture_clip(cfg, 2, dtype=dt)
    hw, hb = O.fixture_head(cfg, 16, dtype=dt)
    loss, lg, grads = OI.loss_and_grads(p, x, torch.tensor(gold["labels"]), cfg, hw, hb)
    assert O.normalised_max_err(lg, torch.tensor(gold["logits"])) < 1e-12
    assert abs(float(loss) - float(gold["loss"])) < 1e-12
    n = 0
    for k in gold.files:
        if k.startswith("grad/"):
            n += 1
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < 1e-6, k
    assert n == (35 if nt == 2 else 27) == len(grads)


@pytest.mark.skipif(not R.imagenet_available(), reason="/root/reference not mounted (GPU box)")
def test_imagenet_oracle_matches_live_reference_with_droppath():
    """timm DropPath on [(b t), n, d] tensors masks FRAMES (one draw per (b, t)), unlike the LND CLIP variant (tokens)."""
    from oracle import imagenet_oracle as OI
    cfg = O.OracleCfg(**TINY)
    dt = torch.float64
    p = OI.fixture_state_dict(cfg, dtype=dt)
    x = O.fixture_clip(cfg, 2, dtype=dt)
    m = R.reference_imagenet(cfg, p, drop_path_rate=0.5).to(dt).train()
    torch.manual_seed(11)
    ref = m(x)
    torch.manual_seed(11)
    keep, BT = 0.5, 2 * cfg.num_frames
    m1 = torch.empty(BT, 1, 1, dtype=dt).bernoulli_(keep).div_(keep).view(BT)
    m2 = torch.empty(BT, 1, 1, dtype=dt).bernoulli_(keep).div_(keep).view(BT)
    out = OI.backbone(p, x, cfg, drop_masks=[(None, None), (m1, m2)])
    assert O.normalised_max_err(out, ref.detach()) < 1e-12

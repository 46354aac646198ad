"""Recognizer glue (f1) and checkpoint compatibility (f3): CPU for the host logic, GPU for the end-to-end call."""
import pytest
import torch
import torch.nn.functional as F

import aimb200
from aimb200.recognizer import top_k_accuracy_device
from oracle import aim_oracle as O

MODEL = dict(type='Recognizer3D',
             backbone=dict(type='ViT_CLIP', input_resolution=64, patch_size=16, num_frames=4, width=256, layers=2, heads=4,
                           drop_path_rate=0.0),
             cls_head=dict(type='I3DHead', in_channels=256, num_classes=16, spatial_type='avg', dropout_ratio=0.5),
             test_cfg=dict(average_clips='prob'))


def _build():
    cfg = dict(MODEL)
    cfg.pop("type")
    return aimb200.Recognizer3D(**cfg)


def test_topk_and_average_clip_cpu():
    s = torch.tensor([[0.1, 0.7, 0.2], [0.5, 0.3, 0.2], [0.2, 0.3, 0.5]])
    y = torch.tensor([1, 1, 0])
    t1, t2 = top_k_accuracy_device(s, y, (1, 2))
    assert float(t1) == pytest.approx(1 / 3) and float(t2) == pytest.approx(2 / 3)
    m = _build()
    sc = torch.randn(6, 16)
    prob = m.average_clip(sc, 3)
    assert torch.allclose(prob, F.softmax(sc.view(2, 3, 16), 2).mean(1))
    m.average_clips = "score"
    assert torch.allclose(m.average_clip(sc, 3), sc.view(2, 3, 16).mean(1))
    m.average_clips = "bogus"
    with pytest.raises(ValueError):
        m.average_clip(sc, 3)
    with pytest.raises(ValueError, match="Label should not be None"):
        m(torch.zeros(1, 1, 3, 4, 64, 64))


def test_checkpoint_roundtrip_cpu(tmp_path):
    m = _build()
    cfg = O.OracleCfg(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)
    m.backbone.load_state_dict(O.fixture_state_dict(cfg))
    # mmcv-style full checkpoint with DDP prefix
    full = {"meta": {}, "state_dict": {"module." + k: v.clone() for k, v in m.state_dict().items()}, "optimizer": {}}
    p = tmp_path / "latest.pth"
    torch.save(full, p)
    m2 = _build()
    res = aimb200.load_checkpoint(m2, str(p), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])
    # adapters-only checkpoint: ~10x smaller, restores exactly the trained set
    small = aimb200.trainable_state_dict(m)
    assert all(("Adapter" in k) or ("ln_post" in k) or ("temporal_embedding" in k) or ("cls_head" in k) for k in small)
    n_small = sum(v.numel() for v in small.values())
    n_full = sum(v.numel() for v in m.state_dict().values())
    assert n_small < 0.25 * n_full
    m3 = _build()
    res = aimb200.load_checkpoint(m3, small)
    assert not res.unexpected_keys
    for k in small:
        assert torch.equal(m3.state_dict()[k], m.state_dict()[k])
    # CLIP-style visual dict (bare backbone keys + 'proj') loads into the recognizer's backbone
    clip_like = {"visual." + k: v for k, v in O.fixture_state_dict(cfg, seed=7).items()}
    clip_like["visual.proj"] = torch.zeros(4, 4)
    res = aimb200.load_checkpoint(m3, clip_like)
    assert not res.unexpected_keys
    assert torch.equal(m3.backbone.conv1.weight, O.fixture_state_dict(cfg, seed=7)["conv1.weight"])


@pytest.mark.gpu
def test_recognizer_train_and_test_step_gpu():
    cfg = O.OracleCfg(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)
    m = _build()
    p = O.fixture_state_dict(cfg)
    m.backbone.load_state_dict(p)
    m = m.cuda()
    g = torch.Generator().manual_seed(1)
    imgs = torch.randn(2, 1, 3, 4, 64, 64, generator=g)
    labels = torch.tensor([[3], [11]])
    m.eval()                                       # dropout off -> comparable with the oracle
    out = m(imgs.cuda(), labels.cuda())
    out["loss_cls"].backward()
    hw, hb = m.cls_head.fc_cls.weight.detach().cpu(), m.cls_head.fc_cls.bias.detach().cpu()
    ref_loss, ref_lg, ref_g = O.loss_and_grads(p, imgs.reshape(2, 3, 4, 64, 64), labels.reshape(-1), cfg, hw, hb)
    assert abs(float(out["loss_cls"]) - float(ref_loss)) < 2e-2
    k = "transformer.resblocks.1.MLP_Adapter.D_fc1.weight"
    assert O.normalised_max_err(dict(m.backbone.named_parameters())[k].grad.cpu(), ref_g[k]) < 6e-2
    assert out["top1_acc"].is_cuda and out["top1_acc"].dim() == 0     # stays on the device: no host sync
    views = torch.randn(2, 3, 3, 4, 64, 64, generator=g)
    prob = m(views.cuda(), return_loss=False)
    assert prob.shape == (2, 16) and torch.allclose(prob.sum(1).cpu(), torch.ones(2), atol=1e-4)
    ref = O.head_logits(O.backbone(p, views.reshape(6, 3, 4, 64, 64), cfg), hw, hb)
    ref = torch.softmax(ref, -1).view(2, 3, 16).mean(1)
    assert O.normalised_max_err(prob.cpu(), ref) < 2e-2


@pytest.mark.gpu
def test_flat_adamw_matches_torch_adamw():
    """aimb200.FlatAdamW (one kernel over the flat trainable buffer) against torch.optim.AdamW with the same groups, three
    training steps of the tiny backbone: every parameter agrees."""
    import aimb200
    from oracle import aim_oracle as O
    cfg = O.OracleCfg(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)
    x = O.fixture_clip(cfg, 2).cuda()
    hw, hb = O.fixture_head(cfg, 16)
    labels = torch.tensor([3, 11]).cuda()

    def run(flat: bool):
        m = aimb200.build_backbone(dict(type="ViT_CLIP", input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2,
                                        heads=4, drop_path_rate=0.0, compute_dtype="fp32"))
        m.init_weights()
        m.load_state_dict(O.fixture_state_dict(cfg))
        m = m.cuda().train()
        w, b = hw.clone().cuda().requires_grad_(True), hb.clone().cuda().requires_grad_(True)
        named = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
        dec = [p for n, p in named if "Adapter" in n and n.endswith("weight")]
        nod = [p for n, p in named if not ("Adapter" in n and n.endswith("weight"))]
        if flat:
            head = torch.optim.AdamW([{"params": [w], "weight_decay": 0.05}, {"params": [b], "weight_decay": 0.0}], lr=1e-3)
            opt = aimb200.FlatAdamW(m, lr=1e-3, weight_decay=0.05, extra=head)
        else:
            opt = torch.optim.AdamW([{"params": dec + [w], "weight_decay": 0.05}, {"params": nod + [b], "weight_decay": 0.0}], lr=1e-3)
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(O.head_logits(m(x), w, b), labels)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        return {n: p.detach().clone() for n, p in named}, w.detach().clone(), float(loss)

    pa, wa, la = run(True)
    pb, wb, lb = run(False)
    assert abs(la - lb) < 1e-5
    assert torch.allclose(wa, wb, atol=1e-6)
    worst = max(float((pa[n] - pb[n]).abs().max()) for n in pa)
    moved = max(float((pa[n] - O.fixture_state_dict(cfg)[n].cuda()).abs().max()) for n in pa)
    assert moved > 1e-3 and worst < 2e-6, (worst, moved)

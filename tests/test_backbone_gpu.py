"""Model-level parity on a real B200: aimb200.ViT_CLIP (CUDA path through the C ABI) against
 (1) the CPU oracle on the same seeded fixture and (2) the committed golden vectors produced by the
 real reference.  Gates: fp32 mode max|d|/max|ref| <= 1e-3 on logits, bf16 mode
 <= 2e-2 with identical top-1; gradients of the trainable set: <= 1e-3 (fp32), <= 6e-2 (bf16)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import aimb200
from oracle import aim_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TINY = dict(input_resolution=64, num_frames=4, patch_size=16, width=256, layers=2, heads=4)


def _build(cfg: O.OracleCfg, mode: str, drop_path_rate=0.0):
    m = aimb200.build_backbone(dict(type="ViT_CLIP", input_resolution=cfg.input_resolution, num_frames=cfg.num_frames,
                                    patch_size=cfg.patch_size, width=cfg.width, layers=cfg.layers, heads=cfg.heads,
                                    drop_path_rate=drop_path_rate, num_tadapter=cfg.num_tadapter,
                                    adapter_scale=cfg.adapter_scale, compute_dtype=mode, block=cfg.block))
    m.init_weights()
    m.load_state_dict(O.fixture_state_dict(cfg))
    return m.cuda()


def _cuda_logits_and_grads(m, cfg, x, hw, hb, labels):
    hwc, hbc = hw.cuda().requires_grad_(True), hb.cuda().requires_grad_(True)
    m.train()   # drop_path_rate=0 -> deterministic; exercises the saved-activation path
    feat = m(x.cuda())
    lg = O.head_logits(feat, hwc, hbc)
    loss = F.cross_entropy(lg, labels.cuda())
    loss.backward()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.requires_grad}
    return lg.detach().cpu(), float(loss), grads


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("block,nt", [("aim", 1), ("aim", 2), ("fork", 1)])
def test_tiny_logits_and_all_grads_vs_golden(mode, block, nt):
    gold = np.load(os.path.join(G, f"tiny_{block}" + ("_nt2" if nt == 2 else "") + ".npz"))
    cfg = O.OracleCfg(**TINY, block=block, num_tadapter=nt)
    m = _build(cfg, mode)
    x = O.fixture_clip(cfg, 2)
    hw, hb = O.fixture_head(cfg, 16)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, torch.tensor(gold["labels"]))
    tol_l, tol_g = (1e-3, 1e-3) if mode == "fp32" else (2e-2, 6e-2)
    assert O.normalised_max_err(lg, torch.tensor(gold["logits"])) < tol_l
    assert abs(loss - float(gold["loss"])) < tol_l
    n = 0
    for k in gold.files:
        if k.startswith("grad/"):
            n += 1
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < tol_g, k
    assert n == len(grads)
    # eval-mode (inference buffers) forward agrees with the training-mode forward
    m.eval()
    with torch.no_grad():
        lg2 = O.head_logits(m(x.cuda()), hw.cuda(), hb.cuda()).cpu()
    assert O.normalised_max_err(lg2, lg) < (1e-6 if mode == "fp32" else 5e-3)  # bf16: eval skips the pre-activation rounding


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("block", ["aim", "fork"])
def test_vitb16_8x224_logits_vs_golden_and_oracle(mode, block):
    """cfg1 of BASELINE.json (the reference correctness fixture): ViT-B/16, 8x224, batch 1.
    block='aim' is checked against the reference class AIM (vitclip_aim.py), block='fork' against the in-tree
    ViT_CLIP (vit_clip.py) — goldens generated from each."""
    gold = np.load(os.path.join(G, f"vitb16_8x224_{block}.npz"))
    cfg = O.OracleCfg(block=block)
    m = _build(cfg, mode)
    x = O.fixture_clip(cfg, 1)
    hw, hb = O.fixture_head(cfg, 400)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, torch.tensor(gold["labels"]))
    ref = torch.tensor(gold["logits_f64"])
    err = O.normalised_max_err(lg, ref)
    print(f"[{mode}] logit err vs reference fp64 = {err:.3e}; loss {loss:.6f} vs {float(gold['loss_f64']):.6f}")
    assert err < (1e-3 if mode == "fp32" else 2e-2)
    assert int(lg.argmax()) == int(ref.argmax())
    # gradient summaries of all 147 trainable tensors + 5 full tensors
    names = [str(s) for s in gold["grad_names"]]
    assert sorted(names) == sorted(grads)
    tol = 1e-3 if mode == "fp32" else 6e-2
    worst = 0.0
    for i, k in enumerate(names):
        g = grads[k].double().reshape(-1)
        e = abs(float(g.norm()) - gold["grad_norm"][i]) / max(gold["grad_norm"][i], 1e-30)
        worst = max(worst, e)
        assert e < tol, (k, e)
    for k in gold.files:
        if k.startswith("grad/"):
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < tol, k
    print(f"[{mode}] worst grad-norm rel err over 147 tensors = {worst:.3e}")
    # the live oracle on this box gives the same answer as the committed golden (oracle not drifting)
    p = O.fixture_state_dict(cfg)
    with torch.no_grad():
        lo = O.logits(p, x, cfg, hw, hb)
    assert O.normalised_max_err(lo, ref) < 5e-6


@pytest.mark.parametrize("env", [dict(), dict(AIMB200_FUSE_T_OUTPROJ="0"), dict(AIMB200_PAIR_MLP="0"), dict(AIMB200_FUSE_S_OUTPROJ="0"),
                                 dict(AIMB200_FUSE_T_OUTPROJ="0", AIMB200_FUSE_S_OUTPROJ="0", AIMB200_PAIR_MLP="0"),
                                 dict(AIMB200_WGRAD_STREAM="0"), dict(AIMB200_LN_FOLD="0")])
def test_vitb16_launch_fusions_vs_golden(env, monkeypatch):
    """Every launch-count variant of the block (temporal out_proj folded into T_Adapter.D_fc1 with batched per-step weight
    products; MLP-adapter GEMMs riding on c_fc / c_proj as N- / K-concatenated segments; side streams on / off) gives the
    reference's logits and all 147 gradients within the bf16 gates."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    gold = np.load(os.path.join(G, "vitb16_8x224_aim.npz"))
    cfg = O.OracleCfg(block="aim")
    m = _build(cfg, "bf16")
    x = O.fixture_clip(cfg, 1)
    hw, hb = O.fixture_head(cfg, 400)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, torch.tensor(gold["labels"]))
    eng = m._engine
    assert eng.t_fused == (env.get("AIMB200_FUSE_T_OUTPROJ", "1") == "1")
    assert eng.s_fused == (env.get("AIMB200_FUSE_S_OUTPROJ", "1") == "1")
    assert eng.pair_mlp == (env.get("AIMB200_PAIR_MLP", "1") == "1")
    ref = torch.tensor(gold["logits_f64"])
    assert O.normalised_max_err(lg, ref) < 2e-2 and int(lg.argmax()) == int(ref.argmax())
    names = [str(s) for s in gold["grad_names"]]
    for i, k in enumerate(names):
        g = grads[k].double().reshape(-1)
        e = abs(float(g.norm()) - gold["grad_norm"][i]) / max(gold["grad_norm"][i], 1e-30)
        assert e < 6e-2, (k, e)
    for k in gold.files:
        if k.startswith("grad/"):
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < 6e-2, k
    m.eval()
    with torch.no_grad():
        lg2 = O.head_logits(m(x.cuda()), hw.cuda(), hb.cuda()).cpu()
    assert O.normalised_max_err(lg2, lg) < 5e-3


def test_bf16_top1_identical_batch4():
    cfg = O.OracleCfg(block="aim")
    m = _build(cfg, "bf16").eval()
    x = O.fixture_clip(cfg, 4, seed=5)
    hw, hb = O.fixture_head(cfg, 400)
    with torch.no_grad():
        lg = O.head_logits(m(x.cuda()), hw.cuda(), hb.cuda()).cpu()
        ref = O.logits(O.fixture_state_dict(cfg), x, cfg, hw, hb)
    assert O.normalised_max_err(lg, ref) < 2e-2
    assert torch.equal(lg.argmax(1), ref.argmax(1))


@pytest.mark.parametrize("block", ["aim", "fork"])
def test_droppath_training_matches_oracle_with_same_masks(block):
    cfg = O.OracleCfg(**TINY, block=block)
    m = _build(cfg, "fp32", drop_path_rate=0.5).train()
    x = O.fixture_clip(cfg, 2)
    torch.manual_seed(3)
    masks = m._drop_masks(m._dims(2), torch.device("cuda"))
    torch.manual_seed(3)
    feat = m(x.cuda())
    om = [(None, None) if a is None else (a.cpu(), b.cpu()) for a, b in masks]
    ref = O.backbone(O.fixture_state_dict(cfg), x, cfg, drop_masks=om)
    assert any(a is not None and float(a.min()) == 0.0 for a, _ in masks), "mask with a dropped token expected"
    assert O.normalised_max_err(feat.detach().cpu(), ref) < 1e-3


def test_uint8_input_normalisation_fused():
    cfg = O.OracleCfg(**TINY, block="aim")
    m = _build(cfg, "fp32").eval()
    mean, std = [122.769, 116.74, 104.04], [68.493, 66.63, 70.321]
    m.set_input_normalization(mean, std)
    g = torch.Generator().manual_seed(9)
    xu = torch.randint(0, 256, (2, 3, 4, 64, 64), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        feat = m(xu.cuda()).cpu()
    xf = (xu.float() - torch.tensor(mean).view(1, 3, 1, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1, 1)
    ref = O.backbone(O.fixture_state_dict(cfg), xf, cfg)
    assert O.normalised_max_err(feat, ref) < 1e-3


def test_edge_shapes_and_errors():
    cfg = O.OracleCfg(**TINY, block="aim")
    m = _build(cfg, "bf16").eval()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 5, 64, 64, device="cuda"))      # T != num_frames (vit_clip.py:344,444)
    with torch.no_grad():
        out = m(torch.zeros(0, 3, 4, 64, 64, device="cuda"))  # empty batch
    assert out.shape == (0, 256, 4, 1, 1)
    with torch.no_grad():
        out = m(torch.randn(3, 3, 4, 64, 64, device="cuda"))
    assert out.shape == (3, 256, 4, 1, 1) and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kw,batch", [
    # cfg4 / cfg5 shapes of BASELINE.json at reduced depth: ViT-L/14 (patch 14 -> K = 588 padded to 640, 257 tokens,
    # width 1024, 16 heads, r = 256), 8 and 32 frames; cfg3: ViT-B/16 with 16 frames
    ("vitl14_t8", dict(input_resolution=224, num_frames=8, patch_size=14, width=1024, layers=2, heads=16), 1),
    ("vitl14_t32", dict(input_resolution=224, num_frames=32, patch_size=14, width=1024, layers=1, heads=16), 1),
    ("vitb16_t16", dict(input_resolution=224, num_frames=16, patch_size=16, width=768, layers=2, heads=12), 1),
])
def test_other_baseline_shapes_vs_oracle(mode, name, kw, batch):
    """Logits and gradients against the CPU oracle (live, same seeded fixture) for the model shapes of cfg3-cfg5."""
    cfg = O.OracleCfg(**kw, block="aim")
    m = _build(cfg, mode)
    x = O.fixture_clip(cfg, batch)
    hw, hb = O.fixture_head(cfg, 400)
    labels = torch.tensor([5] * batch)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, labels)
    ref_loss, ref_lg, ref_g = O.loss_and_grads(O.fixture_state_dict(cfg), x, labels, cfg, hw, hb)
    tol_l, tol_g = (1e-3, 1e-3) if mode == "fp32" else (2e-2, 6e-2)
    assert O.normalised_max_err(lg, ref_lg) < tol_l
    assert int(lg.argmax()) == int(ref_lg.argmax())
    worst = max(O.normalised_max_err(grads[k], ref_g[k]) for k in grads)
    print(f"[{name} {mode}] logits {O.normalised_max_err(lg, ref_lg):.2e}, worst grad {worst:.2e}")
    assert worst < tol_g


def test_inference_three_views_sharded_like_recognizer3d():
    """cfg4-style inference: [videos, views, C, T, H, W] -> views are batch rows (recognizer3d.py:36), softmax-mean over
    views ('prob', recognizers/base.py:186-192); a 2-way shard of the videos gives the same scores as one pass."""
    cfg = O.OracleCfg(**TINY, block="aim")
    m = _build(cfg, "bf16").eval()
    hw, hb = O.fixture_head(cfg, 16)
    g = torch.Generator().manual_seed(4)
    vids = torch.randn(4, 3, 3, 4, 64, 64, generator=g)                  # 4 videos x 3 views

    def scores(v):
        with torch.no_grad():
            feat = m(v.reshape(-1, 3, 4, 64, 64).cuda())
            lg = O.head_logits(feat, hw.cuda(), hb.cuda())
        return torch.softmax(lg, -1).reshape(v.shape[0], 3, -1).mean(1).cpu()

    full = scores(vids)
    shard = torch.cat([scores(vids[0::2]), scores(vids[1::2])])           # rank-strided shards, then gather
    order = torch.cat([torch.arange(0, 4, 2), torch.arange(1, 4, 2)])
    assert torch.allclose(full[order], shard, atol=2e-3)
    ref = O.head_logits(O.backbone(O.fixture_state_dict(cfg), vids.reshape(-1, 3, 4, 64, 64), cfg), hw, hb)
    ref = torch.softmax(ref, -1).reshape(4, 3, -1).mean(1)
    assert O.normalised_max_err(full, ref) < 2e-2


# ---------------------------------------------------------------------------------------------- full-depth configs
# BASELINE.json cfg2 (training batch), cfg4 (ViT-L/14 3-view inference) and cfg5 (ViT-L/14 32 frames) at their FULL
# depth against the live CPU oracle: the bf16 residual stream accumulates rounding with depth, so the 2e-2 / 6e-2 gates
# are checked exactly where they are at risk.  The oracle legs take 10 s - 2 min of host time each (cached per config).
_ORACLE_CACHE = {}


def _oracle_full(name):
    if name in _ORACLE_CACHE:
        return _ORACLE_CACHE[name]
    if name == "cfg4":      # configs/recognition/vit/vitclip_large_k400.py:6 with 8 frames, 1 video x 3 views, eval
        cfg = O.OracleCfg(input_resolution=224, num_frames=8, patch_size=14, width=1024, layers=24, heads=16)
        x = O.fixture_clip(cfg, 3)
        hw, hb = O.fixture_head(cfg, 400)
        with torch.no_grad():
            ref = O.logits(O.fixture_state_dict(cfg), x, cfg, hw, hb)
        out = (cfg, x, hw, hb, ref)
    elif name == "cfg5":    # ViT-L/14, 32 frames, one clip, training step
        cfg = O.OracleCfg(input_resolution=224, num_frames=32, patch_size=14, width=1024, layers=24, heads=16)
        x = O.fixture_clip(cfg, 1)
        hw, hb = O.fixture_head(cfg, 400)
        labels = torch.tensor([5])
        out = (cfg, x, hw, hb, labels) + tuple(O.loss_and_grads(O.fixture_state_dict(cfg), x, labels, cfg, hw, hb))
    else:
        raise KeyError(name)
    _ORACLE_CACHE[name] = out
    return out


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_cfg4_vitl14_24_layers_three_views(mode):
    cfg, x, hw, hb, ref = _oracle_full("cfg4")
    m = _build(cfg, mode).eval()
    with torch.no_grad():
        lg = O.head_logits(m(x.cuda()), hw.cuda(), hb.cuda()).cpu()
    err = O.normalised_max_err(lg, ref)
    prob = torch.softmax(lg, -1).mean(0)                    # average_clips='prob' over the 3 views (recognizers/base.py:186-192)
    prob_ref = torch.softmax(ref, -1).mean(0)
    print(f"[cfg4 full depth {mode}] logits err {err:.3e}, prob err {O.normalised_max_err(prob, prob_ref):.3e}")
    assert err < (1e-3 if mode == "fp32" else 2e-2)
    assert torch.equal(lg.argmax(1), ref.argmax(1)) and int(prob.argmax()) == int(prob_ref.argmax())


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_cfg5_vitl14_24_layers_32_frames_train(mode):
    cfg, x, hw, hb, labels, ref_loss, ref_lg, ref_g = _oracle_full("cfg5")
    m = _build(cfg, mode)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, labels)
    err = O.normalised_max_err(lg, ref_lg)
    worst = max(O.normalised_max_err(grads[k], ref_g[k]) for k in grads)
    print(f"[cfg5 full depth {mode}] logits err {err:.3e}, worst gradient err over {len(grads)} tensors {worst:.3e}")
    assert err < (1e-3 if mode == "fp32" else 2e-2)
    assert int(lg.argmax()) == int(ref_lg.argmax())
    assert worst < (1e-3 if mode == "fp32" else 6e-2)


def test_cfg2_batch8_droppath_all_gradients():
    """The bench configuration itself: ViT-B/16 8x224, 8 clips, drop_path_rate 0.2 (vitclip_base_k400.py:6), training
    mode; the masks drawn by the module are fed to the oracle, all 147 gradients are compared."""
    cfg = O.OracleCfg(block="aim")
    m = _build(cfg, "bf16", drop_path_rate=0.2).train()
    x = O.fixture_clip(cfg, 8, seed=7)
    hw, hb = O.fixture_head(cfg, 400)
    labels = torch.arange(8) * 37 % 400
    torch.manual_seed(11)
    masks = m._drop_masks(m._dims(8), torch.device("cuda"))
    torch.manual_seed(11)                                                   # the forward below draws the same masks
    hwc, hbc = hw.cuda().requires_grad_(True), hb.cuda().requires_grad_(True)
    lg = O.head_logits(m(x.cuda()), hwc, hbc)
    F.cross_entropy(lg, labels.cuda()).backward()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.requires_grad}
    om = [(None, None) if a is None else (a.cpu(), b.cpu()) for a, b in masks]
    assert any(a is not None and float(a.min()) == 0.0 for a, _ in om)
    _, ref_lg, ref_g = O.loss_and_grads(O.fixture_state_dict(cfg), x, labels, cfg, hw, hb, drop_masks=om)
    err = O.normalised_max_err(lg.detach().cpu(), ref_lg)
    worst = max(O.normalised_max_err(grads[k], ref_g[k]) for k in grads)
    print(f"[cfg2 B=8 DropPath 0.2 bf16] logits err {err:.3e}, worst gradient err over {len(grads)} tensors {worst:.3e}")
    assert len(grads) == 147
    assert err < 2e-2 and torch.equal(lg.detach().cpu().argmax(1), ref_lg.argmax(1))
    assert worst < 6e-2


# ---------------------------------------------------------------------------------------------- module-level behaviour
def test_second_inflight_forward_raises_clear_error():
    """One set of saved activations per module: backward of an overwritten forward is a clear AimbError, not a wrong
    gradient (the reference has no such limit; Recognizer3D batches views into one call, recognizer3d.py:16)."""
    from aimb200 import lib
    cfg = O.OracleCfg(**TINY, block="aim")
    m = _build(cfg, "bf16").train()
    x = O.fixture_clip(cfg, 2).cuda()
    f1 = m(x)
    f2 = m(x)
    f2.sum().backward()                                   # the latest forward is fine
    with pytest.raises(lib.AimbError, match="ONE training forward in flight"):
        f1.sum().backward()
    f3 = m(x)                                             # and the module keeps working afterwards
    f3.sum().backward()
    with torch.no_grad():                                 # forwards without grad never disturb a pending one
        f4 = m(x)
        m(x)
    assert torch.isfinite(f4).all()


def test_fp16_input_accepted():
    """auto_fp16 / apex O1 callers hand the backbone half-precision clips (recognizers/base.py:141)."""
    cfg = O.OracleCfg(**TINY, block="aim")
    x = O.fixture_clip(cfg, 2)
    for mode, tol in (("fp32", 2e-3), ("bf16", 2e-2)):
        m = _build(cfg, mode).eval()
        with torch.no_grad():
            a = m(x.half().cuda()).cpu()
            b = m(x.half().float().cuda()).cpu()
        assert O.normalised_max_err(a, b) < (1e-6 if mode == "fp32" else 5e-3)
        ref = O.backbone(O.fixture_state_dict(cfg), x.half().float(), cfg)
        assert O.normalised_max_err(a, ref) < tol


def test_batch_size_change_drops_the_old_buffer_set():
    cfg = O.OracleCfg(**TINY, block="aim")
    m = _build(cfg, "bf16").eval()
    with torch.no_grad():
        m(O.fixture_clip(cfg, 4).cuda())
        n4 = sum(t.numel() * t.element_size() for t in m._engine._bufs.values())
        m(O.fixture_clip(cfg, 2).cuda())
        n2 = sum(t.numel() * t.element_size() for t in m._engine._bufs.values())
        out = m(O.fixture_clip(cfg, 3).cuda())
    assert n2 < n4 and out.shape[0] == 3


def test_distributed_data_parallel_wrapper_sees_ordinary_grads():
    """stock (MM)DistributedDataParallel around the recognizer (apis/train.py:106-110).  World size 1 on
    this box: what is checked is that DDP's autograd hooks, bucket copies and the flat-buffer parameter views coexist."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    cfg = O.OracleCfg(**TINY, block="aim")
    x = O.fixture_clip(cfg, 2).cuda()
    ref = _build(cfg, "bf16").train()
    ref(x).square().mean().backward()
    want = {k: p.grad.clone() for k, p in ref.named_parameters() if p.requires_grad}
    own_pg = not dist.is_initialized()
    if own_pg:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        m = DDP(_build(cfg, "bf16").train(), device_ids=[0], broadcast_buffers=False, find_unused_parameters=False)
        for _ in range(2):                                # second step: the flat-buffer views are already in place
            m.zero_grad(set_to_none=True)
            m(x).square().mean().backward()
        got = {k: p.grad for k, p in m.module.named_parameters() if p.requires_grad}
        assert sorted(got) == sorted(want)
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=1e-4, atol=1e-7), k
    finally:
        if own_pg:
            dist.destroy_process_group()


@pytest.mark.parametrize("block", ["aim", "fork"])
def test_checkpoint_true_recomputes_and_matches(block):
    """checkpoint=True (vit_clip.py:318-319: torch.utils.checkpoint per block): same logits and gradients as with stored
    activations (the recompute is deterministic, DropPath masks are reused), with one shared set of per-block buffers."""
    cfg = O.OracleCfg(**TINY, block=block)
    x = O.fixture_clip(cfg, 2).cuda()
    res = {}
    for ck in (False, True):
        m = aimb200.build_backbone(dict(type="ViT_CLIP", drop_path_rate=0.3, block=block, compute_dtype="bf16", checkpoint=ck, **TINY))
        m.init_weights()
        m.load_state_dict(O.fixture_state_dict(cfg))
        m = m.cuda().train()
        torch.manual_seed(5)
        torch.cuda.reset_peak_memory_stats()
        feat = m(x)
        feat.square().mean().backward()
        nb = sum(t.numel() * t.element_size() for t in m._engine._bufs.values())
        res[ck] = (feat.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.requires_grad}, nb)
    assert torch.equal(res[False][0], res[True][0])
    for k in res[False][1]:
        assert torch.allclose(res[False][1][k], res[True][1][k], rtol=1e-5, atol=1e-8), k    # fp32 atomics: order may differ
    assert res[True][2] < 0.8 * res[False][2]          # 2 layers: one shared activation set instead of two


# ------------------------------------------------------------------------------------------------ ViT_ImageNet variant (f4)
def _build_imagenet(cfg: O.OracleCfg, mode: str, drop_path_rate=0.0):
    from oracle import imagenet_oracle as OI
    m = aimb200.build_backbone(dict(type="ViT_ImageNet", img_size=cfg.input_resolution, num_frames=cfg.num_frames,
                                    patch_size=cfg.patch_size, embed_dim=cfg.width, depth=cfg.layers, num_heads=cfg.heads,
                                    num_tadapter=cfg.num_tadapter, adapter_scale=cfg.adapter_scale,
                                    drop_path_rate=drop_path_rate, compute_dtype=mode))
    m.init_weights()
    m.load_state_dict(OI.fixture_state_dict(cfg))
    return m.cuda()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("nt", [1, 2])
def test_imagenet_tiny_logits_and_all_grads_vs_golden(mode, nt):
    """aimb200.ViT_ImageNet against goldens generated from vit_imagenet.py::ViT_ImageNet itself."""
    gold = np.load(os.path.join(G, "tiny_imagenet" + ("_nt2" if nt == 2 else "") + ".npz"))
    cfg = O.OracleCfg(**TINY, num_tadapter=nt)
    m = _build_imagenet(cfg, mode)
    x = O.fixture_clip(cfg, 2)
    hw, hb = O.fixture_head(cfg, 16)
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, torch.tensor(gold["labels"]))
    tol_l, tol_g = (1e-3, 1e-3) if mode == "fp32" else (2e-2, 6e-2)
    assert O.normalised_max_err(lg, torch.tensor(gold["logits"])) < tol_l
    assert abs(loss - float(gold["loss"])) < tol_l
    n = 0
    for k in gold.files:
        if k.startswith("grad/"):
            n += 1
            assert O.normalised_max_err(grads[k[5:]], torch.tensor(gold[k])) < tol_g, k
    assert n == len(grads)
    m.eval()
    with torch.no_grad():
        feat = m(x.cuda()).cpu()
    assert O.normalised_max_err(feat, torch.tensor(gold["feat"])) < (1e-3 if mode == "fp32" else 2e-2)


def test_imagenet_droppath_is_per_frame_and_matches_oracle():
    from oracle import imagenet_oracle as OI
    cfg = O.OracleCfg(**TINY)
    m = _build_imagenet(cfg, "fp32", drop_path_rate=0.5).train()
    x = O.fixture_clip(cfg, 2)
    torch.manual_seed(3)
    masks = m._drop_masks(m._dims(2), torch.device("cuda"))
    torch.manual_seed(3)
    feat = m(x.cuda())
    n, BT = cfg.tokens, 2 * cfg.num_frames
    a, b = masks[1]
    assert a.numel() == BT * n and torch.equal(a.view(BT, n), a.view(BT, n)[:, :1].expand(BT, n))     # one value per frame
    assert float(a.min()) == 0.0 or float(b.min()) == 0.0, "a dropped frame expected"
    om = [(None, None), (a.view(BT, n)[:, 0].cpu(), b.view(BT, n)[:, 0].cpu())]
    ref = OI.backbone(OI.fixture_state_dict(cfg), x, cfg, drop_masks=om)
    assert O.normalised_max_err(feat.detach().cpu(), ref) < 1e-3


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_imagenet_vitb16_8x224_vs_live_oracle(mode):
    """Full-size timm ViT-B/16 AIM (vit_imagenet_k400.py), one clip: logits and every adapter gradient against the oracle
    (pinned to the reference class by the tiny goldens above)."""
    from oracle import imagenet_oracle as OI
    cfg = O.OracleCfg()
    m = _build_imagenet(cfg, mode)
    x = O.fixture_clip(cfg, 1)
    hw, hb = O.fixture_head(cfg, 400)
    labels = torch.tensor([7])
    lg, loss, grads = _cuda_logits_and_grads(m, cfg, x, hw, hb, labels)
    rl, rlg, rg = OI.loss_and_grads(OI.fixture_state_dict(cfg), x, labels, cfg, hw, hb)
    err = O.normalised_max_err(lg, rlg)
    print(f"[ViT_ImageNet {mode}] logit err {err:.3e}")
    assert err < (1e-3 if mode == "fp32" else 2e-2) and int(lg.argmax()) == int(rlg.argmax())
    tol = 1e-3 if mode == "fp32" else 6e-2
    assert sorted(grads) == sorted(rg)
    for k, g in rg.items():
        assert O.normalised_max_err(grads[k], g) < tol, k

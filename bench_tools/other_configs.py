"""Secondary configurations of BASELINE.json (cfg3-cfg5) at full depth: sanity (finite loss / grads) + throughput.
    python bench_tools/other_configs.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import aimb200  # noqa: E402

dev = torch.device("cuda", 0)


def build(**kw):
    torch.manual_seed(0)
    m = aimb200.build_backbone(dict(type="ViT_CLIP", drop_path_rate=0.2, adapter_scale=0.5, **kw))
    m.init_weights()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "D_fc2" in n or n.endswith("bias") or "temporal_embedding" in n:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return m.to(dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def train_cfg(name, B, **kw):
    m = build(**kw).train()
    D = kw["width"]
    hw = (0.01 * torch.randn(400, D)).to(dev).requires_grad_(True)
    hb = torch.zeros(400, device=dev, requires_grad=True)
    params = [p for p in m.parameters() if p.requires_grad] + [hw, hb]
    opt = torch.optim.AdamW(params, lr=3e-4, fused=True)
    x = torch.randn(B, 3, kw["num_frames"], 224, 224, device=dev)
    y = torch.randint(0, 400, (B,), device=dev)
    state = {}

    def step():
        feat = m(x)
        loss = F.cross_entropy(F.linear(F.dropout(feat.mean((2, 3, 4)), 0.5, True), hw, hb), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        state["loss"] = loss

    ms = timeit(step)
    gn = torch.stack([p.grad.float().norm() for p in params]).norm()
    ok = bool(torch.isfinite(state["loss"])) and bool(torch.isfinite(gn))
    print(f"{name}: B={B} {ms:8.2f} ms/step {B / ms * 1e3:8.1f} clips/s  loss {float(state['loss']):.4f} |grad| {float(gn):.4f} "
          f"finite={ok} peak_mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del m, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def infer_cfg(name, videos, views, **kw):
    m = build(**kw).eval()
    x = torch.randn(videos * views, 3, kw["num_frames"], 224, 224, device=dev)
    hw = (0.01 * torch.randn(400, kw["width"])).to(dev)

    def run():
        with torch.no_grad():
            lg = F.linear(m(x).mean((2, 3, 4)), hw)
            return torch.softmax(lg, -1).view(videos, views, -1).mean(1)

    ms = timeit(run)
    print(f"{name}: {videos} videos x {views} views {ms:8.2f} ms  {videos / ms * 1e3:8.1f} videos/s  finite={bool(torch.isfinite(run()).all())}",
          flush=True)
    del m
    torch.cuda.empty_cache()


VITB = dict(input_resolution=224, patch_size=16, width=768, layers=12, heads=12)
VITL = dict(input_resolution=224, patch_size=14, width=1024, layers=24, heads=16)
train_cfg("cfg3 ViT-B/16 16x224 train (eager)", 8, num_frames=16, **VITB)
infer_cfg("cfg4 ViT-L/14 8x224 3-view inference (eager)", 8, 3, num_frames=8, **VITL)
train_cfg("cfg5 ViT-L/14 32x224 train (eager)", 2, num_frames=32, **VITL)

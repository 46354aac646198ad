"""Spatial attention fwd + bwd at the cfg2 shape (64 frames x 197 tokens x 12 heads), timed with events; few launches
(for ncu).  python bench_tools/attn_one.py [frames] [n] [heads]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 197
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 12
D = heads * 64
dev = "cuda"
qkv = (torch.randn(frames * n, 3 * D, device=dev) * 0.5).bfloat16()
o = torch.empty(frames * n, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(frames * heads * n, device=dev)
d_o = torch.randn(frames * n, D, device=dev).bfloat16()
d_qkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("fwd", lambda: lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)),
                 ("bwd", lambda: lib.attn_spatial_bwd(qkv, o, d_o, lse, d_qkv, frames, n, heads))):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    flops = (4 if name == "fwd" else 10) * frames * heads * n * n * 64
    print(f"attn_spatial_{name} frames={frames} n={n} heads={heads}: {ts[2]:.1f} us  {flops / ts[2] / 1e6:.1f} TFLOP/s (algorithmic)")

# temporal attention at the same shape: B clips x T = 8 frames
B, T = frames // 8, 8
qkv_t = (torch.randn(B * T * n, 3 * D, device=dev) * 0.5).bfloat16()
o_t = torch.empty(B * T * n, D, device=dev, dtype=torch.bfloat16)
d_qkv_t = torch.empty_like(qkv_t)
for name, fn, nbytes in (("fwd", lambda: lib.attn_temporal_fwd(qkv_t, o_t, B, T, n, heads), qkv_t.numel() * 2 + o_t.numel() * 2),
                         ("bwd", lambda: lib.attn_temporal_bwd(qkv_t, d_o, d_qkv_t, B, T, n, heads), 2 * qkv_t.numel() * 2 + o_t.numel() * 2)):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    print(f"attn_temporal_{name} B={B} T={T} n={n} heads={heads}: {ts[2]:.1f} us  {nbytes / ts[2] / 1e3:.0f} GB/s (algorithmic bytes)")

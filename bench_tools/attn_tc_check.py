"""tcgen05 spatial attention against an fp64 torch reference and against the mma.sync kernel (correctness + timing).
python bench_tools/attn_tc_check.py [fwd|bwd|both]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
L = lib.load()
dev = "cuda"


def ref(qkv, frames, n, heads):
    D = heads * 64
    q, k, v = [t.reshape(frames, n, heads, 64).transpose(1, 2) for t in qkv.double().split(D, dim=1)]
    aff = q @ k.transpose(-1, -2) / 8.0
    o = torch.softmax(aff, -1) @ v
    return o.transpose(1, 2).reshape(frames * n, D), torch.logsumexp(aff, -1)


def err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


ok = True
for frames, n, heads in [(1, 16, 1), (2, 17, 2), (1, 33, 1), (2, 64, 1), (1, 128, 2), (3, 197, 2), (2, 129, 1), (2, 256, 1), (5, 197, 12)]:
    g = torch.Generator(device=dev).manual_seed(n)
    D = heads * 64
    qkv = torch.randn(frames * n, 3 * D, device=dev, generator=g).bfloat16()
    qr = qkv.double().requires_grad_(True)
    oref, lref = ref(qr, frames, n, heads)
    if what in ("fwd", "both"):
        o = torch.full((frames * n, D), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.full((frames, heads, n), float("nan"), device=dev)
        L.aimb_debug_attn_mode(0)
        lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)
        torch.cuda.synchronize()
        eo, el = err(o, oref.detach()), err(lse, lref.detach())
        good = eo < 1e-2 and el < 1e-4
        ok &= good
        print(f"fwd frames={frames} n={n} heads={heads}: o err {eo:.2e} lse err {el:.2e} {'OK' if good else 'FAIL'}", flush=True)
    if what in ("bwd", "both"):
        o = torch.empty(frames * n, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(frames, heads, n, device=dev)
        L.aimb_debug_attn_mode(1)
        lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)
        do = torch.randn(frames * n, D, device=dev, generator=g).bfloat16()
        oref.backward(do.double())
        dq = torch.full_like(qkv, float("nan"))
        L.aimb_debug_attn_mode(0)
        lib.attn_spatial_bwd(qkv, o, do, lse, dq, frames, n, heads)
        torch.cuda.synchronize()
        e = [err(dq[:, i * D:(i + 1) * D], qr.grad[:, i * D:(i + 1) * D]) for i in range(3)]
        good = max(e) < 2.5e-2
        ok &= good
        print(f"bwd frames={frames} n={n} heads={heads}: dq {e[0]:.2e} dk {e[1]:.2e} dv {e[2]:.2e} {'OK' if good else 'FAIL'}", flush=True)

# timing at the cfg2 shape, both implementations
frames, n, heads = 64, 197, 12
D = heads * 64
qkv = (torch.randn(frames * n, 3 * D, device=dev) * 0.5).bfloat16()
o = torch.empty(frames * n, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(frames * heads * n, device=dev)
d_o = torch.randn(frames * n, D, device=dev).bfloat16()
d_qkv = torch.empty_like(qkv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode, tag in ((1, "mma.sync"), (0, "tcgen05")):
    L.aimb_debug_attn_mode(mode)
    for name, fn in (("fwd", lambda: lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)),
                     ("bwd", lambda: lib.attn_spatial_bwd(qkv, o, d_o, lse, d_qkv, frames, n, heads))):
        if name not in (what, "fwd" if what == "both" else what, "bwd" if what == "both" else what):
            continue
        for _ in range(2):
            fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        flops = (4 if name == "fwd" else 10) * frames * heads * n * n * 64
        print(f"{tag} {name} 64x197x12: median {ts[3]:.1f} us (min {ts[0]:.1f})  {flops / ts[3] / 1e6:.0f} TFLOP/s algorithmic", flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)

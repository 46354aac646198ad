"""Per-op device-time breakdown of one training step (CUDA events around every C-ABI call).
    python bench_tools/breakdown.py [--steps 3] [--ncu]   (with --ncu: 1 step between cudaProfilerStart/Stop)"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from aimb200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--ncu", action="store_true")
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()
dev = torch.device("cuda", 0)
tr = bench.Trainer(dev, 1, dtype=args.dtype)
g = torch.Generator().manual_seed(2)
x = torch.randn(8, 3, 8, 224, 224, generator=g).to(dev)
y = torch.randint(0, 400, (8,), generator=g).to(dev)
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
if args.ncu:
    torch.cuda.profiler.start()
    tr.step(x, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

recs = []
names = ["layernorm_fwd", "layernorm_bwd", "im2col", "stem_assemble_ln", "temb_grad", "tail_fwd", "tail_bwd", "gemm_nt",
         "gemm_wgrad", "adapter_fused", "colsum", "transpose", "attn_spatial_fwd", "attn_spatial_bwd", "attn_temporal_fwd", "attn_temporal_bwd"]
for nme in names:
    orig = getattr(lib, nme)

    def mk(orig, nme):
        def f(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **k)
            e.record()
            tag = nme
            if nme == "gemm_nt":
                tag += f" M{a[0].shape[0]} N{a[1].shape[0]} K{a[0].shape[1]}"
                fl = 2.0 * a[0].shape[0] * a[1].shape[0] * a[0].shape[1]
            elif nme == "adapter_fused":
                tag += f" M{a[0].shape[0]} D{a[0].shape[1]} R{a[1].shape[0]} " + ("bwd" if "dact_src" in a[5] else "fwd")
                fl = 4.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[0]
            elif nme == "gemm_wgrad":
                tag += f" R{a[0].shape[0]} N{a[0].shape[1]} K{a[1].shape[1]}"
                fl = 2.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1]
            else:
                fl = 0.0
            recs.append((tag, s, e, fl))
            return r
        return f
    setattr(lib, nme, mk(orig, nme))

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(args.steps):
    tr.step(x, y)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / args.steps
agg = collections.OrderedDict()
for tag, s, e, fl in recs:
    a = agg.setdefault(tag, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += s.elapsed_time(e)
    a[2] += fl
print(f"step {tot:.3f} ms (instrumented, eager)")
acc = 0.0
for tag, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    ms /= args.steps
    acc += ms
    tf = (fl / args.steps) / (ms / 1e3) / 1e12 if fl else 0
    print(f"{tag:40s} calls/step {n // args.steps:4d}  {ms:9.3f} ms  {100 * ms / tot:5.1f}%  {tf:8.1f} TFLOP/s")
print(f"sum of C-ABI calls {acc:.3f} ms = {100 * acc / tot:.1f}% of step")

"""wgrad_tc_kernel at the step's two shapes (dW2 = dy^T g: [768, 192]; dW1 = d_h^T a: [192, 768]), M = 12608.
Timed as a CUDA graph of 12 launches into the same gradient buffer (L2 flushed before)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

lib.load()
M = 12608
fl = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for N, K in ((768, 192), (192, 768)):
    dy = torch.randn(M, N, device="cuda").bfloat16()
    x = torch.randn(M, K, device="cuda").bfloat16()
    dw = torch.zeros(N, K, device="cuda")
    for _ in range(3):
        lib.gemm_wgrad(dy, x, dw, accumulate=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(12):
            lib.gemm_wgrad(dy, x, dw, accumulate=True)
    ts = []
    for _ in range(5):
        fl.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 12)
    print(f"wgrad dW[{N},{K}] over M={M}: {min(ts):.2f} us per launch ({2.0 * M * N * K / min(ts) / 1e6:.0f} TFLOP/s, "
          f"{(M * (N + K) * 2) / min(ts) / 1e6:.2f} TB/s of operand bytes)")

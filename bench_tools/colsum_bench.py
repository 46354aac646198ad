"""colsum_kernel: rows per block (one wave vs the fixed 128 of rounds 1-2a) at the two shapes of the step.
Timed as a CUDA graph of 12 launches (L2 flushed before)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
L.aimb_debug_colsum_rpb.argtypes = [__import__("ctypes").c_int]
L.aimb_debug_colsum_rpb.restype = None
fl = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for C in (768, 192):
    x = torch.randn(12608, C, device="cuda").bfloat16()
    o = torch.zeros(C, device="cuda")
    for rpb in (128, 0, 96, 160):
        L.aimb_debug_colsum_rpb(rpb)
        for _ in range(3):
            lib.colsum(x, o, accumulate=True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(12):
                lib.colsum(x, o, accumulate=True)
        ts = []
        for _ in range(5):
            fl.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 12)
        print(f"C={C} rows_per_block={'auto' if rpb == 0 else rpb}: {min(ts):.2f} us per launch ({12608 * C * 2 / min(ts) / 1e6:.2f} TB/s)")
L.aimb_debug_colsum_rpb(0)

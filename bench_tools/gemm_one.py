"""One GEMM config, few launches (for ncu).  python bench_tools/gemm_one.py M N K bn mode"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
M, N, K, bn, mode = [int(v) for v in sys.argv[1:6]]
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / 30).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda").bfloat16()
L.aimb_debug_force_bn(bn)
L.aimb_debug_cta_mode(mode)
res = torch.randn(M, N, device="cuda").bfloat16() if len(sys.argv) > 6 and sys.argv[6] == "res" else None
for _ in range(5):
    lib.gemm_nt(a, w, out, bias=bias, res1=res)
torch.cuda.synchronize()

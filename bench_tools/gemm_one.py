"""A few GEMM configs, few launches each (for ncu).  python bench_tools/gemm_one.py mode N:K:kind[:bn] [N:K:kind[:bn] ...]
kinds as in gemm_epi.py; every config is launched 3 times (profile with --launch-skip / -c to taste)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
M = 12608
mode = int(sys.argv[1])
for spec in sys.argv[2:]:
    parts = spec.split(":")
    N, K, kind = int(parts[0]), int(parts[1]), parts[2]
    bn = int(parts[3]) if len(parts) > 3 else 0
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / 30).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    r1 = torch.randn(M, N, device="cuda").bfloat16()
    r2 = torch.randn(M, N, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda").bfloat16()
    cs = torch.zeros(N, device="cuda")
    kw = {"bias": dict(bias=bias), "none": dict(),
          "qgelu+pre": dict(bias=bias, act=lib.ACT_QUICKGELU, out_pre=pre),
          "gelu+pre": dict(bias=bias, act=lib.ACT_GELU, out_pre=pre),
          "dqgelu": dict(dact_src=r1, dact=lib.ACT_QUICKGELU),
          "dgelu+cs": dict(dact_src=r1, dact=lib.ACT_GELU, colsum_out=cs),
          "res1": dict(bias=bias, res1=r1), "res1+res2": dict(bias=bias, res1=r1, res2=r2, alpha=0.5)}[kind]
    L.aimb_debug_force_bn(bn)
    L.aimb_debug_cta_mode(mode)
    for _ in range(3):
        lib.gemm_nt(a, w, out, **kw)
    torch.cuda.synchronize()

"""GEMM timing per model epilogue (the exact epilogues engine.py issues).  python bench_tools/gemm_epi.py [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
import ctypes  # noqa: E402
L.aimb_debug_skip_epilogue.argtypes = [ctypes.c_int]
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = 12608


def t(N, K, kind, mode, bn=0, iters=12, do_flush=True, skip=0):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / 30).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    r1 = torch.randn(M, N, device=dev).bfloat16()
    r2 = torch.randn(M, N, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    cs = torch.zeros(N, device=dev)
    kw = {"bias": dict(bias=bias), "none": dict(),
          "qgelu+pre": dict(bias=bias, act=lib.ACT_QUICKGELU, out_pre=pre),
          "qgelu+alias": dict(bias=bias, act=lib.ACT_QUICKGELU, out_pre=out),
          "gelu+pre": dict(bias=bias, act=lib.ACT_GELU, out_pre=pre),
          "dqgelu": dict(dact_src=r1, dact=lib.ACT_QUICKGELU),
          "dgelu+cs": dict(dact_src=r1, dact=lib.ACT_GELU, colsum_out=cs),
          "res1": dict(bias=bias, res1=r1), "res1+res2": dict(bias=bias, res1=r1, res2=r2, alpha=0.5)}[kind]
    L.aimb_debug_force_bn(bn)
    L.aimb_debug_cta_mode(mode)
    L.aimb_debug_skip_epilogue(skip)
    for _ in range(3):
        lib.gemm_nt(a, w, out, **kw)
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lib.gemm_nt(a, w, out, **kw)
        lib.gemm_nt(a, w, out, **kw)
        lib.gemm_nt(a, w, out, **kw)
        lib.gemm_nt(a, w, out, **kw)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / 4)
    ts.sort()
    L.aimb_debug_force_bn(0)
    L.aimb_debug_cta_mode(0)
    L.aimb_debug_skip_epilogue(0)
    us = ts[len(ts) // 2] * 1e3
    return us, 2.0 * M * N * K / us / 1e6


modes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 3]
skips = [int(x) for x in os.environ.get('SKIPS', '0').split(',')]
if os.environ.get('DIRECT'):   # 0 never / 1 auto (K >= 512) / 2 always: the register-store epilogue
    L.aimb_debug_direct_epilogue.argtypes = [ctypes.c_int]
    L.aimb_debug_direct_epilogue(int(os.environ['DIRECT']))
bns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
cases = [(2304, 768, "bias"), (768, 768, "bias"), (768, 768, "res1"), (768, 768, "none"), (3072, 768, "qgelu+pre"), (3072, 768, "dqgelu"),
         (768, 3072, "res1"), (768, 3072, "none"), (768, 2304, "none"), (192, 768, "gelu+pre"), (192, 768, "dgelu+cs"), (768, 192, "res1"),
         (768, 192, "res1+res2"), (768, 192, "none")]
if len(sys.argv) > 3 and sys.argv[3] == "epi":     # K = 64: mainloop negligible -> time ~ waves x epilogue time per tile
    cases = [(3072, 64, k) for k in ("none", "bias", "qgelu+pre", "dqgelu", "res1", "res1+res2")] + [(768, 64, k) for k in ("none", "res1", "res1+res2")]
if len(sys.argv) > 3 and sys.argv[3] == "st":
    cases = [(3072, 768, k) for k in ("bias", "qgelu+pre", "qgelu+alias", "dqgelu")]
if len(sys.argv) > 3 and sys.argv[3] == "ad":
    cases = [(192, 768, "gelu+pre"), (192, 768, "dgelu+cs"), (768, 192, "res1"), (768, 192, "res1+res2"), (768, 192, "none"), (768, 768, "res1"), (768, 768, "bias")]
if len(sys.argv) > 3 and sys.argv[3] == "big":
    cases = [(3072, 768, k) for k in ("bias", "qgelu+pre", "dqgelu")] + [(768, 3072, "res1"), (768, 768, "res1"), (768, 192, "res1"), (768, 192, "res1+res2"), (192, 768, "gelu+pre")]
for (N, K, kind) in cases:
    for mode in modes:
        for bn in bns:
            if bn and N % bn:
                continue
            for skip in skips:
                us, tf = t(N, K, kind, mode, bn, skip=skip)
                print(f"N{N:5d} K{K:5d} {kind:10s} mode{mode} bn{bn:3d} skip{skip} {us:8.1f} us {tf:7.1f} TFLOP/s", flush=True)

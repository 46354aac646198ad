"""GEMM kernel sweep: tile width, 1-CTA vs CTA-pair, with / without epilogue.  python bench_tools/gemm_sweep.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
L.aimb_debug_skip_epilogue.argtypes = [ctypes.c_int]
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(M, N, K, bn, mode, skip, iters=10, **kw):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / 30).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev).bfloat16()
    L.aimb_debug_force_bn(bn)
    L.aimb_debug_cta_mode(mode)
    L.aimb_debug_skip_epilogue(skip)
    for _ in range(3):
        lib.gemm_nt(a, w, out, bias=bias, **kw)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lib.gemm_nt(a, w, out, bias=bias, **kw)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    t = ts[len(ts) // 2]
    L.aimb_debug_force_bn(0)
    L.aimb_debug_cta_mode(0)
    L.aimb_debug_skip_epilogue(0)
    return t * 1e3, 2.0 * M * N * K / (t * 1e-3) / 1e12


shapes = [(12608, 2304, 768), (12608, 768, 3072), (12608, 768, 768), (12608, 3072, 768), (12608, 768, 192), (12608, 192, 768)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for (M, N, K) in shapes:
    for mode in (1, 3):
        for bn in (128, 192, 256):
            if N % bn or (mode == 2 and bn < 192):
                continue
            for skip in (0,):
                us, tf = bench(M, N, K, bn, mode, skip)
                print(f"M{M} N{N} K{K} { {1:'1cta',2:'2cta',3:'wide'}[mode] } BN{bn} {['epi  ','ldtm ','trans'][skip]} {us:8.1f} us {tf:7.1f} TFLOP/s", flush=True)
# reference point: cuBLAS through torch for the same shape (library call, context only)
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    for _ in range(3):
        torch.matmul(a, w.t())
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, w.t())
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    print(f"cuBLAS M{M} N{N} K{K} {ts[5]*1e3:8.1f} us {2.0*M*N*K/(ts[5]*1e-3)/1e12:7.1f} TFLOP/s")

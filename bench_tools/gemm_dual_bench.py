"""Paired GEMM (aimb_gemm_dual) against the separate launches it replaces, at the cfg2 MLP shapes (M = 12608, D = 768, r = 192).
  python bench_tools/gemm_dual_bench.py [M D R]
Each candidate is captured into a CUDA graph of REP back-to-back launches (L2 flushed before), timed with CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

lib.load()
M, D, R = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (12608, 768, 192)
bf = torch.bfloat16
REP = 12
rnd = lambda *s: torch.randn(*s, device="cuda").to(bf)
a = rnd(M, D)
wfc, w1 = rnd(4 * D, D) / 28, rnd(R, D) / 28
bfc, b1 = rnd(4 * D), rnd(R)
wp, w2 = rnd(D, 4 * D) / 55, rnd(D, R) / 14
bp, b2 = rnd(D), rnd(D)
x2 = rnd(M, D)
rs = (torch.rand(197, device="cuda") > 0.2).float() / 0.8
gf, hf = torch.empty(M, 4 * D, device="cuda", dtype=bf), torch.empty(M, 4 * D, device="cuda", dtype=bf)
g, h = torch.empty(M, R, device="cuda", dtype=bf), torch.empty(M, R, device="cuda", dtype=bf)
tmp, xo = torch.empty(M, D, device="cuda", dtype=bf), torch.empty(M, D, device="cuda", dtype=bf)
d_hf, d_h, d_xn = torch.empty_like(gf), torch.empty_like(g), torch.empty_like(xo)
cs = torch.zeros(R, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(REP):
            fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / REP)
    print(f"{name:58s} {min(ts):8.2f} us")
    return min(ts)


print(f"M={M} D={D} R={R}")
# ---- forward, N-concatenation
t1 = timeit("c_fc (QuickGELU + pre)", lambda: lib.gemm_nt(a, wfc, gf, bias=bfc, act=lib.ACT_QUICKGELU, out_pre=hf))
t2 = timeit("D_fc1 (GELU + pre, row_scale)", lambda: lib.gemm_nt(a, w1, g, bias=b1, act=lib.ACT_GELU, out_pre=h, row_scale=rs))
t3 = timeit("paired [c_fc | D_fc1]", lambda: lib.gemm_dual_ncat(
    a, wfc, w1, gf, g, dict(bias=bfc, act=lib.ACT_QUICKGELU, out_pre=hf),
    dict(bias=b1, act=lib.ACT_GELU, out_pre=h, row_scale=rs, alpha=0.5)))
print(f"  -> separate {t1 + t2:.2f} us, paired {t3:.2f} us")
# ---- forward, K-concatenation
t1 = timeit("D_fc2 (+ x2, row-scaled bias)", lambda: lib.gemm_nt(g, w2, tmp, bias=b2, row_scale=rs, bias_rowscaled=True, alpha=0.5, res1=x2))
t2 = timeit("c_proj (+ res1)", lambda: lib.gemm_nt(gf, wp, xo, bias=bp, res1=tmp))
t3 = timeit("paired c_proj + D_fc2", lambda: lib.gemm_dual_kcat(gf, wp, g, w2, xo, bias2=b2, bias2_row_scale=rs, bias2_scale=0.5,
                                                               bias=bp, res1=x2))
print(f"  -> separate {t1 + t2:.2f} us, paired {t3:.2f} us")
# ---- backward, N-concatenation (dx = xo)
wpT, w2T = wp.t().contiguous(), w2.t().contiguous()
t1 = timeit("d_hf = dx Wp . QuickGELU'(hf)", lambda: lib.gemm_nt(xo, wpT, d_hf, dact_src=hf, dact=lib.ACT_QUICKGELU))
t2 = timeit("d_h = dx W2 . GELU'(h) (+ db1 column sums)", lambda: lib.gemm_nt(xo, w2T, d_h, dact_src=h, dact=lib.ACT_GELU, alpha=0.5, row_scale=rs,
                                                                             colsum_out=cs, colsum_accumulate=True))
t3 = timeit("paired [d_hf | d_h]", lambda: lib.gemm_dual_ncat(xo, wpT, w2T, d_hf, d_h, dict(dact_src=hf, dact=lib.ACT_QUICKGELU),
                                                             dict(dact_src=h, dact=lib.ACT_GELU, alpha=0.5, row_scale=rs)))
t4 = timeit("colsum(d_h)", lambda: lib.colsum(d_h, cs, accumulate=True))
print(f"  -> separate {t1 + t2:.2f} us, paired {t3:.2f} us (+ {t4:.2f} us column sums on the side stream)")
# ---- backward, K-concatenation
wfcT, w1T = wfc.t().contiguous(), w1.t().contiguous()
t1 = timeit("d_xn = d_hf Wfc", lambda: lib.gemm_nt(d_hf, wfcT, d_xn))
t2 = timeit("d_xn += d_h W1", lambda: lib.gemm_nt(d_h, w1T, d_xn, res1=d_xn))
t3 = timeit("paired d_hf Wfc + d_h W1", lambda: lib.gemm_dual_kcat(d_hf, wfcT, d_h, w1T, d_xn))
print(f"  -> separate {t1 + t2:.2f} us, paired {t3:.2f} us")

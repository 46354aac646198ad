"""The QKV GEMM with the LayerNorm folded in (variant 8 of gemm_tc4_kernel) at the cfg2 shape, a few launches (for ncu):
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc4 -c 6 -o gpurun_out/r2_qkv_lnfold python bench_tools/qkv_lnfold_one.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

lib.load()
M, D = 12608, 768
x = (torch.randn(M, D, device="cuda") * 2 + 0.5).bfloat16()
w = (torch.randn(3 * D, D, device="cuda") / 28).bfloat16()
b = torch.randn(3 * D, device="cuda").bfloat16()
g, be = torch.ones(D, device="cuda").bfloat16(), torch.zeros(D, device="cuda").bfloat16()
mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
ws = w.float().sum(1).contiguous()
out = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    lib.layernorm_fwd(x, g, be, None, mean, rstd)                      # statistics only
    lib.gemm_nt(x, w, out, bias=b, ln_mean=mean, ln_rstd=rstd, ln_wsum=ws)
    lib.gemm_nt(x, w, out, bias=b)                                      # the plain variant for comparison
torch.cuda.synchronize()
ref = torch.nn.functional.layer_norm(x.float(), (D,)) @ w.float().T + b.float()
lib.gemm_nt(x, w, out, bias=b, ln_mean=mean, ln_rstd=rstd, ln_wsum=ws)
torch.cuda.synchronize()
print("max err vs fp32 LayerNorm + matmul:", float((out.float() - ref).abs().max() / ref.abs().max()))

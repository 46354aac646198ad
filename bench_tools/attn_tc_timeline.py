"""Event timeline of CTA 0 of the tcgen05 spatial attention forward (clock64 stamps per unit).
python bench_tools/attn_tc_timeline.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aimb200 import lib  # noqa: E402

L = lib.load()
L.aimb_debug_attn_timeline.argtypes = [ctypes.c_void_p]
L.aimb_debug_attn_timeline.restype = None
frames, n, heads = 64, 197, 12
D = heads * 64
qkv = (torch.randn(frames * n, 3 * D, device="cuda") * 0.5).bfloat16()
o = torch.empty(frames * n, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(frames * heads * n, device="cuda")
for _ in range(2):
    lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)
tl = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.aimb_debug_attn_timeline(tl.data_ptr())
lib.attn_spatial_fwd(qkv, o, lse, frames, n, heads)
torch.cuda.synchronize()
L.aimb_debug_attn_timeline(None)
t = tl.cpu().view(64, 16)
t0 = int(t[t > 0].min())
names = ["s_full seen", "pass1 done", "p_full arrive", "o_full seen", "O loaded", "stored", "", "", "mma: p_full seen", "mma: PV issued", "mma: S issued"]
print("== forward")
print("unit wg | " + " | ".join(f"{x:>16s}" for x in names if x))
for u in range(64):
    if int(t[u].max()) == 0:
        continue
    print(f"{u:4d} {u & 1:2d} | " + " | ".join(f"{(int(t[u, e]) - t0) if int(t[u, e]) else -1:16d}" for e, x in enumerate(names) if x))

# ---- backward: one row per step (key block x query half) of CTA 0
d_o = torch.randn(frames * n, D, device="cuda").bfloat16()
d_qkv = torch.empty_like(qkv)
for _ in range(2):
    lib.attn_spatial_bwd(qkv, o, d_o, lse, d_qkv, frames, n, heads)
tl.zero_()
L.aimb_debug_attn_timeline(tl.data_ptr())
lib.attn_spatial_bwd(qkv, o, d_o, lse, d_qkv, frames, n, heads)
torch.cuda.synchronize()
L.aimb_debug_attn_timeline(None)
t = tl.cpu().view(64, 16)
t0 = int(t[t > 0].min())
names = ["sdp_full seen", "math done", "pds arrive", "dvk_full seen", "dv/dk stored", "dq_full seen", "acc loaded", "staged+bar", "mma: S,dP issued", "mma: pds seen", "mma: dV,dK,dQ iss"]
print("== backward")
print("step | " + " | ".join(f"{x:>17s}" for x in names if x))
for u in range(64):
    if int(t[u].max()) == 0:
        continue
    print(f"{u:4d} | " + " | ".join(f"{(int(t[u, e]) - t0) if int(t[u, e]) else -1:17d}" for e, x in enumerate(names) if x))

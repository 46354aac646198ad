"""ctypes binding of ``libaimb200.so`` (C ABI declared in ``include/aimb200.h``).

PyTorch is only the plumbing here (device memory, streams): every wrapper takes CUDA tensors,
checks them, and enqueues ONE library call on the current torch stream.  There is no CPU or
eager-PyTorch fallback: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaimb200.so")

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_QUICKGELU, ACT_GELU = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_MMA = 0, 1, 2

_ERR = {-1: "AIMB_ERR_ARG (bad shape/alignment/null)", -2: "AIMB_ERR_CUDA", -3: "AIMB_ERR_UNSUPPORTED",
        -4: "AIMB_ERR_DRIVER (cuTensorMapEncodeTiled)"}


class AimbError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("row_scale", C.c_void_p), ("res1", C.c_void_p), ("res2", C.c_void_p),
                ("dact_src", C.c_void_p), ("out", C.c_void_p), ("out_pre", C.c_void_p), ("alpha", C.c_float),
                ("row_mod", C.c_int32), ("act", C.c_int32), ("dact", C.c_int32), ("bias_rowscaled", C.c_int32),
                ("out_f32", C.c_int32), ("accumulate", C.c_int32), ("ldo", C.c_int64), ("colsum_out", C.c_void_p),
                ("colsum_accumulate", C.c_int32), ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p), ("ln_wsum", C.c_void_p)]


# name -> argtypes; every function returns int except where noted.  Must list EVERY symbol of aimb200.h.
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "aimb_version": [],
    "aimb_device_ok": [_I],
    "aimb_im2col": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "aimb_stem_assemble_ln": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "aimb_temb_grad": [_P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _L, _I, _F, _I, _P],
    "aimb_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "aimb_layernorm_bwd_colsum": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _P, _L, _I, _I, _P],
    "aimb_tail_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "aimb_tail_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_gemm_nt": [_P, _L, _P, _L, C.POINTER(Epilogue), _L, _I, _I, _I, _I, _P],
    "aimb_gemm_strided": [_P, _L, _L, _P, _L, _L, C.POINTER(Epilogue), _L, _I, _I, _I, _P],
    "aimb_adapter_fused": [_P, _L, _P, _P, C.POINTER(Epilogue), C.POINTER(Epilogue), _L, _I, _I, _I, _P],
    "aimb_gemm_dual": [_I, _P, _L, _P, _L, _P, _L, _P, _L, C.POINTER(Epilogue), C.POINTER(Epilogue), _P, _P, _I, _F, _L, _I, _I,
                       _I, _I, _I, _P],
    "aimb_gemm_wgrad": [_P, _L, _P, _L, _P, _L, _I, _I, _F, _I, _I, _I, _P],
    "aimb_colsum": [_P, _L, _P, _I, _F, _P, _L, _I, _I, _I, _P],
    "aimb_transpose": [_P, _P, _I, _I, _I, _P],
    "aimb_adamw_flat": [_P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _L, _P],
    "aimb_transpose_batched": [_P, _P, _P, _I, _I, _P],
    "aimb_attn_spatial_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_attn_spatial_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_attn_temporal_fwd": [_P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_attn_temporal_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "aimb_fork_weights": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "aimb_fork_combine": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "aimb_fork_combine_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
}

_lib = None
launches = 0  # number of library kernels-launching calls made through this module (bench reports it)


def load():
    """Load the shared library (once).  Raises AimbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise AimbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.aimb_last_error.restype = C.c_char_p
    lib.aimb_last_error.argtypes = []
    lib.aimb_debug_force_bn.argtypes = [C.c_int]
    lib.aimb_debug_force_bn.restype = None
    lib.aimb_debug_cta_mode.argtypes = [C.c_int]
    lib.aimb_debug_cta_mode.restype = None
    lib.aimb_debug_set_pdl.argtypes = [C.c_int]
    lib.aimb_debug_set_pdl.restype = None
    lib.aimb_debug_attn_mode.argtypes = [C.c_int]
    lib.aimb_debug_attn_mode.restype = None
    if os.environ.get("AIMB200_ATTN", "tc") == "mma":       # A/B switch: legacy mma.sync spatial attention
        lib.aimb_debug_attn_mode(1)
    # programmatic dependent launch: every kernel implements the protocol (griddepcontrol.launch_dependents at entry,
    # griddepcontrol.wait before the first global access), so the next kernel's CTAs are scheduled and run their prologue
    # (barrier init, TMEM allocation, descriptor prefetch) while the previous grid drains.  Neutral at 17 ms/step in round 1;
    # at 10.4 ms/step with 440 launches it is worth 3 % (771 -> 795 clips/s, same box A/B) -> default on
    lib.aimb_debug_set_pdl(1 if os.environ.get("AIMB200_PDL", "1") == "1" else 0)
    if os.environ.get("AIMB200_GEMM_MODE"):
        lib.aimb_debug_cta_mode(int(os.environ["AIMB200_GEMM_MODE"]))
    _lib = lib
    return lib


def _chk(rc: int, what: str):
    if rc != 0:
        raise AimbError(f"{what} failed: {_ERR.get(rc, rc)}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dt_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.uint8:
        return U8
    raise AimbError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise AimbError("aimb200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_contiguous():
        raise AimbError("tensor must be contiguous")
    return t


def _count():
    global launches
    launches += 1


# ------------------------------------------------------------------------------------------------ wrappers
def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-5):
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    _count()
    _chk(load().aimb_layernorm_fwd(_ptr(_c(x)), _ptr(gamma), _ptr(beta), _ptr(_c(y)) if y is not None else None, _ptr(mean),
                                   _ptr(rstd), rows, D,
                                   eps, dt_code(x), _stream()), "layernorm_fwd")
    return y


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, colsum_out=None, colsum_row_scale=None, colsum_alpha=1.0):
    """dx = dres + LN'(dy); optionally colsum_out[c] = colsum_alpha * sum_r dx[r,c] * row_scale[r % len] (fused)."""
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    _count()
    if colsum_out is None:
        _chk(load().aimb_layernorm_bwd(_ptr(_c(dy)), _ptr(_c(x)), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres),
                                       _ptr(_c(dx)), rows, D, dt_code(x), _stream()), "layernorm_bwd")
    else:
        assert colsum_out.dtype == torch.float32 and colsum_out.numel() == D
        rs = colsum_row_scale
        _chk(load().aimb_layernorm_bwd_colsum(_ptr(_c(dy)), _ptr(_c(x)), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres),
                                              _ptr(_c(dx)), _ptr(rs), rs.numel() if rs is not None else 0, colsum_alpha,
                                              _ptr(colsum_out), rows, D, dt_code(x), _stream()), "layernorm_bwd_colsum")
    return dx


def im2col(x, cols, patch, mean=None, std=None):
    B, Cc, T, H, W = x.shape
    assert Cc == 3
    _count()
    _chk(load().aimb_im2col(_ptr(_c(x)), dt_code(x), _ptr(mean), _ptr(std), _ptr(_c(cols)), dt_code(cols), B, T, H, W,
                            patch, cols.shape[-1], _stream()), "im2col")
    return cols


def stem_assemble_ln(tok, cls, pos, temb, gamma, beta, z, x, mean, rstd, B, T, n, eps=1e-5):
    D = tok.shape[-1]
    _count()
    _chk(load().aimb_stem_assemble_ln(_ptr(_c(tok)), _ptr(cls), _ptr(pos), _ptr(temb), _ptr(gamma), _ptr(beta), _ptr(z),
                                      _ptr(_c(x)), _ptr(mean), _ptr(rstd), B, T, n, D, eps, dt_code(tok), _stream()),
         "stem_assemble_ln")


def temb_grad(dz, out, B, T, n):
    _count()
    _chk(load().aimb_temb_grad(_ptr(_c(dz)), _ptr(out), B, T, n, dz.shape[-1], dt_code(dz), _stream()), "temb_grad")


def tail_fwd(x, gamma, beta, feat, mean, rstd, B, T, n, eps=1e-5):
    _count()
    _chk(load().aimb_tail_fwd(_ptr(_c(x)), _ptr(gamma), _ptr(beta), _ptr(feat), _ptr(mean), _ptr(rstd), B, T, n,
                              x.shape[-1], eps, dt_code(x), _stream()), "tail_fwd")


def tail_bwd(dfeat, x, mean, rstd, gamma, dx, dgamma, dbeta, B, T, n):
    assert dfeat.dtype == torch.float32 and dgamma.dtype == torch.float32
    _count()
    _chk(load().aimb_tail_bwd(_ptr(_c(dfeat)), _ptr(_c(x)), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(_c(dx)),
                              _ptr(dgamma), _ptr(dbeta), B, T, n, x.shape[-1], dt_code(x), _stream()), "tail_bwd")


def make_epilogue(out, bias=None, row_scale=None, res1=None, res2=None, dact_src=None, out_pre=None, alpha=1.0, act=0,
                  dact=0, bias_rowscaled=False, out_f32=False, accumulate=False, ldo=0, colsum_out=None,
                  colsum_accumulate=False, ln_mean=None, ln_rstd=None, ln_wsum=None) -> Epilogue:
    e = Epilogue()
    e.bias, e.row_scale, e.res1, e.res2 = _ptr(bias), _ptr(row_scale), _ptr(res1), _ptr(res2)
    e.dact_src, e.out, e.out_pre = _ptr(dact_src), _ptr(out), _ptr(out_pre)
    e.alpha = alpha
    e.row_mod = row_scale.numel() if row_scale is not None else 1
    e.act, e.dact = act, dact
    e.bias_rowscaled, e.out_f32, e.accumulate = int(bias_rowscaled), int(out_f32), int(accumulate)
    e.ldo = ldo
    e.colsum_out = _ptr(colsum_out)
    e.colsum_accumulate = int(colsum_accumulate)
    if ln_mean is not None:      # LayerNorm folded into this GEMM: fp32 row statistics of A + row sums of the scaled weight
        assert ln_rstd is not None and ln_wsum is not None and ln_wsum.dtype == torch.float32 and ln_mean.dtype == torch.float32
    e.ln_mean, e.ln_rstd, e.ln_wsum = _ptr(ln_mean), _ptr(ln_rstd), _ptr(ln_wsum)
    return e


def gemm_nt(a, w, out, impl=IMPL_AUTO, **epi):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T); a / w may be row-strided views (last dim contiguous)."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1 and out.shape[0] == M and out.shape[1] == N
    e = make_epilogue(out, ldo=out.stride(0), **epi)
    _count()
    _chk(load().aimb_gemm_nt(_ptr(a), a.stride(0), _ptr(w), w.stride(0), C.byref(e), M, N, K, dt_code(a), impl,
                             _stream()), f"gemm_nt[{M}x{N}x{K}]")
    return out


def adapter_fused_supported(a, R, D) -> bool:
    return (a.dtype == torch.bfloat16 and a.shape[0] >= 128 and R in (64, 192, 256) and (D % 192 == 0 or D % 256 == 0)
            and not (R == 64 and D % 256 != 0) and not (R == 192 and D % 192 != 0) and not (R == 256 and D % 256 != 0)
            and a.stride(1) == 1 and a.stride(0) % 8 == 0)


def adapter_fused(a, w1, w2, hidden_out, out, epi1: dict, epi2: dict):
    """out = epi2(epi1(a @ w1^T) @ w2^T); a [M,D], w1 [R,D], w2 [D,R]; hidden_out [M,R] receives epi1's result."""
    M, D = a.shape
    R = w1.shape[0]
    assert w1.shape == (R, D) and w2.shape == (D, R) and w1.is_contiguous() and w2.is_contiguous()
    e1 = make_epilogue(hidden_out, ldo=hidden_out.stride(0), **epi1)
    e2 = make_epilogue(out, ldo=out.stride(0), **epi2)
    _count()
    _chk(load().aimb_adapter_fused(_ptr(a), a.stride(0), _ptr(w1), _ptr(w2), C.byref(e1), C.byref(e2), M, D, R, dt_code(a),
                                   _stream()), f"adapter_fused[{M}x{D}x{R}]")
    return out


def dual_supported_dims(M: int, K: int, dtype, n1: int, n2: int, k2: int = 0) -> bool:
    """shapes the paired tcgen05 GEMM (aimb_gemm_dual) tiles; anything else takes the separate gemm_nt calls"""
    if dtype != torch.bfloat16 or M < 128 or K % 64:
        return False
    if k2:      # KCAT: one [M, n1] output over K + k2
        return k2 % 64 == 0 and (n1 % 192 == 0 or n1 % 256 == 0) and n1 % 16 == 0
    return n1 % 256 == 0 and n2 in (192, 256)


def dual_supported(a, n1: int, n2: int, k2: int = 0) -> bool:
    return a.stride(1) == 1 and a.stride(0) % 8 == 0 and dual_supported_dims(a.shape[0], a.shape[1], a.dtype, n1, n2, k2)


def gemm_dual_ncat(a, w1, w2, out1, out2, epi1: dict, epi2: dict):
    """[out1 | out2] = [epi1(a @ w1^T) | epi2(a @ w2^T)] in one launch (same left operand)."""
    M, K = a.shape
    N1, N2 = w1.shape[0], w2.shape[0]
    assert w1.shape[1] == K and w2.shape[1] == K and out1.shape == (M, N1) and out2.shape == (M, N2)
    e1 = make_epilogue(out1, ldo=out1.stride(0), **epi1)
    e2 = make_epilogue(out2, ldo=out2.stride(0), **epi2)
    _count()
    _chk(load().aimb_gemm_dual(0, _ptr(a), a.stride(0), _ptr(w1), w1.stride(0), None, 0, _ptr(w2), w2.stride(0), C.byref(e1),
                               C.byref(e2), None, None, 0, 1.0, M, N1, N2, K, 0, dt_code(a), _stream()),
         f"gemm_dual_ncat[{M}x({N1}+{N2})x{K}]")


def gemm_dual_kcat(a1, w1, a2, w2, out, bias2=None, bias2_row_scale=None, bias2_scale=1.0, **epi):
    """out = epilogue(a1 @ w1^T + a2 @ w2^T + bias2 * bias2_scale * bias2_row_scale[m % len]) in one launch."""
    M, K1 = a1.shape
    K2 = a2.shape[1]
    N = w1.shape[0]
    assert w1.shape == (N, K1) and w2.shape == (N, K2) and a2.shape[0] == M and out.shape == (M, N)
    e = make_epilogue(out, ldo=out.stride(0), **epi)
    rs = bias2_row_scale
    _count()
    _chk(load().aimb_gemm_dual(1, _ptr(a1), a1.stride(0), _ptr(w1), w1.stride(0), _ptr(a2), a2.stride(0), _ptr(w2), w2.stride(0),
                               C.byref(e), None, _ptr(bias2), _ptr(rs), rs.numel() if rs is not None else 0, bias2_scale, M, N, 0,
                               K1, K2, dt_code(a1), _stream()), f"gemm_dual_kcat[{M}x{N}x({K1}+{K2})]")
    return out


def gemm_strided(a, a_sm, a_sk, b, b_sn, b_sk, out, M, N, K, **epi):
    e = make_epilogue(out, ldo=out.stride(0), **epi)
    _count()
    _chk(load().aimb_gemm_strided(_ptr(a), a_sm, a_sk, _ptr(b), b_sn, b_sk, C.byref(e), M, N, K, dt_code(a), _stream()),
         "gemm_strided")
    return out


def gemm_wgrad(dy, x, dw, alpha=1.0, accumulate=False, impl=IMPL_AUTO):
    """dw[N,K] (fp32) (+)= alpha * dy[R,N]^T @ x[R,K]"""
    R, N = dy.shape
    K = x.shape[1]
    assert x.shape[0] == R and dw.dtype == torch.float32 and dw.shape == (N, K) and dw.is_contiguous()
    _count()
    _chk(load().aimb_gemm_wgrad(_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(dw), R, N, K, alpha, int(accumulate),
                                dt_code(dy), impl, _stream()), "gemm_wgrad")
    return dw


def colsum(x, out, row_scale=None, alpha=1.0, accumulate=False):
    R, Cn = x.shape
    assert out.dtype == torch.float32 and x.stride(1) == 1
    _count()
    _chk(load().aimb_colsum(_ptr(x), x.stride(0), _ptr(row_scale), row_scale.numel() if row_scale is not None else 0,
                            alpha, _ptr(out), R, Cn, int(accumulate), dt_code(x), _stream()), "colsum")
    return out


def transpose(src, dst):
    R, Cn = src.shape
    _count()
    _chk(load().aimb_transpose(_ptr(_c(src)), _ptr(_c(dst)), R, Cn, dt_code(src), _stream()), "transpose")
    return dst


def adamw_flat(p, g, m, v, wd_mask, step, lr, beta1, beta2, eps, weight_decay):
    """one AdamW step over flat fp32 arrays; step = 1-element fp32 CUDA tensor holding the (already incremented) step count"""
    n = p.numel()
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and g.numel() == n and step.dtype == torch.float32
    assert wd_mask is None or (wd_mask.dtype == torch.uint8 and wd_mask.numel() == n)
    _count()
    _chk(load().aimb_adamw_flat(_ptr(_c(p)), _ptr(_c(g)), _ptr(_c(m)), _ptr(_c(v)), _ptr(wd_mask), _ptr(step), lr, beta1, beta2,
                                eps, weight_decay, n, _stream()), "adamw_flat")


def transpose_batched(src_flat, dst_flat, table_dev, nmat):
    """table_dev: int64 CUDA tensor [nmat, 3] = (element offset, rows, cols) into src_flat / dst_flat."""
    assert table_dev.dtype == torch.int64 and table_dev.is_cuda and src_flat.dtype == dst_flat.dtype
    _count()
    _chk(load().aimb_transpose_batched(_ptr(_c(src_flat)), _ptr(_c(dst_flat)), _ptr(table_dev), nmat, dt_code(src_flat),
                                       _stream()), "transpose_batched")


def attn_spatial_fwd(qkv, o, lse, frames, n, heads, impl=IMPL_AUTO):
    _count()
    _chk(load().aimb_attn_spatial_fwd(_ptr(_c(qkv)), _ptr(_c(o)), _ptr(lse), frames, n, heads, dt_code(qkv), impl,
                                      _stream()), "attn_spatial_fwd")


def attn_spatial_bwd(qkv, o, d_o, lse, d_qkv, frames, n, heads, impl=IMPL_AUTO):
    _count()
    _chk(load().aimb_attn_spatial_bwd(_ptr(_c(qkv)), _ptr(_c(o)), _ptr(_c(d_o)), _ptr(lse), _ptr(_c(d_qkv)), frames, n,
                                      heads, dt_code(qkv), impl, _stream()), "attn_spatial_bwd")


def attn_temporal_fwd(qkv, o, B, T, n, heads):
    _count()
    _chk(load().aimb_attn_temporal_fwd(_ptr(_c(qkv)), _ptr(_c(o)), B, T, n, heads, dt_code(qkv), _stream()),
         "attn_temporal_fwd")


def attn_temporal_bwd(qkv, d_o, d_qkv, B, T, n, heads):
    _count()
    _chk(load().aimb_attn_temporal_bwd(_ptr(_c(qkv)), _ptr(_c(d_o)), _ptr(_c(d_qkv)), B, T, n, heads, dt_code(qkv),
                                       _stream()), "attn_temporal_bwd")


def fork_weights(qkv, kc, w_o, w_c, frames, n, D):
    _count()
    _chk(load().aimb_fork_weights(_ptr(_c(qkv)), _ptr(_c(kc)), _ptr(w_o), _ptr(w_c), frames, n, D, dt_code(qkv),
                                  _stream()), "fork_weights")


def fork_combine(x, a_o, s_frame, lam, rs, out, BT, n):
    _count()
    _chk(load().aimb_fork_combine(_ptr(_c(x)), _ptr(_c(a_o)), _ptr(_c(s_frame)), _ptr(lam), _ptr(rs), _ptr(_c(out)), BT, n,
                                  x.shape[-1], dt_code(x), _stream()), "fork_combine")


def fork_combine_bwd(dx, lam, rs, d_ao, d_s, BT, n):
    _count()
    _chk(load().aimb_fork_combine_bwd(_ptr(_c(dx)), _ptr(lam), _ptr(rs), _ptr(_c(d_ao)), _ptr(_c(d_s)), BT, n, dx.shape[-1],
                                      dt_code(dx), _stream()), "fork_combine_bwd")

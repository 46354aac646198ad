"""Explicit forward / adapter-only backward of the AIM ViT_CLIP backbone on the aimb200 kernels.

The engine owns no parameters.  ``forward`` takes the clip and a dict of weights already in the
compute dtype, runs the stem, the L blocks and the tail as a fixed sequence of C-ABI calls and (in
training) keeps exactly the tensors the hand-written backward needs.  ``backward`` walks the blocks
in reverse: grad-input flows through every frozen GEMM (pre-transposed frozen weights, same tcgen05
kernel), weight/bias grads are produced only for the adapters, ``ln_post`` and
``temporal_embedding`` (freeze rule of vit_clip.py:413-415) and are written straight into a flat
fp32 gradient buffer whose per-block slices can be all-reduced while earlier blocks are still in
backward.

Reference math: block 'aim' = vitclip_aim.py:196-211, stem = vit_clip.py:433-447, tail = :450-456.
Rows of every activation are (b*T + t)*n + token (frame-major), so spatial attention reads
contiguous row ranges and temporal attention reads rows at stride n in place.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch

from . import lib

ADAPTERS_AIM = ("T_Adapter", "S_Adapter", "MLP_Adapter")


@dataclass(frozen=True)
class Dims:
    B: int
    T: int
    n: int
    D: int
    heads: int
    L: int
    r: int
    patch: int
    res: int
    kpad: int
    num_tadapter: int
    scale: float
    block: str = "aim"
    # ViT_ImageNet variant of the same block (vit_imagenet.py:88-126, 232-260): timm LayerNorm eps, exact-GELU MLP, no ln_pre
    eps: float = 1e-5
    mlp_act: int = lib.ACT_QUICKGELU
    ln_pre: bool = True

    @property
    def BT(self):
        return self.B * self.T

    @property
    def M(self):
        return self.B * self.T * self.n

    @property
    def G(self):
        return self.res // self.patch


class Engine:
    def __init__(self, dtype: torch.dtype, device: torch.device):
        assert dtype in (torch.float32, torch.bfloat16)
        self.dtype = dtype
        self.device = device
        self._bufs: Dict[tuple, torch.Tensor] = {}
        self._buf_mode: Dict[tuple, str] = {}      # which mode ("train" / "eval") created a buffer
        self._mode_sig: Dict[str, tuple] = {}      # problem size last seen per mode
        self._cur_mode = "eval"
        self.saved: Optional[dict] = None
        self.attn_impl = lib.IMPL_AUTO
        self.grads_prezeroed = False
        self.ckpt = False
        self.gemm_impl = lib.IMPL_AUTO
        # T_Adapter has no skip and is the only consumer of the temporal out_proj (vitclip_aim.py:203-204), so
        # D_fc1(out_proj(o)) = o (W1 Wo)^T + (W1 bo + b1): the temporal out_proj GEMM and its dgrad are never launched
        # (-24 GEMMs of 12608 x 768 x 768 per step, -5 % of the step's FLOPs, a_t is never written).  The per-step products
        # W1 Wo, their transposes and dW1 = (d_h^T o) Wo^T + db1 (x) bo are batched over all blocks (3 launches in forward,
        # 4 per gradient bucket in backward); the first version issued them per block (96 tiny launches) and lost more
        # than the removed GEMMs gave (675 vs 709 clips/s); batched: 731 -> 753 clips/s.
        self.fuse_t_outproj = os.environ.get("AIMB200_FUSE_T_OUTPROJ", "1") == "1"
        self.t_fused = False
        # S_Adapter (skip connection, vitclip_aim.py:208): x2 = x1 + a + fc2(gelu(fc1(a))), a = out_proj(o_s).  With the same
        # product, h = o_s (W1 Wo)^T + b1o and x2 = x1 + [o_s | g] [Wo | W2]^T + bo + b2: out_proj and D_fc2 share ONE
        # K-concatenated launch and a is never materialised; backward: d_o = [dx2 | d_h] [Wo^T | W1o^T]^T the same way.
        self.fuse_s_outproj = os.environ.get("AIMB200_FUSE_S_OUTPROJ", "1") == "1"
        self.s_fused = False
        self.one_kernel_t = os.environ.get("AIMB200_ONE_KERNEL_T", "0") == "1"   # experiment: adapter_tc_kernel for T_Adapter
        # ln_1 is frozen, so LN(x) Wqkv^T + b = rstd * (x (Wqkv * gamma)^T - mean * rowsum(Wqkv * gamma)) + (b + Wqkv beta):
        # the QKV GEMM reads the un-normalised residual stream and applies the row statistics in its epilogue; the
        # normalised activations are never written (a statistics-only pass replaces the LayerNorm kernel)
        self.ln_fold = os.environ.get("AIMB200_LN_FOLD", "1") == "1"
        self._ln_fold_cache: Dict[int, tuple] = {}
        self.fuse_adapters = os.environ.get("AIMB200_FUSE_ADAPTERS", "0") == "1"   # opt-in: measured on par at M = 12608
        # MLP_Adapter shares its input with mlp.c_fc and its sum with mlp.c_proj (vitclip_aim.py:210-211): its two GEMMs ride
        # on the frozen ones as N- / K-concatenated segments of ONE launch each, forward and backward (-4 launches per block)
        self.pair_mlp = os.environ.get("AIMB200_PAIR_MLP", "1") == "1"
        # adapter weight-gradient kernels are off the critical path of backward: they run on a side stream (captured
        # into the step graph as parallel branches) and fill the SMs the small dgrad GEMMs / kernel tails leave idle
        self.wgrad_side = os.environ.get("AIMB200_WGRAD_STREAM", "1") == "1"
        self._side = None
        self._side_busy = False
        # bias-gradient column sums that used to be fused into the LayerNorm backward (slower variant of the kernel,
        # on the critical path) run as a separate reduction on a second side stream
        self.ln_colsum_side = os.environ.get("AIMB200_LN_COLSUM_STREAM", "1") == "1"
        self._side2 = None
        self._side2_busy = False

    # ------------------------------------------------------------------ buffers (stable pointers across steps)
    def buf(self, name, shape, dtype=None, key=None):
        dtype = dtype or self.dtype
        k = (name, tuple(shape), dtype, key)
        t = self._bufs.get(k)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[k] = t
            self._buf_mode[k] = self._cur_mode
        return t

    def _enter_mode(self, mode: str, d: "Dims"):
        """Buffers are cached by exact shape (stable pointers for TMA descriptors / CUDA graphs).  When the batch size of
        a mode changes (last partial batch, a different number of views), the old set is dropped instead of living next
        to the new one: at 32-frame ViT-L sizes a second set of saved activations is tens of GB."""
        self._cur_mode = mode
        sig = (d.B, d.T, d.n, d.D, d.L)
        if self._mode_sig.get(mode) not in (None, sig):
            for k in [k for k, m in self._buf_mode.items() if m == mode]:
                del self._bufs[k]
                del self._buf_mode[k]
        self._mode_sig[mode] = sig

    # ------------------------------------------------------------------ side stream for the weight-gradient kernels
    def _side_begin(self):
        """side stream, ordered after everything issued so far on the current stream"""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._side.wait_event(ev)
        self._side_busy = True
        return self._side

    def _join_side(self):
        """the current stream waits for the side stream: call before anything overwrites a buffer a wgrad reads"""
        if self._side_busy:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side_busy = False

    def _ln_bwd(self, dy, x, mean, rstd, gamma, dres, dx, colsum_out=None, colsum_row_scale=None, colsum_alpha=1.0):
        """LayerNorm backward (+ optional weighted column sums of its result = an adapter's fc2 bias gradient)"""
        if colsum_out is None or not (self.wgrad_side and self.ln_colsum_side):
            self._join_side2()
            return lib.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, colsum_out=colsum_out,
                                     colsum_row_scale=colsum_row_scale, colsum_alpha=colsum_alpha)
        self._join_side2()             # a previous column sum may still read the buffer this call rewrites
        lib.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx)
        if self._side2 is None:
            self._side2 = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._side2.wait_event(ev)
        with torch.cuda.stream(self._side2):
            self._colsum(dx, colsum_out, row_scale=colsum_row_scale, alpha=colsum_alpha)
        self._side2_busy = True
        return dx

    def _join_side2(self):
        if self._side2_busy:
            ev = torch.cuda.Event()
            ev.record(self._side2)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side2_busy = False

    def release(self):
        self._bufs.clear()
        self._buf_mode.clear()
        self._mode_sig.clear()
        self._ln_fold_cache.clear()
        self._t_stack_key = None
        self.saved = None

    def gemm(self, a, w, out, **kw):
        if self.grads_prezeroed and kw.get("colsum_out") is not None:
            kw["colsum_accumulate"] = True
        return lib.gemm_nt(a, w, out, impl=self.gemm_impl, **kw)

    # gradient outputs: when the caller has zeroed the whole flat gradient buffer once, the ~200 per-call memsets of the
    # weight-gradient / column-sum kernels (graph memset nodes between the kernels) are not issued
    def _wgrad(self, dy, x, dw, alpha=1.0):
        return lib.gemm_wgrad(dy, x, dw, alpha=alpha, accumulate=self.grads_prezeroed)

    def _colsum(self, x, out, row_scale=None, alpha=1.0):
        return lib.colsum(x, out, row_scale=row_scale, alpha=alpha, accumulate=self.grads_prezeroed)

    # ------------------------------------------------------------------ forward
    def prep_stream(self):
        """side stream of the per-step weight preparation (backbone.py::_run_forward)"""
        if getattr(self, "_prep", None) is None:
            self._prep = torch.cuda.Stream(device=self.device)
        return self._prep

    def forward(self, x: torch.Tensor, W: Dict[str, torch.Tensor], d: Dims, training: bool,
                drop_masks: Optional[List] = None, WT: Optional[Dict[str, torch.Tensor]] = None,
                checkpoint: bool = False, weights_ready=None) -> torch.Tensor:
        """x [B,3,T,H,W] (fp32 / bf16 / uint8) -> feat fp32 [B, D, T].  W: weights in compute dtype.
        drop_masks[i] = (mask_t, mask_m): fp32 [n] DropPath multipliers (0 or 1/keep) or None."""
        M, D, r, n, BT = d.M, d.D, d.r, d.n, d.BT
        sv = {"d": d, "blocks": []} if training else None
        key = "train" if training else "eval"
        self._enter_mode(key, d)
        # checkpoint=True (vit_clip.py:318-319, torch.utils.checkpoint per block): only the block INPUTS are kept; every
        # block's saved activations live in ONE shared buffer set and are recomputed block by block in backward
        self.ckpt = bool(checkpoint and training)
        self._eps = d.eps
        # weights_ready = (event, stream): the trainable weights (W / WT / drop_masks) were prepared on that side stream; the
        # batched weight products of the fused adapters are appended to it, and the main stream joins it twice: for
        # temporal_embedding (ready at `event`) before the stem assembly, and for everything before block 0
        main = torch.cuda.current_stream(self.device)
        self.t_fused = (self.fuse_t_outproj and d.block == "aim" and d.num_tadapter == 1 and WT is not None
                        and self.dtype == torch.bfloat16)
        self.s_fused = (self.fuse_s_outproj and d.block == "aim" and WT is not None and self.dtype == torch.bfloat16
                        and self.gemm_impl == lib.IMPL_AUTO and not self.fuse_adapters
                        and lib.dual_supported_dims(M, D, self.dtype, D, 0, r))
        with torch.cuda.stream(weights_ready[1]) if weights_ready is not None else contextlib.nullcontext():
            if self.t_fused:
                self._prep_fused("T_Adapter", W, WT, d, training)
            if self.s_fused:
                self._prep_fused("S_Adapter", W, WT, d, training)
        # ---- stem
        cols = self.buf("cols", (BT * d.G * d.G, d.kpad))
        lib.im2col(x, cols, d.patch, W.get("input_mean"), W.get("input_std"))
        tok = self.buf("tok", (BT * d.G * d.G, D))
        self.gemm(cols, W["conv1.weight"], tok, bias=W.get("conv1.bias"))       # patch_embed.proj has a bias in ViT_ImageNet
        xcur = self.buf("x", (M, D), key=(key, 0))
        mean0 = self.buf("ln_pre_mean", (M,), torch.float32, key)
        rstd0 = self.buf("ln_pre_rstd", (M,), torch.float32, key)
        if weights_ready is not None:
            main.wait_event(weights_ready[0])              # temporal_embedding comes out of the per-step cast
        if d.ln_pre:
            z = self.buf("z", (M, D), key=key) if training else None
            lib.stem_assemble_ln(tok, W["class_embedding"], W["positional_embedding"], W["temporal_embedding"],
                                 W["ln_pre.weight"], W["ln_pre.bias"], z, xcur, mean0, rstd0, d.B, d.T, n, eps=d.eps)
            if training:
                sv["z"], sv["ln_pre"] = z, (mean0, rstd0)
        else:
            # no ln_pre (vit_imagenet.py:244-251): the assembled sum IS the block input; the kernel's LayerNorm output goes
            # to a scratch buffer (one pass per step, not worth a second kernel variant)
            lib.stem_assemble_ln(tok, W["class_embedding"], W["positional_embedding"], W["temporal_embedding"],
                                 W["ln_post.weight"], W["ln_post.bias"], xcur, self.buf("stem_scratch", (M, D)), mean0, rstd0,
                                 d.B, d.T, n, eps=d.eps)
        if weights_ready is not None:
            main.wait_stream(weights_ready[1])
        for i in range(d.L):
            masks = drop_masks[i] if (drop_masks is not None) else (None, None)
            if d.block == "fork":
                xcur = self._block_fwd_fork(i, xcur, W, d, training, masks, sv)
            else:
                xcur = self._block_fwd(i, xcur, W, d, training, masks, sv)
        # ---- tail
        feat = torch.empty(d.B, D, d.T, device=self.device, dtype=torch.float32)
        tm = self.buf("tail_mean", (BT,), torch.float32, key)
        tr = self.buf("tail_rstd", (BT,), torch.float32, key)
        lib.tail_fwd(xcur, W["ln_post.weight"], W["ln_post.bias"], feat, tm, tr, d.B, d.T, n, eps=d.eps)
        if training:
            sv["x_last"], sv["tail"] = xcur, (tm, tr)
            self.saved = sv
        return feat

    def _per_block(self, T: Dict[str, torch.Tensor], leaf: str, d: Dims) -> torch.Tensor:
        """[L, *shape] strided view of one per-block trainable tensor (parameters / gradients of all blocks live in ONE flat
        buffer with a uniform block stride: backbone.py::_flatten_trainable)."""
        t0 = T[f"transformer.resblocks.0.{leaf}"]
        if d.L == 1:
            return t0.unsqueeze(0)
        t1 = T[f"transformer.resblocks.1.{leaf}"]
        step = (t1.data_ptr() - t0.data_ptr()) // t0.element_size()
        return t0.as_strided((d.L,) + tuple(t0.shape), (step,) + tuple(t0.stride()))

    def _t_frozen_stacks(self, W, WT, d):
        """out_proj weights of all blocks stacked once (frozen): Wo [L, D_out, D_in], Wo^T, bo (compute dtype and fp32)."""
        key = tuple(W[f"transformer.resblocks.{i}.attn.out_proj.weight"].data_ptr() for i in range(d.L))
        if getattr(self, "_t_stack_key", None) != key:
            pre = "transformer.resblocks.{}."
            Wo = torch.stack([W[pre.format(i) + "attn.out_proj.weight"] for i in range(d.L)])
            WoT = torch.stack([WT[pre.format(i) + "attn.out_proj.weight"] for i in range(d.L)])
            bo = torch.stack([W[pre.format(i) + "attn.out_proj.bias"] for i in range(d.L)])
            self._t_stack_key, self._t_stack = key, (Wo, WoT, bo, bo.float())
        return self._t_stack

    def _prep_fused(self, ad, W, WT, d, training):
        """Per step (D_fc1 is trainable), for ALL blocks in three batched launches: W1o = W1 Wo [r, D], b1o = W1 bo + b1,
        and W1o^T for the dgrad.  These are weight x weight products (0.23 GFLOP per block), not part of the activation
        path: one strided-batched library GEMM each instead of 36 tile-kernel launches."""
        r, D, L = d.r, d.D, d.L
        Wo, WoT, bo, _ = self._t_frozen_stacks(W, WT, d)
        W1 = self._per_block(W, ad + ".D_fc1.weight", d)                        # [L, r, D]
        b1 = self._per_block(W, ad + ".D_fc1.bias", d)                          # [L, r]
        w1o = self.buf(ad + "_w1o", (L, r, D))
        torch.bmm(W1, Wo, out=w1o)                                              # sum_j W1[r, j] Wo[j, i]
        b1o = self.buf(ad + "_b1o", (L, 1, r))
        torch.baddbmm(b1.unsqueeze(1), bo.unsqueeze(1), W1.transpose(1, 2), out=b1o)
        if training:
            w1oT = self.buf(ad + "_w1oT", (L, D, r))                            # the N x K operand of d_o = d_h W1o
            w1oT.copy_(w1o.transpose(1, 2))
            self.buf(ad + "_G", (L, r, D), torch.float32).zero_()               # d_h^T o of every block, see _flush_fused
        for i in range(L):
            pre = f"transformer.resblocks.{i}."
            W[pre + ad + ".w1o"], W[pre + ad + ".b1o"] = w1o[i], b1o[i].view(r)
            if training:
                W[pre + ad + ".w1oT"] = w1oT[i]

    def _flush_fused(self, ad, lo, hi, W, WT, grads, d):
        """<ad>.D_fc1 weight gradients of blocks lo..hi-1 from the accumulated G = d_h^T o:
        dW1 = G Wo^T + db1 (x) bo   (a = o Wo^T + bo was never materialised).  One batched launch chain per gradient
        bucket, issued when the bucket's last block has finished its backward."""
        if hi <= lo:
            return
        r, D, L = d.r, d.D, d.L
        _, WoT, _, bo32 = self._t_frozen_stacks(W, WT, d)
        self._join_side()                                                       # the wgrads that fill G run on the side stream
        G = self.buf(ad + "_G", (L, r, D), torch.float32)[lo:hi]
        Gb = self.buf(ad + "_Gb", (L, r, D))[lo:hi]
        Gb.copy_(G)
        Gw = self.buf(ad + "_Gw", (L, r, D))[lo:hi]
        torch.bmm(Gb, WoT[lo:hi], out=Gw)                                       # sum_j G[r, j] Wo[j', j]
        dW1 = self._per_block(grads, ad + ".D_fc1.weight", d)[lo:hi]
        db1 = self._per_block(grads, ad + ".D_fc1.bias", d)[lo:hi]
        dW1.copy_(Gw)
        dW1.baddbmm_(db1.unsqueeze(2), bo32[lo:hi].unsqueeze(1))

    def _ln_folded(self, i, W):
        """(W * gamma in the compute dtype, fp32 row sums of exactly those values, b + W beta) of block i's in_proj / ln_1;
        cached while the frozen source tensors are the same objects (the backbone replaces them when a parameter changes)."""
        pre = f"transformer.resblocks.{i}."
        src = (W[pre + "attn.in_proj_weight"], W[pre + "attn.in_proj_bias"], W[pre + "ln_1.weight"], W[pre + "ln_1.bias"])
        ent = self._ln_fold_cache.get(i)
        if ent is None or any(a is not b for a, b in zip(ent[0], src)):
            w, b, g, be = (t.float() for t in src)
            wf = (w * g[None, :]).to(self.dtype).contiguous()
            ent = (src, wf, wf.float().sum(1).contiguous(), (b + w @ be).to(self.dtype).contiguous())
            self._ln_fold_cache[i] = ent
        return ent[1], ent[2], ent[3]

    def _qkv_ln(self, i, x, W, out, mean, rstd, fold, xn_key=None):
        """out = ln_1(x) Wqkv^T + b (vit_clip.py:132-138 after :71-77); returns ln_1(x) when it was materialised."""
        pre = f"transformer.resblocks.{i}."
        if fold:
            lib.layernorm_fwd(x, W[pre + "ln_1.weight"], W[pre + "ln_1.bias"], None, mean, rstd, eps=self._eps)      # statistics only
            wf, ws, cf = self._ln_folded(i, W)
            self.gemm(x, wf, out, bias=cf, ln_mean=mean, ln_rstd=rstd, ln_wsum=ws)
            return None
        xn = self.buf("xn", x.shape, key=xn_key)
        lib.layernorm_fwd(x, W[pre + "ln_1.weight"], W[pre + "ln_1.bias"], xn, mean, rstd, eps=self._eps)
        self.gemm(xn, W[pre + "attn.in_proj_weight"], out, bias=W[pre + "attn.in_proj_bias"])
        return xn

    def _adapter_fwd(self, name, pre, a, W, d, bk, training, rs, alpha, res1, res2, out):
        """out = res1 + res2 + alpha * rs * (fc2(gelu(fc1(a))))   (rs folded into the hidden, see backward)."""
        M, r = a.shape[0], d.r
        h = self.buf(name + "_h", (M, r), key=bk) if training else None
        g = self.buf(name + "_g", (M, r), key=bk)
        w1, w2 = W[pre + name + ".D_fc1.weight"], W[pre + name + ".D_fc2.weight"]
        epi1 = dict(bias=W[pre + name + ".D_fc1.bias"], act=lib.ACT_GELU, out_pre=h, row_scale=rs)
        epi2 = dict(bias=W[pre + name + ".D_fc2.bias"], row_scale=rs, bias_rowscaled=rs is not None, alpha=alpha,
                    res1=res1, res2=res2)
        if self.fuse_adapters and self.gemm_impl == lib.IMPL_AUTO and lib.adapter_fused_supported(a, r, d.D):
            lib.adapter_fused(a, w1, w2, g, out, epi1, epi2)      # one kernel: the hidden tile stays on chip
        else:
            self.gemm(a, w1, g, **epi1)
            self.gemm(g, w2, out, **epi2)
        return h, g

    def _mlp_fwd(self, i, x2, W, d, training, mask_m, bk):
        """joint adaptation (vitclip_aim.py:210-211 == vit_clip.py:285-286):
        x_out = x2 + c_proj(QuickGELU(c_fc(ln_2(x2)))) + drop_path(scale * MLP_Adapter(ln_2(x2)))"""
        M, D = d.M, d.D
        pre = f"transformer.resblocks.{i}."
        f32 = torch.float32
        xn2 = self.buf("xn2", (M, D), key=bk)
        m3, r3 = self.buf("ln2_m", (M,), f32, bk), self.buf("ln2_r", (M,), f32, bk)
        lib.layernorm_fwd(x2, W[pre + "ln_2.weight"], W[pre + "ln_2.bias"], xn2, m3, r3, eps=d.eps)
        hf = self.buf("hf", (M, 4 * D), key=bk) if training else None
        gf = self.buf("gf", (M, 4 * D))
        xo = self.buf("x", (M, D), key=("train", i + 1) if training else ("eval", (i + 1) % 2))
        if d.mlp_act == lib.ACT_QUICKGELU and self._pair_mlp_ok(xn2, d):
            # [hf | h_m] = xn2 [Wfc ; W1]^T in one launch; g_m = scale * mask_m * GELU(h_m) carries the whole branch factor
            h_m = self.buf("MLP_Adapter_h", (M, d.r), key=bk) if training else None
            g_m = self.buf("MLP_Adapter_g", (M, d.r), key=bk)
            lib.gemm_dual_ncat(xn2, W[pre + "mlp.c_fc.weight"], W[pre + "MLP_Adapter.D_fc1.weight"], gf, g_m,
                               dict(bias=W[pre + "mlp.c_fc.bias"], act=lib.ACT_QUICKGELU, out_pre=hf),
                               dict(bias=W[pre + "MLP_Adapter.D_fc1.bias"], act=lib.ACT_GELU, out_pre=h_m, row_scale=mask_m,
                                    alpha=d.scale))
            # x_out = [gf | g_m] [Wp | W2]^T + bp + scale * mask_m * b2 + x2
            lib.gemm_dual_kcat(gf, W[pre + "mlp.c_proj.weight"], g_m, W[pre + "MLP_Adapter.D_fc2.weight"], xo,
                               bias2=W[pre + "MLP_Adapter.D_fc2.bias"], bias2_row_scale=mask_m, bias2_scale=d.scale,
                               bias=W[pre + "mlp.c_proj.bias"], res1=x2)
            return xo, dict(ln2=(m3, r3), xn2=xn2, h_m=h_m, g_m=g_m, hf=hf, paired=True)
        tmp = self.buf("tmp", (M, D))
        h_m, g_m = self._adapter_fwd("MLP_Adapter", pre, xn2, W, d, bk, training, mask_m, d.scale, x2, None, tmp)
        self.gemm(xn2, W[pre + "mlp.c_fc.weight"], gf, bias=W[pre + "mlp.c_fc.bias"], act=d.mlp_act, out_pre=hf)
        self.gemm(gf, W[pre + "mlp.c_proj.weight"], xo, bias=W[pre + "mlp.c_proj.bias"], res1=tmp)
        return xo, dict(ln2=(m3, r3), xn2=xn2, h_m=h_m, g_m=g_m, hf=hf, paired=False)

    def _pair_mlp_ok(self, xn2, d):
        return (self.pair_mlp and self.dtype == torch.bfloat16 and self.gemm_impl == lib.IMPL_AUTO and not self.fuse_adapters
                and lib.dual_supported(xn2, 4 * d.D, d.r) and lib.dual_supported(xn2, d.D, 0, d.r))

    def _mlp_bwd(self, i, dx, W, WT, grads, d, S, mask_m, db2_fused, colsum_for=None):
        """backward of _mlp_fwd: returns dx2 = d(x2) (accumulated in place in `dx`)."""
        M, D = d.M, d.D
        pre = f"transformer.resblocks.{i}."
        d_hf = self.buf("d_big", (M, 4 * D))
        d_xn2 = self.buf("d_xn", (M, D))
        if S.get("paired"):
            k1w, k1b = pre + "MLP_Adapter.D_fc1.weight", pre + "MLP_Adapter.D_fc1.bias"
            k2w, k2b = pre + "MLP_Adapter.D_fc2.weight", pre + "MLP_Adapter.D_fc2.bias"
            d_h = self.buf("d_h_m", (M, d.r))
            self._join_side()          # the previous block's fc1 wgrad / db1 column sums still read d_h_m
            # [d_hf | d_h] = dx [Wp^T ; W2^T]^T . [QuickGELU'(hf) | scale * mask_m * GELU'(h_m)]
            lib.gemm_dual_ncat(dx, WT[pre + "mlp.c_proj.weight"], WT[k2w], d_hf, d_h,
                               dict(dact_src=S["hf"], dact=lib.ACT_QUICKGELU),
                               dict(dact_src=S["h_m"], dact=lib.ACT_GELU, alpha=d.scale, row_scale=mask_m))

            def _wgrads():             # g_m already carries scale * mask_m
                self._wgrad(dx, S["g_m"], grads[k2w])
                self._colsum(d_h, grads[k1b])
                self._wgrad(d_h, S["xn2"], grads[k1w])

            if self.wgrad_side:
                with torch.cuda.stream(self._side_begin()):
                    _wgrads()
            if not db2_fused:
                self._colsum(dx, grads[k2b], row_scale=mask_m, alpha=d.scale)
            lib.gemm_dual_kcat(d_hf, WT[pre + "mlp.c_fc.weight"], d_h, WT[k1w], d_xn2)      # d_xn2 = d_hf Wfc + d_h W1
            if not self.wgrad_side:
                _wgrads()
            self._join_side()          # dW2 reads dx, which the LayerNorm backward below rewrites in place
        else:
            self.gemm(dx, WT[pre + "mlp.c_proj.weight"], d_hf, dact_src=S["hf"], dact=d.mlp_act)
            self.gemm(d_hf, WT[pre + "mlp.c_fc.weight"], d_xn2)
            self._adapter_bwd("MLP_Adapter", pre, dx, S["xn2"], S["h_m"], S["g_m"], W, WT, grads, d, mask_m, d.scale,
                              d_xn2, d_xn2, db2_fused=db2_fused)
        m3, r3 = S["ln2"]
        self._ln_bwd(d_xn2, S["x2"], m3, r3, W[pre + "ln_2.weight"], dx, dx, colsum_out=colsum_for)
        return dx

    def _block_fwd(self, i, x, W, d, training, masks, sv):
        M, D, r, n = d.M, d.D, d.r, d.n
        pre = f"transformer.resblocks.{i}."
        bk = (("train", "ckpt") if self.ckpt else ("train", i)) if training else "eval"   # per-block buffers only when they must survive
        mask_t, mask_m = masks
        S = {} if training else None
        f32 = torch.float32
        Wqkv, bqkv = W[pre + "attn.in_proj_weight"], W[pre + "attn.in_proj_bias"]
        Wo, bo = W[pre + "attn.out_proj.weight"], W[pre + "attn.out_proj.bias"]
        ln1w, ln1b = W[pre + "ln_1.weight"], W[pre + "ln_1.bias"]
        # ---------------- temporal adaptation (vitclip_aim.py:199-206)
        fold = self.ln_fold and self.dtype == torch.bfloat16 and self.gemm_impl == lib.IMPL_AUTO and M >= 128 and D % 64 == 0
        m1, r1 = self.buf("ln1t_m", (M,), f32, bk), self.buf("ln1t_r", (M,), f32, bk)
        qkv_t = self.buf("qkv_t", (M, 3 * D), key=bk)
        if d.num_tadapter == 2:
            xn = self.buf("xn", (M, D), key=bk if training else None)
            lib.layernorm_fwd(x, ln1w, ln1b, xn, m1, r1, eps=d.eps)
            xin = self.buf("xn_in", (M, D), key=bk)
            hi, gi = self._adapter_fwd("T_Adapter_in", pre, xn, W, d, bk, training, None, 1.0, xn, None, xin)
            if training:
                S["tin"] = (xn, hi, gi)
            self.gemm(xin, Wqkv, qkv_t, bias=bqkv)
        else:
            self._qkv_ln(i, x, W, qkv_t, m1, r1, fold)
        o_t = self.buf("o_t", (M, D), key=bk)
        lib.attn_temporal_fwd(qkv_t, o_t, d.B, d.T, n, d.heads)
        x1 = self.buf("x1", (M, D), key=bk)
        if self.t_fused:
            a_t = None
            h_t = self.buf("T_Adapter_h", (M, r), key=bk) if training else None
            g_t = self.buf("T_Adapter_g", (M, r), key=bk)
            epi1 = dict(bias=W[pre + "T_Adapter.b1o"], act=lib.ACT_GELU, out_pre=h_t, row_scale=mask_t)
            epi2 = dict(bias=W[pre + "T_Adapter.D_fc2.bias"], row_scale=mask_t, bias_rowscaled=mask_t is not None, res1=x)
            if self.one_kernel_t and lib.adapter_fused_supported(o_t, r, D):
                # both GEMMs of the adapter in one kernel: the hidden tile stays in shared memory (adapter_tc_kernel)
                lib.adapter_fused(o_t, W[pre + "T_Adapter.w1o"], W[pre + "T_Adapter.D_fc2.weight"], g_t, x1, epi1, epi2)
            else:
                self.gemm(o_t, W[pre + "T_Adapter.w1o"], g_t, **epi1)
                self.gemm(g_t, W[pre + "T_Adapter.D_fc2.weight"], x1, **epi2)
        else:
            a_t = self.buf("a_t", (M, D), key=bk)
            self.gemm(o_t, Wo, a_t, bias=bo)
            h_t, g_t = self._adapter_fwd("T_Adapter", pre, a_t, W, d, bk, training, mask_t, 1.0, x, None, x1)
        # ---------------- spatial adaptation (:208)
        m2, r2 = self.buf("ln1s_m", (M,), f32, bk), self.buf("ln1s_r", (M,), f32, bk)
        qkv_s = self.buf("qkv_s", (M, 3 * D), key=bk)
        self._qkv_ln(i, x1, W, qkv_s, m2, r2, fold)
        o_s = self.buf("o_s", (M, D), key=bk)
        lse = self.buf("lse_s", (d.BT, d.heads, n), f32, bk) if training else None
        lib.attn_spatial_fwd(qkv_s, o_s, lse, d.BT, n, d.heads, impl=self.attn_impl)
        x2 = self.buf("x2", (M, D), key=bk)
        if self.s_fused:
            a_s = None
            h_s = self.buf("S_Adapter_h", (M, r), key=bk) if training else None
            g_s = self.buf("S_Adapter_g", (M, r), key=bk)
            self.gemm(o_s, W[pre + "S_Adapter.w1o"], g_s, bias=W[pre + "S_Adapter.b1o"], act=lib.ACT_GELU, out_pre=h_s)
            lib.gemm_dual_kcat(o_s, Wo, g_s, W[pre + "S_Adapter.D_fc2.weight"], x2, bias2=W[pre + "S_Adapter.D_fc2.bias"],
                               bias=bo, res1=x1)                       # x2 = x1 + (o_s Wo^T + bo) + (g W2^T + b2)
        else:
            a_s = self.buf("a_s", (M, D), key=bk)
            self.gemm(o_s, Wo, a_s, bias=bo)
            h_s, g_s = self._adapter_fwd("S_Adapter", pre, a_s, W, d, bk, training, None, 1.0, x1, a_s, x2)
        xo, mlp_saved = self._mlp_fwd(i, x2, W, d, training, mask_m, bk)
        if training:
            S.update(x=x, ln1t=(m1, r1), qkv_t=qkv_t, o_t=o_t, a_t=a_t, h_t=h_t, g_t=g_t, x1=x1, ln1s=(m2, r2),
                     qkv_s=qkv_s, o_s=o_s, lse=lse, a_s=a_s, h_s=h_s, g_s=g_s, x2=x2, masks=masks, **mlp_saved)
            sv["blocks"].append(S)
        return xo


    # ------------------------------------------------------------------ block 'fork' (vit_clip.py:199-288, shift=False)
    def _block_fwd_fork(self, i, x, W, d, training, masks, sv):
        """cls-only temporal attention kept aside as xt; spatial self-attention a_o; cross attention of every token to the
        single key xt_f (softmax == 1 -> a_c is one row per frame); lambda_f = w_c / (w_c + w_o) (no grad);
        x += (1 - lambda) a_o + drop_path(scale * S_Adapter(lambda a_c)); then the shared MLP half."""
        M, D, n, BT = d.M, d.D, d.n, d.BT
        pre = f"transformer.resblocks.{i}."
        bk = (("train", "ckpt") if self.ckpt else ("train", i)) if training else "eval"
        mask_s, mask_m = masks
        f32 = torch.float32
        Wqkv, bqkv = W[pre + "attn.in_proj_weight"], W[pre + "attn.in_proj_bias"]
        Wo, bo = W[pre + "attn.out_proj.weight"], W[pre + "attn.out_proj.bias"]
        xn = self.buf("xn", (M, D))
        m1, r1 = self.buf("ln1s_m", (M,), f32, bk), self.buf("ln1s_r", (M,), f32, bk)
        lib.layernorm_fwd(x, W[pre + "ln_1.weight"], W[pre + "ln_1.bias"], xn, m1, r1, eps=d.eps)
        # ---- temporal attention over the T cls tokens of each clip (:218-229); ln_1 is row-wise, so ln_1(cls) = cls rows of xn
        ct = xn.view(BT, n, D)[:, 0, :]                                    # [BT, D], row stride n*D (read in place)
        qkv_c = self.buf("qkv_c", (BT, 3 * D), key=bk)
        self.gemm(ct, Wqkv, qkv_c, bias=bqkv)
        o_c = self.buf("o_c", (BT, D), key=bk)
        lib.attn_temporal_fwd(qkv_c, o_c, d.B, d.T, 1, d.heads)
        a_ct = self.buf("a_ct", (BT, D), key=bk)
        self.gemm(o_c, Wo, a_ct, bias=bo)
        xt = self.buf("xt", (BT, D), key=bk)
        h_t, g_t = self._adapter_fwd("T_Adapter", pre, a_ct, W, d, bk, training, None, 1.0, None, None, xt)
        # ---- spatial self attention (:264, :128-162)
        qkv_s = self.buf("qkv_s", (M, 3 * D), key=bk)
        self.gemm(xn, Wqkv, qkv_s, bias=bqkv)
        o_s = self.buf("o_s", (M, D), key=bk)
        lse = self.buf("lse_s", (BT, d.heads, n), f32, bk) if training else None
        lib.attn_spatial_fwd(qkv_s, o_s, lse, BT, n, d.heads, impl=self.attn_impl)
        a_o = self.buf("a_s", (M, D))
        self.gemm(o_s, Wo, a_o, bias=bo)
        # ---- cross attention to xt (:265, :164-197): k, v from the raw xt; one key -> softmax == 1 -> a_c = out_proj(v_c)
        k_c = self.buf("k_c", (BT, D))
        v_c = self.buf("v_c", (BT, D))
        self.gemm(xt, Wqkv[D:2 * D], k_c, bias=bqkv[D:2 * D])
        self.gemm(xt, Wqkv[2 * D:], v_c, bias=bqkv[2 * D:])
        a_c = self.buf("a_c", (BT, D), key=bk)
        self.gemm(v_c, Wo, a_c, bias=bo)
        # ---- lambda from the un-normalised attention masses (:147-151, :182-186), fp32, no max subtraction
        w_o = self.buf("w_o", (BT,), f32)
        w_c = self.buf("w_c", (BT,), f32)
        lib.fork_weights(qkv_s, k_c, w_o, w_c, BT, n, D)
        lam = self.buf("lam", (BT,), f32, bk)
        torch.div(w_c, w_c + w_o, out=lam)
        # ---- S_Adapter (no skip in the fork) on one row per frame
        u = self.buf("u", (BT, D), key=bk)
        u.copy_(a_c.float() * lam.unsqueeze(1))
        s_fr = self.buf("s_fr", (BT, D))
        h_s, g_s = self._adapter_fwd("S_Adapter", pre, u, W, d, bk, training, None, d.scale, None, None, s_fr)
        x2 = self.buf("x2", (M, D), key=bk)
        lib.fork_combine(x, a_o, s_fr, lam, mask_s, x2, BT, n)
        xo, mlp_saved = self._mlp_fwd(i, x2, W, d, training, mask_m, bk)
        if training:
            sv["blocks"].append(dict(x=x, ln1s=(m1, r1), qkv_c=qkv_c, a_ct=a_ct, h_t=h_t, g_t=g_t, qkv_s=qkv_s, o_s=o_s,
                                     lse=lse, lam=lam, u=u, h_s=h_s, g_s=g_s, x2=x2, masks=masks, **mlp_saved))
        return xo

    def _block_bwd_fork(self, i, dx, W, WT, grads, d, S):
        M, D, n, BT = d.M, d.D, d.n, d.BT
        pre = f"transformer.resblocks.{i}."
        mask_s, mask_m = S["masks"]
        lam = S["lam"]
        WqkvT, WoT = WT[pre + "attn.in_proj_weight"], WT[pre + "attn.out_proj.weight"]   # [D, 3D], [D, D]
        dx2 = self._mlp_bwd(i, dx, W, WT, grads, d, S, mask_m, db2_fused=False)
        # x2 = x + (1 - lam) a_o + mask_s * s_fr   (lam carries no gradient: computed under no_grad in the reference)
        d_ao = self.buf("d_a", (M, D))
        d_s = self.buf("d_s", (BT, D))
        lib.fork_combine_bwd(dx2, lam, mask_s, d_ao, d_s, BT, n)
        d_u = self.buf("d_u", (BT, D))
        self._adapter_bwd("S_Adapter", pre, d_s, S["u"], S["h_s"], S["g_s"], W, WT, grads, d, None, d.scale, d_u, None)
        d_ac = self.buf("d_ac", (BT, D))
        d_ac.copy_(d_u.float() * lam.unsqueeze(1))
        d_vc = self.buf("d_vc", (BT, D))
        self.gemm(d_ac, WoT, d_vc)                                         # a_c = out_proj(v_c)
        d_xt = self.buf("d_xt", (BT, D))
        self.gemm(d_vc, WqkvT[:, 2 * D:], d_xt)                            # v_c = xt W_v^T + b_v ; k_c feeds only lam
        d_act = self.buf("d_act", (BT, D))
        self._adapter_bwd("T_Adapter", pre, d_xt, S["a_ct"], S["h_t"], S["g_t"], W, WT, grads, d, None, 1.0, d_act, None)
        d_oc = self.buf("d_oc", (BT, D))
        self.gemm(d_act, WoT, d_oc)
        d_qkv_c = self.buf("d_qkv_c", (BT, 3 * D))
        lib.attn_temporal_bwd(S["qkv_c"], d_oc, d_qkv_c, d.B, d.T, 1, d.heads)
        d_ct = self.buf("d_ct", (BT, D))
        self.gemm(d_qkv_c, WqkvT, d_ct)                                    # grad wrt ln_1(x) at the cls rows
        # spatial self attention
        d_os = self.buf("d_o", (M, D))
        self.gemm(d_ao, WoT, d_os)
        d_qkv = self.buf("d_qkv", (M, 3 * D))
        lib.attn_spatial_bwd(S["qkv_s"], S["o_s"], d_os, S["lse"], d_qkv, BT, n, d.heads, impl=self.attn_impl)
        d_xn = self.buf("d_xn", (M, D))
        self.gemm(d_qkv, WqkvT, d_xn)
        d_xn.view(BT, n, D)[:, 0, :].add_(d_ct)
        m1, r1 = S["ln1s"]
        lib.layernorm_bwd(d_xn, S["x"], m1, r1, W[pre + "ln_1.weight"], dx2, dx)
        return dx

    # ------------------------------------------------------------------ backward
    def backward(self, dfeat: torch.Tensor, W: Dict[str, torch.Tensor], WT: Dict[str, torch.Tensor],
                 grads: Dict[str, torch.Tensor], on_block_done: Optional[Callable[[int], None]] = None,
                 grads_prezeroed: bool = False, bucket_ends_at: Optional[Callable[[int], bool]] = None):
        """dfeat fp32 [B, D, T].  WT: transposed weights ([K,N] contiguous) for the dgrad GEMMs.
        grads: name -> fp32 tensor (views of the flat gradient buffer), overwritten.
        on_block_done(i) is called after block i's gradients are complete (i = L for ln_post, -1 for
        temporal_embedding) so the caller can start that bucket's all-reduce."""
        sv = self.saved
        if sv is None:
            raise lib.AimbError("Engine.backward() without a pending training forward")
        d: Dims = sv["d"]
        self._cur_mode = "train"
        self.grads_prezeroed = bool(grads_prezeroed)
        M, D, n = d.M, d.D, d.n
        dx = self.buf("dx", (M, D))
        tm, tr = sv["tail"]
        lib.tail_bwd(dfeat.contiguous(), sv["x_last"], tm, tr, W["ln_post.weight"], dx, grads["ln_post.weight"],
                     grads["ln_post.bias"], d.B, d.T, n)
        if on_block_done:
            on_block_done(d.L)
        t_hi = d.L                     # blocks [i, t_hi) still owe their T_Adapter.D_fc1 weight gradient (fused out_proj)
        for i in reversed(range(d.L)):
            prev_mask_m = sv["blocks"][i - 1]["masks"][1] if i > 0 else None
            S = sv["blocks"][i]
            self._cur_block = i
            if self.ckpt:
                # recompute this block's activations from its saved input (same DropPath masks -> identical values)
                self._join_side()
                self._join_side2()
                tmp = {"blocks": []}
                fwd = self._block_fwd_fork if d.block == "fork" else self._block_fwd
                fwd(i, S["x"], W, d, True, S["masks"], tmp)
                S = tmp["blocks"][0]
            if d.block == "fork":
                dx = self._block_bwd_fork(i, dx, W, WT, grads, d, S)
            else:
                dx = self._block_bwd(i, dx, W, WT, grads, d, S, prev_mask_m)
            if (self.t_fused or self.s_fused) and (i == 0 or (bucket_ends_at is not None and bucket_ends_at(i))):
                if self.t_fused:
                    self._flush_fused("T_Adapter", i, t_hi, W, WT, grads, d)
                if self.s_fused:
                    self._flush_fused("S_Adapter", i, t_hi, W, WT, grads, d)
                t_hi = i
            if on_block_done:
                on_block_done(i)
        # ln_pre backward -> dz ; temporal_embedding grad = sum over (b, token)   (vit_clip.py:443-447)
        if d.ln_pre:
            m0, r0 = sv["ln_pre"]
            dz = self.buf("dz", (M, D))
            lib.layernorm_bwd(dx, sv["z"], m0, r0, W["ln_pre.weight"], None, dz)
        else:
            dz = dx
        lib.temb_grad(dz, grads["temporal_embedding"], d.B, d.T, n)
        if on_block_done:
            on_block_done(-1)
        self.saved = None

    def _adapter_bwd(self, name, pre, dy, a, h, g, W, WT, grads, d, rs, alpha, d_a_out, d_a_res, db2_fused=False,
                     lazy_join=False):
        """y = alpha * rs * (fc2(gelu(fc1(a)))).  Given dy: adapter weight/bias grads, and
        d_a_out = d_a_res + d(a) (d_a_res may be None).  db2_fused: the fc2 bias gradient (a weighted column sum
        of dy) was already produced by the kernel that wrote dy.  lazy_join: the caller calls _join_side() itself
        before `dy` (or `a`, `g`) is overwritten, so the two wgrad kernels may overlap the kernels that follow."""
        M, r, D = dy.shape[0], d.r, d.D
        k1w, k1b = pre + name + ".D_fc1.weight", pre + name + ".D_fc1.bias"
        k2w, k2b = pre + name + ".D_fc2.weight", pre + name + ".D_fc2.bias"
        # fc2: g' = rs*gelu(h) was stored, so dW2 = alpha * dy^T g' ; db2 = alpha * sum_m rs[m] dy[m]
        self._join_side()              # the previous adapter's fc1 wgrad still reads the shared d_h scratch
        side = self.wgrad_side and not (self.fuse_adapters and lib.adapter_fused_supported(dy, r, D))
        if side:
            with torch.cuda.stream(self._side_begin()):
                self._wgrad(dy, g, grads[k2w], alpha=alpha)
        else:
            self._wgrad(dy, g, grads[k2w], alpha=alpha)
        if not db2_fused:
            self._colsum(dy, grads[k2b], row_scale=rs, alpha=alpha)
        # d_h = rs * alpha * (dy W2) * gelu'(h) ; db1 = column sums of d_h, taken in the same epilogue
        d_h = self.buf("d_h", (M, r))
        epi1 = dict(dact_src=h, dact=lib.ACT_GELU, alpha=alpha, row_scale=rs, colsum_out=grads[k1b])
        epi2 = dict(res1=d_a_res)
        if self.fuse_adapters and self.gemm_impl == lib.IMPL_AUTO and lib.adapter_fused_supported(dy, r, D):
            lib.adapter_fused(dy, WT[k2w], WT[k1w], d_h, d_a_out, epi1, epi2)
            self._wgrad(d_h, a, grads[k1w])
            return d_a_out
        self.gemm(dy, WT[k2w], d_h, **epi1)
        if side:
            with torch.cuda.stream(self._side_begin()):      # ordered after the d_h GEMM, concurrent with the d_a GEMM
                self._wgrad(d_h, a, grads[k1w])
        self.gemm(d_h, WT[k1w], d_a_out, **epi2)
        if not side:
            self._wgrad(d_h, a, grads[k1w])
        elif not lazy_join:
            self._join_side()
        return d_a_out

    def _t_fused_bwd(self, pre, dy, S, W, WT, grads, d, rs, d_o_out):
        """Backward of x1 = x + rs * fc2(gelu(o W1o^T + b1o)) without the out_proj activation a = o Wo^T + bo:
        d_o = d_h W1o;  dW1 = d_h^T a = (d_h^T o) Wo^T + colsum(d_h) (x) bo;  db1 = colsum(d_h).  The fc2 bias gradient was
        fused into the LayerNorm backward that produced dy.  Weight gradients run on the side stream."""
        M, r, D = dy.shape[0], d.r, d.D
        k1w, k1b, k2w = pre + "T_Adapter.D_fc1.weight", pre + "T_Adapter.D_fc1.bias", pre + "T_Adapter.D_fc2.weight"
        h, g, o = S["h_t"], S["g_t"], S["o_t"]
        self._join_side()
        side = self.wgrad_side
        if side:
            with torch.cuda.stream(self._side_begin()):
                self._wgrad(dy, g, grads[k2w])
        else:
            self._wgrad(dy, g, grads[k2w])
        d_h = self.buf("d_h", (M, r))
        G = self.buf("T_Adapter_G", (d.L, r, D), torch.float32)[self._cur_block]
        epi1 = dict(dact_src=h, dact=lib.ACT_GELU, row_scale=rs, colsum_out=grads[k1b],
                    colsum_accumulate=self.grads_prezeroed)
        if self.one_kernel_t and lib.adapter_fused_supported(dy, r, D):
            lib.adapter_fused(dy, WT[k2w], W[pre + "T_Adapter.w1oT"], d_h, d_o_out, epi1, {})
            if side:
                with torch.cuda.stream(self._side_begin()):
                    lib.gemm_wgrad(d_h, o, G, accumulate=True)
            else:
                lib.gemm_wgrad(d_h, o, G, accumulate=True)
            return
        lib.gemm_nt(dy, WT[k2w], d_h, impl=self.gemm_impl, **epi1)
        if side:
            with torch.cuda.stream(self._side_begin()):       # ordered after the d_h GEMM (and its fused db1 column sums)
                lib.gemm_wgrad(d_h, o, G, accumulate=True)                      # d_h^T o, fp32 (zeroed in _prep_fused)
        self.gemm(d_h, W[pre + "T_Adapter.w1oT"], d_o_out)
        if not side:
            lib.gemm_wgrad(d_h, o, G, accumulate=True)

    def _s_fused_bwd(self, pre, dy, S, W, WT, grads, d, d_o_out):
        """Backward of x2 = x1 + [o | g] [Wo | W2]^T + bo + b2, g = gelu(o W1o^T + b1o), dy = d(x2):
        d_h = (dy W2) . gelu'(h) (+ db1 column sums);  d_o = [dy | d_h] [Wo^T | W1o^T]^T in one K-concatenated launch;
        dW2 = dy^T g and G = d_h^T o on the side stream (dW1 = G Wo^T + db1 (x) bo in _flush_fused).  db2 = colsum(dy) was
        fused into the LayerNorm backward that produced dy.  The caller joins the side stream before dy is rewritten."""
        M, r, D = dy.shape[0], d.r, d.D
        k1b, k2w = pre + "S_Adapter.D_fc1.bias", pre + "S_Adapter.D_fc2.weight"
        h, g, o = S["h_s"], S["g_s"], S["o_s"]
        self._join_side()              # the previous adapter's weight gradients still read the shared d_h scratch
        side = self.wgrad_side
        if side:
            with torch.cuda.stream(self._side_begin()):
                self._wgrad(dy, g, grads[k2w])
        else:
            self._wgrad(dy, g, grads[k2w])
        d_h = self.buf("d_h", (M, r))
        self.gemm(dy, WT[k2w], d_h, dact_src=h, dact=lib.ACT_GELU, colsum_out=grads[k1b])
        G = self.buf("S_Adapter_G", (d.L, r, D), torch.float32)[self._cur_block]
        if side:
            with torch.cuda.stream(self._side_begin()):
                lib.gemm_wgrad(d_h, o, G, accumulate=True)
        lib.gemm_dual_kcat(dy, WT[pre + "attn.out_proj.weight"], d_h, W[pre + "S_Adapter.w1oT"], d_o_out)
        if not side:
            lib.gemm_wgrad(d_h, o, G, accumulate=True)

    def _block_bwd(self, i, dx, W, WT, grads, d, S, prev_mask_m=None):
        M, D, n = d.M, d.D, d.n
        pre = f"transformer.resblocks.{i}."
        mask_t, mask_m = S["masks"]
        # ---------------- joint adaptation: x_out = x2 + mlp(xn2) + scale*mask_m*MLP_Adapter(xn2)
        # (for i < L-1 the MLP-adapter fc2 bias grad was fused into the LN backward of block i+1 that produced dx;
        #  the LN2 backward below produces dx2 and, fused, the S_Adapter fc2 bias grad = column sums of dx2)
        dx2 = self._mlp_bwd(i, dx, W, WT, grads, d, S, mask_m, db2_fused=(i < d.L - 1),
                            colsum_for=grads[pre + "S_Adapter.D_fc2.bias"])
        # ---------------- spatial: x2 = x1 + a_s + S_Adapter_noskip(a_s)
        d_os = self.buf("d_o", (M, D))
        if self.s_fused:
            self._s_fused_bwd(pre, dx2, S, W, WT, grads, d, d_os)
        else:
            d_as = self.buf("d_a", (M, D))
            self._adapter_bwd("S_Adapter", pre, dx2, S["a_s"], S["h_s"], S["g_s"], W, WT, grads, d, None, 1.0, d_as, dx2,
                              db2_fused=True, lazy_join=True)
            self.gemm(d_as, WT[pre + "attn.out_proj.weight"], d_os)
        d_qkv = self.buf("d_qkv", (M, 3 * D))
        lib.attn_spatial_bwd(S["qkv_s"], S["o_s"], d_os, S["lse"], d_qkv, d.BT, n, d.heads, impl=self.attn_impl)
        d_xn1 = self.buf("d_xn", (M, D))
        self.gemm(d_qkv, WT[pre + "attn.in_proj_weight"], d_xn1)
        m2, r2 = S["ln1s"]
        dx1 = dx
        self._join_side()              # S_Adapter's fc2 wgrad reads dx2, which this LN backward rewrites in place
        self._ln_bwd(d_xn1, S["x1"], m2, r2, W[pre + "ln_1.weight"], dx2, dx1,
                     colsum_out=grads[pre + "T_Adapter.D_fc2.bias"], colsum_row_scale=mask_t)
        # ---------------- temporal: x1 = x + mask_t * T_Adapter(attn(ln_1(x)))
        d_ot = self.buf("d_o", (M, D))
        if self.t_fused:
            self._t_fused_bwd(pre, dx1, S, W, WT, grads, d, mask_t, d_ot)
        else:
            d_at = self.buf("d_a", (M, D))
            self._adapter_bwd("T_Adapter", pre, dx1, S["a_t"], S["h_t"], S["g_t"], W, WT, grads, d, mask_t, 1.0, d_at, None,
                              db2_fused=True, lazy_join=True)
            self.gemm(d_at, WT[pre + "attn.out_proj.weight"], d_ot)
        lib.attn_temporal_bwd(S["qkv_t"], d_ot, d_qkv, d.B, d.T, n, d.heads)
        d_xn1t = self.buf("d_xn", (M, D))
        if d.num_tadapter == 2:
            self.gemm(d_qkv, WT[pre + "attn.in_proj_weight"], d_xn1t, colsum_out=grads[pre + "T_Adapter_in.D_fc2.bias"])
            xn, hi, gi = S["tin"]
            d_xn_tot = self.buf("d_xn_b", (M, D))
            self._adapter_bwd("T_Adapter_in", pre, d_xn1t, xn, hi, gi, W, WT, grads, d, None, 1.0, d_xn_tot, d_xn1t,
                              db2_fused=True)
            d_xn1t = d_xn_tot
        else:
            self.gemm(d_qkv, WT[pre + "attn.in_proj_weight"], d_xn1t)
        m1, r1 = S["ln1t"]
        self._join_side()              # T_Adapter's fc2 wgrad reads dx1 == dx, rewritten below; block grads complete
        if i > 0:   # dx is the output gradient of block i-1: its MLP-adapter fc2 bias grad = scale * sum_m mask_m[m] dx[m]
            self._ln_bwd(d_xn1t, S["x"], m1, r1, W[pre + "ln_1.weight"], dx1, dx,
                         colsum_out=grads[f"transformer.resblocks.{i - 1}.MLP_Adapter.D_fc2.bias"],
                         colsum_row_scale=prev_mask_m, colsum_alpha=d.scale)
        else:
            self._ln_bwd(d_xn1t, S["x"], m1, r1, W[pre + "ln_1.weight"], dx1, dx)
        return dx

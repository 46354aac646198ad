/* aimb200 — C ABI of the B200-native AIM `ViT_CLIP` hot path (libaimb200.so).
 *
 * The reference (bobochow/adapt-image-models) is 100 % Python and has no FFI of its own
 * (SURVEY.md §2.2); what it *launches* on this path are ATen / cuBLAS / cuDNN library calls.
 * Each entry point below replaces one group of those calls; the reference call site it
 * stands in for is cited as file:line (paths relative to the reference root).
 *
 * Conventions
 *   - plain device pointers + explicit shapes; the CALLER allocates every output/workspace
 *   - `dtype`: AIMB_F32 (fp32 storage + fp32 SIMT math; the 1e-3 parity mode) or
 *              AIMB_BF16 (bf16 storage, fp32 accumulate/statistics; tcgen05 tensor cores)
 *   - `stream`: a cudaStream_t passed as void*; every call only enqueues work on it
 *   - return 0 on success, negative AIMB_ERR_* otherwise; never throws, never exits
 *   - activations are row-major [rows, D]; row m = (b*T + t)*n + token  (frame-major; token 0 = cls,
 *     token 1+gy*G+gx = patch (gy,gx)); head h = columns h*64 .. h*64+63; q/k/v = columns
 *     0:D / D:2D / 2D:3D of the fused QKV buffer (== rows of attn.in_proj_weight)
 */
#ifndef AIMB200_H
#define AIMB200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIMB_OK 0
#define AIMB_ERR_ARG (-1)         /* bad shape / alignment / null pointer */
#define AIMB_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed */
#define AIMB_ERR_UNSUPPORTED (-3) /* configuration outside what the kernels implement */
#define AIMB_ERR_DRIVER (-4)      /* cuTensorMapEncodeTiled unavailable / failed */

#define AIMB_F32 0
#define AIMB_BF16 1
#define AIMB_U8 2 /* input clips only */

#define AIMB_ACT_NONE 0
#define AIMB_ACT_QUICKGELU 1 /* u*sigmoid(1.702u)   vit_clip.py:80-82 */
#define AIMB_ACT_GELU 2      /* exact erf GELU      vit_clip.py:56 (nn.GELU) */

#define AIMB_IMPL_AUTO 0 /* bf16 -> tensor-core kernels, f32 -> SIMT kernels */
#define AIMB_IMPL_SIMT 1 /* force the SIMT kernel (debug / cross-check) */
#define AIMB_IMPL_MMA 2  /* attention only: force the mma.sync kernel instead of the tcgen05 one (cross-check) */

/* GEMM epilogue, applied per output element (m, n) of the fp32 accumulator `acc`:
 *   v = acc + bias[n] * (bias_rowscaled ? row_scale[m % row_mod] : 1)
 *   if (out_pre)  out_pre[m,n] = v                      (saved pre-activation for backward)
 *   v = act(v)
 *   if (dact_src) v *= dact'(dact_src[m,n])             (backward through an activation)
 *   v *= alpha
 *   if (row_scale && !bias_rowscaled) v *= row_scale[m % row_mod]   (DropPath per-token multiplier)
 *   v += res1[m,n] + res2[m,n]
 *   out[m,n] = v          (or out[m,n] += v when accumulate && out_f32)
 *   if (colsum_out) colsum_out[n] (+)= sum_m v            (fp32; the bias gradient of the layer that consumes `out`,
 *                                                         produced while the tile is still in registers; `+=` when
 *                                                         colsum_accumulate: the caller has zeroed the buffer)
 * NULL pointers disable the corresponding term. */
typedef struct aimb_epilogue {
    const void* bias;       /* [N], dtype */
    const float* row_scale; /* [row_mod] fp32 */
    const void* res1;       /* [M, ldo] dtype */
    const void* res2;       /* [M, ldo] dtype */
    const void* dact_src;   /* [M, ldo] dtype */
    void* out;              /* [M, ldo] dtype (fp32 if out_f32) */
    void* out_pre;          /* [M, ldo] dtype */
    float alpha;
    int32_t row_mod;
    int32_t act;
    int32_t dact;
    int32_t bias_rowscaled;
    int32_t out_f32;
    int32_t accumulate;
    int64_t ldo; /* 0 -> N */
    float* colsum_out; /* [N] fp32, overwritten (or accumulated into, see colsum_accumulate) */
    int32_t colsum_accumulate; /* 1: do not zero colsum_out first (one memset of the whole gradient buffer by the caller) */
    /* LayerNorm folded into the GEMM that consumes it (vit_clip.py:71-77 feeding :132-138; the LayerNorms on this path are
     * frozen): with A = the UN-normalised rows x, W' = W * gamma (per input column), ln_wsum[n] = sum_k W'[n,k] and
     * bias[n] = b[n] + sum_k beta[k] W[n,k], the accumulator becomes  acc = ln_rstd[m] * (acc - ln_mean[m] * ln_wsum[n])
     * before the steps above: LN(x) W^T + b without ever writing LN(x).  All three NULL (default) or all three set. */
    const float* ln_mean;  /* [M] fp32 row means of A */
    const float* ln_rstd;  /* [M] fp32 1/sqrt(var + eps) */
    const float* ln_wsum;  /* [N] fp32 row sums of the gamma-scaled weight, taken from the values the GEMM reads */
} aimb_epilogue_t;

int aimb_version(void);
/* 0 if `device` is an sm_100 part this library can run on. */
int aimb_device_ok(int device);
const char* aimb_last_error(void);

/* ---- stem: vit_clip.py:433-447 / vitclip_aim.py:445-459 ------------------------------------ */
/* `(b t)` flatten + im2col of the k=s=p patch conv (replaces rearrange + cuDNN conv1 + reshape/permute).
 * x [B,3,T,H,W] (x_dtype f32 / bf16 / u8; u8 applies (v-mean[c])/std[c] = GPUNormalize,
 * mmaction/utils/module_hooks.py:35-87) -> cols [B*T*G*G, kpad] in `dtype`,
 * row (b*T+t)*G*G + gy*G + gx, column c*p*p + ky*p + kx, zero padded up to kpad. */
int aimb_im2col(const void* x, int32_t x_dtype, const float* mean, const float* std_, void* cols, int32_t dtype,
                int32_t B, int32_t T, int32_t H, int32_t W, int32_t patch, int32_t kpad, void* stream);
/* cls prepend + positional + temporal embedding + ln_pre in one pass (vit_clip.py:439-447).
 * tok [B*T*G*G, D]; z (pre-LN sum, optional, kept for backward) and x (ln_pre output) are [B*T*n, D]. */
int aimb_stem_assemble_ln(const void* tok, const void* cls, const void* pos, const void* temb, const void* gamma,
                          const void* beta, void* z, void* x, float* mean, float* rstd, int32_t B, int32_t T,
                          int32_t n, int32_t D, float eps, int32_t dtype, void* stream);
/* grad of temporal_embedding: out[t, d] (+)= sum_{b, token} dz[(b*T+t)*n + token, d]  (fp32 out). */
int aimb_temb_grad(const void* dz, float* out, int32_t B, int32_t T, int32_t n, int32_t D, int32_t dtype,
                   void* stream);

/* ---- LayerNorm: vit_clip.py:71-77 (fp32 statistics, eps 1e-5) ------------------------------ */
/* y may be NULL: statistics only (mean / rstd for a GEMM with the LayerNorm folded in, and for the LayerNorm backward). */
int aimb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                       int64_t rows, int32_t D, float eps, int32_t dtype, void* stream);
/* dx = dres + LN'(dy) ; dres may be NULL; dx may alias dres or dy. */
int aimb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                       const void* dres, void* dx, int64_t rows, int32_t D, int32_t dtype, void* stream);

/* Same, and additionally cs_out[c] = cs_alpha * sum_r dx[r,c] * (cs_row_scale ? cs_row_scale[r % cs_row_mod] : 1)
 * (fp32 [D], overwritten): the bias gradient of the adapter whose output gradient is `dx`, fused into the pass
 * that produces `dx` instead of re-reading it. */
int aimb_layernorm_bwd_colsum(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                              const void* dres, void* dx, const float* cs_row_scale, int32_t cs_row_mod, float cs_alpha,
                              float* cs_out, int64_t rows, int32_t D, int32_t dtype, void* stream);

/* ---- tail: vit_clip.py:450-456 (ln_post on the cls rows only, '(b t) d -> b d t') ----------- */
/* feat is fp32 [B, D, T]. mean/rstd [B*T] saved for backward. */
int aimb_tail_fwd(const void* x, const void* gamma, const void* beta, float* feat, float* mean, float* rstd,
                  int32_t B, int32_t T, int32_t n, int32_t D, float eps, int32_t dtype, void* stream);
/* dfeat fp32 [B, D, T] -> dx [B*T*n, D] (cls rows written, all other rows zeroed); dgamma/dbeta fp32 [D]
 * (overwritten). */
int aimb_tail_bwd(const float* dfeat, const void* x, const float* mean, const float* rstd, const void* gamma,
                  void* dx, float* dgamma, float* dbeta, int32_t B, int32_t T, int32_t n, int32_t D, int32_t dtype,
                  void* stream);

/* ---- GEMMs: every nn.Linear on the path (vit_clip.py:60-69, 93-97, 132-138, 157) ------------- */
/* C = epilogue(A[M,K] * W[N,K]^T).  A row stride lda, W row stride ldw (elements), K contiguous.
 * bf16 + AIMB_IMPL_AUTO runs the TMA + tcgen05/TMEM kernel (needs K % 64 == 0, N % 64 == 0,
 * 16-byte aligned rows); f32 or AIMB_IMPL_SIMT runs the SIMT kernel. */
int aimb_gemm_nt(const void* A, int64_t lda, const void* W, int64_t ldw, const aimb_epilogue_t* epi, int64_t M,
                 int32_t N, int32_t K, int32_t dtype, int32_t impl, void* stream);
/* Fully strided SIMT GEMM: C[m,n] = epilogue(sum_k A[m*a_sm + k*a_sk] * B[n*b_sn + k*b_sk]). */
int aimb_gemm_strided(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sn, int64_t b_sk,
                      const aimb_epilogue_t* epi, int64_t M, int32_t N, int32_t K, int32_t dtype, void* stream);
/* Adapter weight gradient (autograd of vit_clip.py:62,64): dW[N,K] (fp32) (+)= alpha * dY[R,N]^T * X[R,K]. */
int aimb_gemm_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t R, int32_t N,
                    int32_t K, float alpha, int32_t accumulate, int32_t dtype, int32_t impl, void* stream);
/* Adapter bottleneck (vit_clip.py:51-69) as ONE kernel: H = epi1(A[M,D] * W1[R,D]^T) is written to epi1->out (and
 * epi1->out_pre) and, without a round trip through HBM, used as the left operand of epi2(H * W2[D,R]^T) -> epi2->out.
 *   forward : epi1 = {bias b1, act GELU, row_scale, out g', out_pre h}, epi2 = {bias b2 (row-scaled), alpha, res1[, res2]}
 *   backward: A = dy, W1 := W2^T [R,D], W2 := W1^T [D,R]; epi1 = {dact_src h, dact GELU, alpha, row_scale, out d_h,
 *             colsum_out db1}, epi2 = {res1}.
 * bf16 only; M >= 128, R in {64, 192, 256}, D a multiple of 192 or 256 (returns AIMB_ERR_UNSUPPORTED otherwise: the
 * caller then issues the two aimb_gemm_nt calls this kernel fuses). */
int aimb_adapter_fused(const void* A, int64_t lda, const void* W1, const void* W2, const aimb_epilogue_t* epi1,
                       const aimb_epilogue_t* epi2, int64_t M, int32_t D, int32_t R, int32_t dtype, void* stream);
/* Two nn.Linear of one block in ONE tcgen05 launch (bf16 only).  The MLP adapter reads the same ln_2(x) as mlp.c_fc and
 * adds into the same sum as mlp.c_proj (vitclip_aim.py:210-211 == vit_clip.py:285-286), so its two small GEMMs ride on the
 * frozen ones, forward and backward:
 *   AIMB_DUAL_NCAT  [C1 | C2] = [epi1(A1 W1^T) | epi2(A1 W2^T)]    A1 [M,K1], W1 [N1,K1], W2 [N2,K1]; A2 / K2 ignored;
 *                   each epilogue has its own outputs, leading dimension, activation (c_fc | D_fc1;  d_hf | d_h)
 *   AIMB_DUAL_KCAT  C = epi1(A1 W1^T + A2 W2^T  [+ bias2[n] * bias2_scale * bias2_row_scale[m % bias2_row_mod]])
 *                   A1 [M,K1], W1 [N1,K1], A2 [M,K2], W2 [N1,K2]; epi2 / N2 ignored (c_proj + D_fc2;  c_fc^T + D_fc1^T)
 * No colsum_out / out_f32.  Returns AIMB_ERR_UNSUPPORTED for shapes outside N1 % 256 == 0, N2 in {192, 256} (NCAT) or
 * N1 % 192 == 0 / N1 % 256 == 0 (KCAT), K % 64 == 0, M >= 128, 32-byte aligned outputs with ldo % 16 == 0: the caller then
 * issues the separate aimb_gemm_nt calls. */
#define AIMB_DUAL_NCAT 0
#define AIMB_DUAL_KCAT 1
int aimb_gemm_dual(int32_t mode, const void* A1, int64_t lda1, const void* W1, int64_t ldw1, const void* A2, int64_t lda2,
                   const void* W2, int64_t ldw2, const aimb_epilogue_t* epi1, const aimb_epilogue_t* epi2, const void* bias2,
                   const float* bias2_row_scale, int32_t bias2_row_mod, float bias2_scale, int64_t M, int32_t N1, int32_t N2,
                   int32_t K1, int32_t K2, int32_t dtype, void* stream);
/* Bias gradient: out[c] (+)= alpha * sum_r x[r, c] * (row_scale ? row_scale[r % row_mod] : 1)  (fp32 out). */
int aimb_colsum(const void* x, int64_t ld, const float* row_scale, int32_t row_mod, float alpha, float* out,
                int64_t R, int32_t C, int32_t accumulate, int32_t dtype, void* stream);
int aimb_transpose(const void* in, void* out, int32_t R, int32_t C, int32_t dtype, void* stream);
/* One AdamW step (torch.optim.AdamW semantics: p *= 1 - lr*wd where wd_mask[i] != 0 (NULL: everywhere); m, v moments; bias
 * correction with the DEVICE scalar step[0] >= 1) over the flat fp32 buffer that holds every trainable tensor of the path and
 * the flat gradient buffer the backward fills — replaces the optimizer's multi-tensor launch chain
 * (configs/recognition/vit/vitclip_base_k400.py `optimizer = dict(type='AdamW', ...)`, mmcv build_optimizer). */
int aimb_adamw_flat(float* p, const float* g, float* m, float* v, const uint8_t* wd_mask, const float* step, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t n, void* stream);

/* One launch for many transposes (the per-step transposes of the trainable adapter weights for dgrad):
 * matrix b = src + table[3b] (elements), shape [table[3b+1], table[3b+2]], written transposed at dst + table[3b].
 * `table` is a DEVICE array of 3*nmat int64. */
int aimb_transpose_batched(const void* src, void* dst, const int64_t* table, int32_t nmat, int32_t dtype, void* stream);

/* ---- attention cores: vit_clip.py:140-156 (the part between the QKV and out_proj GEMMs) ------ */
/* Spatial: one softmax(q k^T / 8) v problem per (frame, head); n tokens, head_dim 64.
 * qkv [frames*n, 3D], o [frames*n, D], lse fp32 [frames, heads, n] (may be NULL in inference). */
int aimb_attn_spatial_fwd(const void* qkv, void* o, float* lse, int32_t frames, int32_t n, int32_t heads,
                          int32_t dtype, int32_t impl, void* stream);
int aimb_attn_spatial_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv,
                          int32_t frames, int32_t n, int32_t heads, int32_t dtype, int32_t impl, void* stream);
/* Temporal (T-Adapter attention, vitclip_aim.py:200-206): one problem per (clip b, token, head) over the
 * T frames; rows of the sequence are (b*T + t)*n + token, i.e. the `n (b t) d -> t (b n) d` view is read
 * in place through the row stride n*3D — no transpose kernel. */
int aimb_attn_temporal_fwd(const void* qkv, void* o, int32_t B, int32_t T, int32_t n, int32_t heads, int32_t dtype,
                           void* stream);
int aimb_attn_temporal_bwd(const void* qkv, const void* d_o, void* d_qkv, int32_t B, int32_t T, int32_t n,
                           int32_t heads, int32_t dtype, void* stream);

/* ---- fork block extras (vit_clip.py:147-151, 182-186, 264-275) ------------------------------ */
/* w_o[f] = sum_{i,j} exp(sum_h q_{f,h,i}.k_{f,h,j} / 8)  and  w_c[f] = sum_i exp(q_{f,i}.kc_f / 8)
 * (full-width D dot products; fp32, no max subtraction, as the reference).  kc [frames, D]. */
int aimb_fork_weights(const void* qkv, const void* kc, float* w_o, float* w_c, int32_t frames, int32_t n,
                      int32_t D, int32_t dtype, void* stream);

/* out[f,i,:] = x[f,i,:] + (1 - lam[f]) * a_o[f,i,:] + rs[i] * s_frame[f,:]   (vit_clip.py:275; s_frame = scale *
 * S_Adapter(lam * a_c) is constant over the tokens of a frame because the cross attention has a single key). */
int aimb_fork_combine(const void* x, const void* a_o, const void* s_frame, const float* lam, const float* rs, void* out,
                      int32_t BT, int32_t n, int32_t D, int32_t dtype, void* stream);
/* d_ao[f,i,:] = (1 - lam[f]) * dx[f,i,:] ;  d_s[f,:] = sum_i rs[i] * dx[f,i,:] */
int aimb_fork_combine_bwd(const void* dx, const float* lam, const float* rs, void* d_ao, void* d_s, int32_t BT, int32_t n,
                          int32_t D, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIMB200_H */

// Shared by the tcgen05 GEMM kernels (gemm_tc.cu: gemm_tc4_kernel / generic / wgrad / one-kernel adapter; gemm_dual.cu: the
// paired GEMM): tile constants, the epilogue variants, the 8-column epilogue math in the TMEM-native layout and the slab
// helpers of the row-layout epilogue.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace aimb {

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 192;          // wgrad kernel: 1 producer + 1 MMA + 4 epilogue warps
constexpr int GEMM_THREADS = 320;        // GEMM: 1 producer + 1 MMA + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int STG_LD = 36;                          // row stride (floats) of the transpose tile: 16-byte aligned rows,
                                                    // conflict-free for the 128-bit stores and loads used below
constexpr int STG_WARP_FLOATS = 32 * STG_LD + 128;  // per-warp 32x32 fp32 transpose buffer (padded) + bias slice
constexpr int STG_BYTES = EPI_WARPS * STG_WARP_FLOATS * 4 + 1024;   // + scol[256]: per-CTA column-sum partials
constexpr int A_STAGE_BYTES = BM * BK * 2;

__device__ __forceinline__ void unpack8(const uint4& t, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}

// QuickGELU for the bf16 tensor-core epilogues: sigmoid(z) = 0.5 + 0.5 tanh(z / 2) with the single-instruction
// MUFU.TANH (max relative error 2^-11, an order below bf16 rounding) — 4 instructions per element instead of ~12
// for the ex2 / rcp form (ncu: the c_fc epilogues were ~45 % issue-bound).  fp32 parity mode never comes here.
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float quick_gelu_fast(float u) {
    const float h = 0.5f * u;
    return fmaf(h, tanh_fast(0.851f * u), h);
}
__device__ __forceinline__ float quick_gelu_grad_fast(float u) {   // s + 1.702 u s (1 - s), s = (1 + t) / 2
    const float t = tanh_fast(0.851f * u);
    return fmaf(0.4255f * u, fmaf(-t, t, 1.f), fmaf(0.5f, t, 0.5f));
}
template <int ACT> __device__ __forceinline__ float act_fn(float v, int rt) {
    if (ACT < 0) return apply_act(rt, v);
    if (ACT == AIMB_ACT_QUICKGELU) return quick_gelu_fast(v);
    if (ACT == AIMB_ACT_GELU) return gelu_erf(v);
    return v;
}
template <int ACT> __device__ __forceinline__ float act_grad_fn(float u, int rt) {
    if (ACT < 0) return apply_act_grad(rt, u);
    if (ACT == AIMB_ACT_QUICKGELU) return quick_gelu_grad_fast(u);
    if (ACT == AIMB_ACT_GELU) return gelu_erf_grad(u);
    return 1.f;
}

// Epilogue variants are compiled into SEPARATE kernels (template parameter V) so each gets its own register
// allocation; the host picks V from the C-ABI epilogue description.
//   0 plain/bias (QKV, out_proj, dgrads)   1 bias+QuickGELU (c_fc)        2 bias+GELU (adapter fc1)
//   3 +res1 (c_proj, fc2, d_a)             4 +res1+res2 (S_Adapter fc2)   5 x QuickGELU'(saved) (d_hf)
//   6 x GELU'(saved) (d_h)                 7 anything else the C ABI allows (generic, slower)
template <int V> struct EpiVariant;
template <> struct EpiVariant<0> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_NONE, EXT = 0; };
template <> struct EpiVariant<1> { static constexpr int ACT = AIMB_ACT_QUICKGELU, DACT = AIMB_ACT_NONE, EXT = 0; };
template <> struct EpiVariant<2> { static constexpr int ACT = AIMB_ACT_GELU, DACT = AIMB_ACT_NONE, EXT = 0; };
template <> struct EpiVariant<3> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_NONE, EXT = 2; };
template <> struct EpiVariant<4> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_NONE, EXT = 6; };
template <> struct EpiVariant<5> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_QUICKGELU, EXT = 1; };
template <> struct EpiVariant<6> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_GELU, EXT = 1; };
template <> struct EpiVariant<7> { static constexpr int ACT = -1, DACT = -1, EXT = 7; };
template <> struct EpiVariant<8> { static constexpr int ACT = AIMB_ACT_NONE, DACT = AIMB_ACT_NONE, EXT = 0; };   // 0 + folded LayerNorm (QKV)

inline int pick_variant(const EpiParams& e) {
    const int ext = (e.dact_src ? 1 : 0) | (e.res1 ? 2 : 0) | (e.res2 ? 4 : 0);
    const int act = e.act, dact = e.dact_src ? e.dact : AIMB_ACT_NONE;
    if (e.ln_mean) return (ext == 0 && act == AIMB_ACT_NONE && !e.out_pre) ? 8 : 7;
    if (ext == 0 && act == AIMB_ACT_NONE) return 0;
    if (ext == 0 && act == AIMB_ACT_QUICKGELU) return 1;
    if (ext == 0 && act == AIMB_ACT_GELU) return 2;
    if (ext == 2 && act == AIMB_ACT_NONE) return 3;
    if (ext == 6 && act == AIMB_ACT_NONE) return 4;
    if (ext == 1 && act == AIMB_ACT_NONE && dact == AIMB_ACT_QUICKGELU) return 5;
    if (ext == 1 && act == AIMB_ACT_NONE && dact == AIMB_ACT_GELU) return 6;
    return 7;
}

template <int V, bool DIRECT = false> struct EpiBufs {
    static constexpr int EXT = EpiVariant<V>::EXT;
    static constexpr bool PRE = (V == 1 || V == 2);
    static constexpr int NEXT = ((EXT & 1) ? 1 : 0) + ((EXT & 2) ? 1 : 0) + ((EXT & 4) ? 1 : 0);
    static constexpr int NBUF = DIRECT ? NEXT : (NEXT > 0 ? NEXT : 1) + (PRE ? 1 : 0);   // slab path: out aliases the first operand buffer
    static constexpr int BIAS_BYTES = (V == 8) ? 512 : 256;                  // 64 fp32 bias values (+ 64 LayerNorm weight sums)
    static constexpr int WARP_BYTES = NBUF * 4096 + BIAS_BYTES;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint32_t slab_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// one warp instruction = 4 full 128-byte rows of the slab
__device__ __forceinline__ void slab_fetch(uint32_t buf, const bf16* g, int64_t ldo, int64_t row_base, int n_base, int M, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), ch = lane & 7;
        if (row_base + row < M) cp_async16(buf + slab_off(row, ch), g + (row_base + row) * ldo + n_base + ch * 8);
    }
}

// 8 columns of one row: v = accumulators in, packed bf16 result out (and the packed pre-activation when asked)
template <int ACT, int DACT, int EXT>
__device__ __forceinline__ uint4 epi_math8(const EpiParams& e, float rs, float* v, const float* bias8, const uint4& xd,
                                           const uint4& x1, const uint4& x2, uint4& pre_pk, bool want_pre) {
    float t[8];
    if (e.bias) {
        const float bs = e.bias_rowscaled ? rs : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(bias8[j], bs, v[j]);
    }
    if (want_pre) {
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pre_pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pw[j] = *reinterpret_cast<uint32_t*>(&h2);
            v[2 * j] = __uint_as_float(pw[j] << 16);
            v[2 * j + 1] = __uint_as_float(pw[j] & 0xffff0000u);
        }
    }
    if (ACT != AIMB_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_fn<ACT>(v[j], e.act);
    }
    if ((EXT & 1) && DACT != AIMB_ACT_NONE) {
        unpack8(xd, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= act_grad_fn<DACT>(t[j], e.dact);
    }
    const float sc = e.alpha * ((e.row_scale && !e.bias_rowscaled) ? rs : 1.f);
    if (sc != 1.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sc;
    }
    if (EXT & 2) {
        unpack8(x1, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if (EXT & 4) {
        unpack8(x2, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        ow[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    return o;
}


// 2-D bf16 row-major [rows, cols] tensor map (row stride ld elements); box = 64 columns x box_rows rows, 128B swizzle; cached.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows);

}  // namespace aimb

// Shared device/host helpers for the aimb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/aimb200.h"

#define AIMB_CHECK_LAUNCH()                                   \
    do {                                                      \
        cudaError_t e__ = cudaPeekAtLastError();              \
        if (e__ != cudaSuccess) return AIMB_ERR_CUDA;         \
    } while (0)

namespace aimb {

typedef __nv_bfloat16 bf16;

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerialization allowed: the next kernel's CTAs may be
// scheduled (and run their prologue) while the previous kernel drains.  Protocol, identical in every kernel:
//   pdl_trigger()  - as early as possible: lets the dependent grid start launching once all CTAs have started
//   pdl_wait()     - before the FIRST access to global memory: blocks until the previous grid has fully completed and
//                    its writes are visible.  Safe transitively (a grid cannot complete before its own wait returns).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() { pdl_trigger(); pdl_wait(); }

extern int g_pdl_enabled;   // host side (api.cu); 0 disables the launch attribute (debug)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device): a function attribute belongs to
// the device's context, so a process that drives several GPUs must set it on each of them.  (`kernel` may contain commas.)
#define AIMB_SET_SMEM_ATTR(bytes, ...)                                                                                  \
    do {                                                                                                               \
        static bool done__[64] = {false};                                                                              \
        int dev__ = 0;                                                                                                 \
        cudaGetDevice(&dev__);                                                                                         \
        if (dev__ < 0 || dev__ >= 64) return AIMB_ERR_UNSUPPORTED;                                                     \
        if (!done__[dev__]) {                                                                                          \
            if (cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)) != cudaSuccess) \
                return AIMB_ERR_CUDA;                                                                                  \
            done__[dev__] = true;                                                                                      \
        }                                                                                                              \
    } while (0)

inline int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

template <typename T> struct DT;
template <> struct DT<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct DT<bf16> {
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return DT<T>::ld(p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { DT<T>::st(p, v); }

// Round a float to the storage type and back (used so that "what is stored" == "what is used next").
template <typename T> __device__ __forceinline__ float roundT(float v);
template <> __device__ __forceinline__ float roundT<float>(float v) { return v; }
template <> __device__ __forceinline__ float roundT<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- activations (reference: QuickGELU vit_clip.py:80-82; nn.GELU exact erf vit_clip.py:56) ----
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float quick_gelu(float u) { return u * sigmoidf_(1.702f * u); }
__device__ __forceinline__ float quick_gelu_grad(float u) {
    float s = sigmoidf_(1.702f * u);
    return s * (1.f + 1.702f * u * (1.f - s));
}
__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float u) {
    float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
    float pdf = 0.39894228040143268f * __expf(-0.5f * u * u);
    return cdf + u * pdf;
}
__device__ __forceinline__ float apply_act(int act, float v) {
    if (act == AIMB_ACT_QUICKGELU) return quick_gelu(v);
    if (act == AIMB_ACT_GELU) return gelu_erf(v);
    return v;
}
__device__ __forceinline__ float apply_act_grad(int act, float u) {
    if (act == AIMB_ACT_QUICKGELU) return quick_gelu_grad(u);
    if (act == AIMB_ACT_GELU) return gelu_erf_grad(u);
    return 1.f;
}

// Epilogue shared by the SIMT and the tcgen05 GEMMs.  See include/aimb200.h (aimb_epilogue_t).
// v = acc (+ bias[n] * (bias_rowscaled ? rs[m % row_mod] : 1)); [store pre]; v = act(v);
// [v *= act'(dact_src[m,n])]; v *= alpha; [v *= rs[m % row_mod] unless bias_rowscaled]; v += res1 + res2; store.
struct EpiParams {
    const void* bias;
    const float* row_scale;
    const void* res1;
    const void* res2;
    const void* dact_src;
    void* out;
    void* out_pre;
    float alpha;
    int32_t row_mod;
    int32_t act;
    int32_t dact;
    int32_t bias_rowscaled;
    int32_t out_f32;      // out is float regardless of T (wgrad accumulators)
    int32_t accumulate;   // out += v (only with out_f32)
    int64_t ldo;          // leading dim of out/out_pre/res1/res2/dact_src (elements)
    float* colsum_out;    // [N] fp32: column sums of the stored values (atomically accumulated; zeroed by the host side)
    int32_t colsum_accumulate;   // host side only: skip that zeroing
    const float* ln_mean;        // folded LayerNorm (see aimb200.h): acc = ln_rstd[m] * (acc - ln_mean[m] * ln_wsum[n])
    const float* ln_rstd;
    const float* ln_wsum;
};

inline EpiParams make_epi(const aimb_epilogue_t* e, int64_t ld_default) {
    EpiParams p;
    p.bias = e->bias; p.row_scale = e->row_scale; p.res1 = e->res1; p.res2 = e->res2;
    p.dact_src = e->dact_src; p.out = e->out; p.out_pre = e->out_pre; p.alpha = e->alpha;
    p.row_mod = e->row_mod > 0 ? e->row_mod : 1; p.act = e->act; p.dact = e->dact;
    p.bias_rowscaled = e->bias_rowscaled; p.out_f32 = e->out_f32; p.accumulate = e->accumulate;
    p.ldo = e->ldo > 0 ? e->ldo : ld_default;
    p.colsum_out = e->colsum_out;
    p.colsum_accumulate = e->colsum_accumulate;
    const bool ln = e->ln_mean && e->ln_rstd && e->ln_wsum;
    p.ln_mean = ln ? e->ln_mean : nullptr; p.ln_rstd = ln ? e->ln_rstd : nullptr; p.ln_wsum = ln ? e->ln_wsum : nullptr;
    return p;
}

}  // namespace aimb

// Element-wise GEMM epilogue shared by the SIMT and tcgen05 kernels (semantics: include/aimb200.h).
#pragma once
#include "common.cuh"

namespace aimb {

// Applies the epilogue to one accumulator value and stores it.  T = storage dtype.
template <typename T>
__device__ __forceinline__ void epilogue_store(const EpiParams& e, int64_t m, int n, float acc) {
    float rs = 1.f;
    if (e.row_scale) rs = e.row_scale[m % e.row_mod];
    float v = acc;
    if (e.ln_mean) v = (v - e.ln_mean[m] * e.ln_wsum[n]) * e.ln_rstd[m];
    if (e.bias) v += ldf<T>((const T*)e.bias + n) * (e.bias_rowscaled ? rs : 1.f);
    int64_t off = m * e.ldo + n;
    if (e.out_pre) { stf<T>((T*)e.out_pre + off, v); v = roundT<T>(v); }
    v = apply_act(e.act, v);
    if (e.dact_src) v *= apply_act_grad(e.dact, ldf<T>((const T*)e.dact_src + off));
    v *= e.alpha;
    if (e.row_scale && !e.bias_rowscaled) v *= rs;
    if (e.res1) v += ldf<T>((const T*)e.res1 + off);
    if (e.res2) v += ldf<T>((const T*)e.res2 + off);
    if (e.colsum_out) atomicAdd(e.colsum_out + n, v);
    if (e.out_f32) {
        float* o = (float*)e.out + off;
        *o = e.accumulate ? *o + v : v;
    } else {
        stf<T>((T*)e.out + off, v);
    }
}

}  // namespace aimb

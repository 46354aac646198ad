// HBM-bound kernels of the path: stem assembly, LayerNorm fwd/bwd, tail, bias/temporal-embedding
// reductions, transpose.  One warp per row, 16-byte vector loads, fp32 statistics.
#include "common.cuh"

namespace aimb {

template <typename T> struct VecIO;
template <> struct VecIO<float> {
    static constexpr int N = 4;
    typedef float4 Raw;
    static __device__ __forceinline__ Raw ldraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void unpack(const Raw& t, float* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void ld(const float* p, float* v) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecIO<bf16> {
    static constexpr int N = 8;
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ldraw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ void unpack(const Raw& t, float* v) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void ld(const bf16* p, float* v) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void st(bf16* p, const float* v) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};

// Rows up to MAXD elements are kept in registers: NIT chunks of (32 lanes * VEC) elements.
constexpr int MAXD = 1024;

template <typename T> struct RowRegs {
    static constexpr int V = VecIO<T>::N;
    static constexpr int NIT = MAXD / (32 * V);
    float v[NIT][V];
    __device__ __forceinline__ void load(const T* row, int D, int lane) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = (it * 32 + lane) * V;
            if (c < D) VecIO<T>::ld(row + c, v[it]);
            else {
#pragma unroll
                for (int j = 0; j < V; ++j) v[it][j] = 0.f;
            }
        }
    }
};

// A row held as raw 16-byte vectors (half the registers of RowRegs for bf16): used to keep the NEXT row's loads
// in flight while the current row is being processed.
template <typename T, int NIT> struct RowRaw {
    static constexpr int V = VecIO<T>::N;
    typename VecIO<T>::Raw v[NIT];
    __device__ __forceinline__ void load(const T* row, int D, int lane) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = (it * 32 + lane) * V;
            if (c < D) v[it] = VecIO<T>::ldraw(row + c);
        }
    }
};

// ------------------------------------------------------------------ LayerNorm forward
template <typename T>
__device__ __forceinline__ void ln_row(RowRegs<T>& r, int D, int lane, float eps, float& mean, float& rstd) {
    constexpr int V = RowRegs<T>::V, NIT = RowRegs<T>::NIT;
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
        for (int j = 0; j < V; ++j) s += r.v[it][j];
    mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
#pragma unroll
            for (int j = 0; j < V; ++j) { float d = r.v[it][j] - mean; q += d * d; }
        }
    }
    rstd = rsqrtf(warp_sum(q) / D + eps);
}

template <typename T>
__device__ __forceinline__ void ln_apply_store(const RowRegs<T>& r, const T* gamma, const T* beta, T* y, int D, int lane,
                                               float mean, float rstd) {
    constexpr int V = RowRegs<T>::V, NIT = RowRegs<T>::NIT;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
            float g[V], b[V], o[V];
            VecIO<T>::ld(gamma + c, g);
            VecIO<T>::ld(beta + c, b);
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = (r.v[it][j] - mean) * rstd * g[j] + b[j];
            VecIO<T>::st(y + c, o);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                            const T* __restrict__ beta, T* __restrict__ y,
                                                            float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                            int64_t rows, int D, float eps) {
    pdl_grid_sync();
    int lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    RowRegs<T> r;
    r.load(x + row * D, D, lane);
    float mean, rstd;
    ln_row<T>(r, D, lane, eps, mean, rstd);
    if (y) ln_apply_store<T>(r, gamma, beta, y + row * D, D, lane, mean, rstd);     // y == nullptr: statistics only
    if (lane == 0) {
        if (mean_o) mean_o[row] = mean;
        if (rstd_o) rstd_o[row] = rstd;
    }
}

// ------------------------------------------------------------------ LayerNorm backward (input grad only)
// One warp per row, RPW rows per warp.  Optionally accumulates cs_out[c] += alpha * w[row] * dx[row, c] (the bias
// gradient of the adapter fed by dx) from the values already in registers: per-lane column partials over the
// warp's rows, reduced across the 4 warps through smem, one atomicAdd per column per block.
template <typename T, int RPW, int NIT>      // NIT = ceil(D / (32 lanes * 16-byte vector)): registers sized to the row length
__global__ void __launch_bounds__(128, 3) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ mean_i,
                                                            const float* __restrict__ rstd_i,
                                                            const T* __restrict__ gamma, const T* dres, T* dx,
                                                            const float* __restrict__ cs_w, int cs_mod, float cs_alpha,
                                                            float* __restrict__ cs_out, int64_t rows, int D) {
    pdl_grid_sync();
    constexpr int V = VecIO<T>::N;
    constexpr int NWARP = 4;
    __shared__ float red[RPW > 1 ? NWARP * MAXD : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float cs[NIT][V];
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
        for (int j = 0; j < V; ++j) cs[it][j] = 0.f;
    // each warp handles RPW consecutive rows; with the fused column sums a block flushes one atomic per column
    // for its 4*RPW rows.  The rows are software-pipelined: the raw 16-byte loads (dy, x, dres) of row k+1 are issued
    // before row k is reduced, so one memory latency per row is hidden behind the previous row's math.
    const int64_t row_first = ((int64_t)blockIdx.x * NWARP + warp) * RPW;
    RowRaw<T, NIT> n_dy, n_x, n_res;
    if (row_first < rows) {
        n_dy.load(dy + row_first * D, D, lane);
        n_x.load(x + row_first * D, D, lane);
        if (dres) n_res.load(dres + row_first * D, D, lane);
    }
#pragma unroll 1
    for (int k = 0; k < RPW; ++k) {
        const int64_t row = row_first + k;
        if (row >= rows) break;
        RowRaw<T, NIT> c_dy = n_dy, c_x = n_x, c_res = n_res;
        if (k + 1 < RPW && row + 1 < rows) {
            n_dy.load(dy + (row + 1) * D, D, lane);
            n_x.load(x + (row + 1) * D, D, lane);
            if (dres) n_res.load(dres + (row + 1) * D, D, lane);
        }
        const float mean = mean_i[row], rstd = rstd_i[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = (it * 32 + lane) * V;
            if (c < D) {
                float gm[V], a[V], b[V];
                VecIO<T>::ld(gamma + c, gm);
                VecIO<T>::unpack(c_dy.v[it], a);
                VecIO<T>::unpack(c_x.v[it], b);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float gg = a[j] * gm[j];
                    s1 += gg;
                    s2 += gg * ((b[j] - mean) * rstd);
                }
            }
        }
        s1 = warp_sum(s1) / D;
        s2 = warp_sum(s2) / D;
        const float w = (RPW > 1 && cs_w) ? cs_w[row % cs_mod] : 1.f;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {       // second sweep re-derives g and xhat from the raw registers (cheap ALU)
            int c = (it * 32 + lane) * V;        // instead of keeping two more fp32 copies of the row alive
            if (c < D) {
                float gm[V], a[V], b[V], o[V];
                VecIO<T>::ld(gamma + c, gm);
                VecIO<T>::unpack(c_dy.v[it], a);
                VecIO<T>::unpack(c_x.v[it], b);
                if (dres) VecIO<T>::unpack(c_res.v[it], o);
                else {
#pragma unroll
                    for (int j = 0; j < V; ++j) o[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    o[j] += rstd * (a[j] * gm[j] - s1 - (b[j] - mean) * rstd * s2);
                    if (RPW > 1) cs[it][j] = fmaf(w, o[j], cs[it][j]);
                }
                VecIO<T>::st(dx + row * D + c, o);
            }
        }
    }
    if (RPW > 1) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int c = (it * 32 + lane) * V;
            if (c < D) {
#pragma unroll
                for (int j = 0; j < V; ++j) red[warp * MAXD + c + j] = cs[it][j];
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += 128) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) t += red[w * MAXD + c];
            atomicAdd(cs_out + c, cs_alpha * t);
        }
    }
}

// ------------------------------------------------------------------ stem: cls | tok + pos + temb -> z ; ln_pre -> x
template <typename T>
__global__ void __launch_bounds__(128) stem_assemble_ln_kernel(const T* __restrict__ tok, const T* __restrict__ cls,
                                                               const T* __restrict__ pos, const T* __restrict__ temb,
                                                               const T* __restrict__ gamma, const T* __restrict__ beta,
                                                               T* __restrict__ z, T* __restrict__ xo,
                                                               float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                               int BT, int T_, int n, int D, float eps) {
    pdl_grid_sync();
    constexpr int V = RowRegs<T>::V, NIT = RowRegs<T>::NIT;
    int lane = threadIdx.x & 31;
    int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= (int64_t)BT * n) return;
    int f = (int)(row / n), tk = (int)(row % n), t = f % T_;
    const T* src = (tk == 0) ? cls : tok + ((int64_t)f * (n - 1) + (tk - 1)) * D;
    RowRegs<T> r;
    r.load(src, D, lane);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
            float a[V], b[V];
            VecIO<T>::ld(pos + (int64_t)tk * D + c, a);
            VecIO<T>::ld(temb + (int64_t)t * D + c, b);
            // same association order as the reference: (x + pos) + temb   (vitclip_aim.py:452,456)
#pragma unroll
            for (int j = 0; j < V; ++j) r.v[it][j] = roundT<T>(roundT<T>(r.v[it][j] + a[j]) + b[j]);
            if (z) VecIO<T>::st(z + row * D + c, r.v[it]);
        }
    }
    float mean, rstd;
    ln_row<T>(r, D, lane, eps, mean, rstd);
    ln_apply_store<T>(r, gamma, beta, xo + row * D, D, lane, mean, rstd);
    if (lane == 0) {
        if (mean_o) mean_o[row] = mean;
        if (rstd_o) rstd_o[row] = rstd;
    }
}

// ------------------------------------------------------------------ im2col of the patch conv
template <typename TI> __device__ __forceinline__ float ld_in(const TI* p) { return (float)*p; }
template <> __device__ __forceinline__ float ld_in<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename TI, typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const TI* __restrict__ x, const float* __restrict__ mean,
                                                     const float* __restrict__ std_, T* __restrict__ cols, int B,
                                                     int T_, int H, int W, int p, int kpad) {
    pdl_grid_sync();
    int G = W / p, Gy = H / p;
    int64_t total = (int64_t)B * T_ * 3 * Gy * p * G;  // one thread per (b, t, c, gy, ky, gx): p contiguous pixels
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    int gx = (int)(id % G); id /= G;
    int ky = (int)(id % p); id /= p;
    int gy = (int)(id % Gy); id /= Gy;
    int c = (int)(id % 3); id /= 3;
    int t = (int)(id % T_);
    int b = (int)(id / T_);
    const TI* src = x + ((((int64_t)b * 3 + c) * T_ + t) * H + (gy * p + ky)) * W + gx * p;
    int64_t row = ((int64_t)(b * T_ + t) * Gy + gy) * G + gx;
    T* dst = cols + row * kpad + (c * p + ky) * p;
    float m = 0.f, rs = 1.f;
    if (mean) { m = mean[c]; rs = 1.f / std_[c]; }
    for (int kx = 0; kx < p; ++kx) stf<T>(dst + kx, (ld_in<TI>(src + kx) - m) * rs);
    if (c == 2 && ky == p - 1)
        for (int k = 3 * p * p; k < kpad; ++k) stf<T>(cols + row * kpad + k, 0.f);
}

// ------------------------------------------------------------------ tail: ln_post on cls rows -> feat [B, D, T] fp32
template <typename T>
__global__ void __launch_bounds__(128) tail_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma,
                                                       const T* __restrict__ beta, float* __restrict__ feat,
                                                       float* __restrict__ mean_o, float* __restrict__ rstd_o, int BT,
                                                       int T_, int n, int D, float eps) {
    pdl_grid_sync();
    constexpr int V = RowRegs<T>::V, NIT = RowRegs<T>::NIT;
    int lane = threadIdx.x & 31;
    int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= BT) return;
    int b = f / T_, t = f % T_;
    RowRegs<T> r;
    r.load(x + (int64_t)f * n * D, D, lane);
    float mean, rstd;
    ln_row<T>(r, D, lane, eps, mean, rstd);
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
            float g[V], bb[V];
            VecIO<T>::ld(gamma + c, g);
            VecIO<T>::ld(beta + c, bb);
#pragma unroll
            for (int j = 0; j < V; ++j)
                feat[((int64_t)b * D + c + j) * T_ + t] = (r.v[it][j] - mean) * rstd * g[j] + bb[j];
        }
    }
    if (lane == 0) { mean_o[f] = mean; rstd_o[f] = rstd; }
}

template <typename T>
__global__ void __launch_bounds__(128) tail_bwd_kernel(const float* __restrict__ dfeat, const T* __restrict__ x,
                                                       const float* __restrict__ mean_i,
                                                       const float* __restrict__ rstd_i, const T* __restrict__ gamma,
                                                       T* __restrict__ dx, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta, int BT, int T_, int n, int D) {
    pdl_grid_sync();
    constexpr int V = RowRegs<T>::V, NIT = RowRegs<T>::NIT;
    int lane = threadIdx.x & 31;
    int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= BT) return;
    int b = f / T_, t = f % T_;
    RowRegs<T> xh, g;
    xh.load(x + (int64_t)f * n * D, D, lane);
    float mean = mean_i[f], rstd = rstd_i[f];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
            float gm[V];
            VecIO<T>::ld(gamma + c, gm);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float dy = dfeat[((int64_t)b * D + c + j) * T_ + t];
                float h = (xh.v[it][j] - mean) * rstd;
                atomicAdd(dgamma + c + j, dy * h);
                atomicAdd(dbeta + c + j, dy);
                float gg = dy * gm[j];
                g.v[it][j] = gg;
                xh.v[it][j] = h;
                s1 += gg;
                s2 += gg * h;
            }
        }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        int c = (it * 32 + lane) * V;
        if (c < D) {
            float o[V];
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] = rstd * (g.v[it][j] - s1 - xh.v[it][j] * s2);
            VecIO<T>::st(dx + (int64_t)f * n * D + c, o);
        }
    }
}

// ------------------------------------------------------------------ column sums (bias grads, temporal-embedding grad)
// block = 32 lanes (each a 16-byte column vector) x 32 row lanes; a block reduces `rows_per_block` (128) rows of 32
// column vectors with 4 independent loads in flight per thread, then one atomicAdd per column.  (More, smaller blocks
// were measured slower: every block adds into the same C addresses and the atomics serialise in L2.)
template <typename T>
__global__ void __launch_bounds__(1024) colsum_kernel(const T* __restrict__ x, int64_t ld,
                                                      const float* __restrict__ row_scale, int row_mod, float alpha,
                                                      float* __restrict__ out, int64_t R, int C, int rows_per_block) {
    pdl_grid_sync();
    constexpr int V = VecIO<T>::N;
    __shared__ float red[32][32][V + 1];
    const int c = (blockIdx.x * 32 + threadIdx.x) * V;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
    float s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = 0.f;
    if (c < C) {
        const T* xc = x + c;
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 128) {
            typename VecIO<T>::Raw raw[4];
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t ru = r + 32 * u;
                const bool ok = ru < r1;
                raw[u] = VecIO<T>::ldraw(xc + (ok ? ru : r) * ld);
                w[u] = !ok ? 0.f : (row_scale ? row_scale[(int)(ru % row_mod)] : 1.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float v[V];
                VecIO<T>::unpack(raw[u], v);
#pragma unroll
                for (int j = 0; j < V; ++j) s[j] = fmaf(v[j], w[u], s[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) red[threadIdx.y][threadIdx.x][j] = s[j];
    __syncthreads();
    // Every row group adds into the same C addresses and same-address atomics serialise in L2 (~10 us for 99 groups of
    // scalar adds, whatever the size of x): one 16-byte vector atomic per 4 columns cuts the serialised operations 4x.
    if (threadIdx.y < V / 4 && c < C) {      // row lane q finishes columns 4q..4q+3 of every vector
        float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] += red[i][threadIdx.x][threadIdx.y * 4 + j];
        }
        float* o = out + c + threadIdx.y * 4;
        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            atomicAdd(reinterpret_cast<float4*>(o), make_float4(alpha * t[0], alpha * t[1], alpha * t[2], alpha * t[3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(o + j, alpha * t[j]);
        }
    }
}

// ------------------------------------------------------------------ AdamW over the flat trainable buffer
// The 147 trainable tensors of the path are views of ONE flat fp32 buffer (and their gradients of one flat gradient
// buffer), so the optimizer step of the reference's AdamW (torch.optim.AdamW semantics: decoupled weight decay, bias
// correction; mmcv builds it from configs/recognition/vit/*.py `optimizer = dict(type='AdamW', ...)`) is one pass over
// four arrays instead of a multi-tensor launch chain.  wd_mask[i] != 0 selects the elements that decay (adapter weight
// matrices; biases, LayerNorm and temporal_embedding do not, `paramwise_cfg`).  `step` lives on the device (the call is
// captured into the step's CUDA graph): the caller increments it before the launch.
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, const uint8_t* __restrict__ wd_mask,
                                                         const float* __restrict__ step, float lr, float beta1, float beta2,
                                                         float eps, float wd, int64_t n) {
    pdl_grid_sync();
    const float t = step[0];
    const float bc1 = 1.f - powf(beta1, t), bc2s = sqrtf(1.f - powf(beta2, t));
    const float step_size = lr / bc1;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (i0 + 4 <= n) {
        float4 pp = *reinterpret_cast<float4*>(p + i0), gg = *reinterpret_cast<const float4*>(g + i0);
        float4 mm = *reinterpret_cast<float4*>(m + i0), vv = *reinterpret_cast<float4*>(v + i0);
        const uchar4 wm = wd_mask ? *reinterpret_cast<const uchar4*>(wd_mask + i0) : make_uchar4(1, 1, 1, 1);
        float* P = reinterpret_cast<float*>(&pp); const float* G = reinterpret_cast<const float*>(&gg);
        float* M = reinterpret_cast<float*>(&mm); float* V = reinterpret_cast<float*>(&vv);
        const unsigned char W4[4] = {wm.x, wm.y, wm.z, wm.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (W4[j]) P[j] *= 1.f - lr * wd;
            M[j] = beta1 * M[j] + (1.f - beta1) * G[j];
            V[j] = beta2 * V[j] + (1.f - beta2) * G[j] * G[j];
            P[j] -= step_size * M[j] / (sqrtf(V[j]) / bc2s + eps);
        }
        *reinterpret_cast<float4*>(p + i0) = pp; *reinterpret_cast<float4*>(m + i0) = mm; *reinterpret_cast<float4*>(v + i0) = vv;
    } else {
        for (int64_t i = i0; i < n; ++i) {
            float pi = p[i];
            if (!wd_mask || wd_mask[i]) pi *= 1.f - lr * wd;
            const float mi = beta1 * m[i] + (1.f - beta1) * g[i], vi = beta2 * v[i] + (1.f - beta2) * g[i] * g[i];
            m[i] = mi; v[i] = vi;
            p[i] = pi - step_size * mi / (sqrtf(vi) / bc2s + eps);
        }
    }
}

// Batched 2-D transposes in one launch: matrix b (blockIdx.z) is src + table[3b] of shape [table[3b+1], table[3b+2]],
// written transposed at the same element offset in dst.  (All adapter weights of a step: 6 per block.)
template <typename T>
__global__ void __launch_bounds__(256) transpose_batched_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                                const int64_t* __restrict__ table) {
    pdl_grid_sync();
    __shared__ T tile[32][33];
    const int64_t off = table[3 * blockIdx.z];
    const int R = (int)table[3 * blockIdx.z + 1], C = (int)table[3 * blockIdx.z + 2];
    const T* in = src + off;
    T* out = dst + off;
    const int tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
    for (int tidx = blockIdx.x; tidx < tiles_c * tiles_r; tidx += gridDim.x) {
        const int bx = tidx % tiles_c, by = tidx / tiles_c;
        int c = bx * 32 + threadIdx.x;
        for (int i = threadIdx.y; i < 32; i += 8) {
            int r = by * 32 + i;
            if (r < R && c < C) tile[i][threadIdx.x] = in[(int64_t)r * C + c];
        }
        __syncthreads();
        int r = by * 32 + threadIdx.x;
        for (int i = threadIdx.y; i < 32; i += 8) {
            int cc = bx * 32 + i;
            if (r < R && cc < C) out[(int64_t)cc * R + r] = tile[threadIdx.x][i];
        }
        __syncthreads();
    }
}

// out[t, d] += sum_{tok} dz[(f*n + tok), d] for frame f = blockIdx.y (t = f % T)
template <typename T>
__global__ void __launch_bounds__(256) temb_grad_kernel(const T* __restrict__ dz, float* __restrict__ out, int T_, int n,
                                                        int D) {
    pdl_grid_sync();
    __shared__ float red[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    int f = blockIdx.y;
    float s = 0.f;
    if (c < D)
        for (int tk = threadIdx.y; tk < n; tk += 8) s += ldf<T>(dz + ((int64_t)f * n + tk) * D + c);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(out + (int64_t)(f % T_) * D + c, t);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int R, int C) {
    pdl_grid_sync();
    __shared__ T tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int r = blockIdx.y * 32 + i;
        if (r < R && c < C) tile[i][threadIdx.x] = in[(int64_t)r * C + c];
    }
    __syncthreads();
    int r = blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int cc = blockIdx.x * 32 + i;
        if (r < R && cc < C) out[(int64_t)cc * R + r] = tile[threadIdx.x][i];
    }
}


// ------------------------------------------------------------------ fork block (vit_clip.py:264-275) combine
// out[f,i,:] = x[f,i,:] + (1 - lam[f]) * a_o[f,i,:] + rs[i] * s[f,:]      (s = scale * S_Adapter(lam * a_c), one row per frame)
template <typename T>
__global__ void __launch_bounds__(128) fork_combine_kernel(const T* __restrict__ x, const T* __restrict__ a_o,
                                                           const T* __restrict__ sfr, const float* __restrict__ lam,
                                                           const float* __restrict__ rs, T* __restrict__ out, int BT, int n,
                                                           int D) {
    pdl_grid_sync();
    constexpr int V = VecIO<T>::N;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= (int64_t)BT * n) return;
    const int f = (int)(row / n), i = (int)(row % n);
    const float w = 1.f - lam[f], m = rs ? rs[i] : 1.f;
    for (int c = lane * V; c < D; c += 32 * V) {
        float a[V], b[V], cc[V], o[V];
        VecIO<T>::ld(x + row * D + c, a);
        VecIO<T>::ld(a_o + row * D + c, b);
        VecIO<T>::ld(sfr + (int64_t)f * D + c, cc);
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = a[j] + w * b[j] + m * cc[j];
        VecIO<T>::st(out + row * D + c, o);
    }
}
// d_ao[f,i,:] = (1 - lam[f]) * dx[f,i,:] ;  d_s[f,:] = sum_i rs[i] * dx[f,i,:]     (block = one frame x 32*V columns)
template <typename T>
__global__ void __launch_bounds__(256) fork_combine_bwd_kernel(const T* __restrict__ dx, const float* __restrict__ lam,
                                                               const float* __restrict__ rs, T* __restrict__ d_ao,
                                                               T* __restrict__ d_s, int n, int D) {
    pdl_grid_sync();
    constexpr int V = VecIO<T>::N;
    __shared__ float red[8][32][V + 1];
    const int f = blockIdx.y;
    const int c = (blockIdx.x * 32 + threadIdx.x) * V;
    const float w = 1.f - lam[f];
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    if (c < D)
        for (int i = threadIdx.y; i < n; i += 8) {
            const int64_t off = ((int64_t)f * n + i) * D + c;
            float v[V], o[V];
            VecIO<T>::ld(dx + off, v);
            const float m = rs ? rs[i] : 1.f;
#pragma unroll
            for (int j = 0; j < V; ++j) { o[j] = w * v[j]; acc[j] = fmaf(m, v[j], acc[j]); }
            VecIO<T>::st(d_ao + off, o);
        }
#pragma unroll
    for (int j = 0; j < V; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && c < D) {
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x][j];
            o[j] = t;
        }
        VecIO<T>::st(d_s + (int64_t)f * D + c, o);
    }
}

static inline bool vec_ok(int D, int dtype) {
    int v = dtype == AIMB_BF16 ? 8 : 4;
    return D > 0 && D % v == 0 && D <= MAXD;
}

}  // namespace aimb

using namespace aimb;

extern "C" int aimb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                                  int64_t rows, int32_t D, float eps, int32_t dtype, void* stream) {
    if (!x || !gamma || !beta || (!y && !(mean && rstd)) || rows < 0 || !vec_ok(D, dtype)) return AIMB_ERR_ARG;
    if (rows == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned grid = (unsigned)((rows + 3) / 4);
    if (dtype == AIMB_BF16)
        launch_k((layernorm_fwd_kernel<bf16>), dim3(grid), dim3(128), 0, s, (const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y,
                                                        mean, rstd, rows, D, eps);
    else if (dtype == AIMB_F32)
        launch_k((layernorm_fwd_kernel<float>), dim3(grid), dim3(128), 0, s, (const float*)x, (const float*)gamma, (const float*)beta,
                                                         (float*)y, mean, rstd, rows, D, eps);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

template <typename T, int NIT>
static void ln_bwd_launch_t(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                            const void* dres, void* dx, const float* cs_w, int cs_mod, float cs_alpha, float* cs_out,
                            int64_t rows, int D, cudaStream_t s) {
    if (cs_out) {
        constexpr int RPW = 8;
        unsigned grid = (unsigned)((rows + 4 * RPW - 1) / (4 * RPW));
        launch_k((layernorm_bwd_kernel<T, RPW, NIT>), dim3(grid), dim3(128), 0, s, (const T*)dy, (const T*)x, mean, rstd,
                 (const T*)gamma, (const T*)dres, (T*)dx, cs_w, cs_mod > 0 ? cs_mod : 1, cs_alpha, cs_out, rows, D);
    } else {
        unsigned grid = (unsigned)((rows + 3) / 4);
        launch_k((layernorm_bwd_kernel<T, 1, NIT>), dim3(grid), dim3(128), 0, s, (const T*)dy, (const T*)x, mean, rstd,
                 (const T*)gamma, (const T*)dres, (T*)dx, (const float*)nullptr, 1, 1.f, (float*)nullptr, rows, D);
    }
}

static int ln_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                         const void* dres, void* dx, const float* cs_w, int cs_mod, float cs_alpha, float* cs_out,
                         int64_t rows, int D, int dtype, cudaStream_t s) {
    if (!dy || !x || !mean || !rstd || !gamma || !dx || rows < 0 || !vec_ok(D, dtype)) return AIMB_ERR_ARG;
    if (cs_out && cudaMemsetAsync(cs_out, 0, (size_t)D * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (rows == 0) return AIMB_OK;
    const int vec = dtype == AIMB_BF16 ? 8 : 4;
    const int nit = (D + 32 * vec - 1) / (32 * vec);
#define AIMB_LN_BWD(T_, N_) ln_bwd_launch_t<T_, N_>(dy, x, mean, rstd, gamma, dres, dx, cs_w, cs_mod, cs_alpha, cs_out, rows, D, s)
    if (dtype == AIMB_BF16) {
        if (nit <= 1) AIMB_LN_BWD(bf16, 1); else if (nit <= 2) AIMB_LN_BWD(bf16, 2); else if (nit <= 3) AIMB_LN_BWD(bf16, 3);
        else AIMB_LN_BWD(bf16, 4);
    } else if (dtype == AIMB_F32) {
        if (nit <= 2) AIMB_LN_BWD(float, 2); else if (nit <= 4) AIMB_LN_BWD(float, 4); else if (nit <= 6) AIMB_LN_BWD(float, 6);
        else AIMB_LN_BWD(float, 8);
    } else return AIMB_ERR_ARG;
#undef AIMB_LN_BWD
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                                  const void* dres, void* dx, int64_t rows, int32_t D, int32_t dtype, void* stream) {
    return ln_bwd_launch(dy, x, mean, rstd, gamma, dres, dx, nullptr, 1, 1.f, nullptr, rows, D, dtype, (cudaStream_t)stream);
}

extern "C" int aimb_layernorm_bwd_colsum(const void* dy, const void* x, const float* mean, const float* rstd,
                                         const void* gamma, const void* dres, void* dx, const float* cs_row_scale,
                                         int32_t cs_row_mod, float cs_alpha, float* cs_out, int64_t rows, int32_t D,
                                         int32_t dtype, void* stream) {
    if (!cs_out || (cs_row_scale && cs_row_mod <= 0)) return AIMB_ERR_ARG;
    return ln_bwd_launch(dy, x, mean, rstd, gamma, dres, dx, cs_row_scale, cs_row_mod, cs_alpha, cs_out, rows, D, dtype,
                         (cudaStream_t)stream);
}

extern "C" int aimb_stem_assemble_ln(const void* tok, const void* cls, const void* pos, const void* temb,
                                     const void* gamma, const void* beta, void* z, void* x, float* mean, float* rstd,
                                     int32_t B, int32_t T, int32_t n, int32_t D, float eps, int32_t dtype, void* stream) {
    if (!tok || !cls || !pos || !temb || !gamma || !beta || !x || B < 0 || T <= 0 || n < 2 || !vec_ok(D, dtype))
        return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t rows = (int64_t)B * T * n;
    unsigned grid = (unsigned)((rows + 3) / 4);
    if (dtype == AIMB_BF16)
        launch_k((stem_assemble_ln_kernel<bf16>), dim3(grid), dim3(128), 0, s, (const bf16*)tok, (const bf16*)cls, (const bf16*)pos,
                                                           (const bf16*)temb, (const bf16*)gamma, (const bf16*)beta,
                                                           (bf16*)z, (bf16*)x, mean, rstd, B * T, T, n, D, eps);
    else if (dtype == AIMB_F32)
        launch_k((stem_assemble_ln_kernel<float>), dim3(grid), dim3(128), 0, s, (const float*)tok, (const float*)cls, (const float*)pos,
                                                            (const float*)temb, (const float*)gamma, (const float*)beta,
                                                            (float*)z, (float*)x, mean, rstd, B * T, T, n, D, eps);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

template <typename TI>
static int im2col_dispatch(const void* x, const float* mean, const float* std_, void* cols, int dtype, int B, int T, int H,
                           int W, int p, int kpad, cudaStream_t s) {
    int64_t total = (int64_t)B * T * 3 * (H / p) * p * (W / p);
    unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == AIMB_BF16)
        launch_k((im2col_kernel<TI, bf16>), dim3(grid), dim3(256), 0, s, (const TI*)x, mean, std_, (bf16*)cols, B, T, H, W, p, kpad);
    else if (dtype == AIMB_F32)
        launch_k((im2col_kernel<TI, float>), dim3(grid), dim3(256), 0, s, (const TI*)x, mean, std_, (float*)cols, B, T, H, W, p, kpad);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_im2col(const void* x, int32_t x_dtype, const float* mean, const float* std_, void* cols, int32_t dtype,
                           int32_t B, int32_t T, int32_t H, int32_t W, int32_t patch, int32_t kpad, void* stream) {
    if (!x || !cols || B < 0 || T <= 0 || patch <= 0 || H % patch || W % patch || kpad < 3 * patch * patch)
        return AIMB_ERR_ARG;
    if ((mean == nullptr) != (std_ == nullptr)) return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype == AIMB_F32) return im2col_dispatch<float>(x, mean, std_, cols, dtype, B, T, H, W, patch, kpad, s);
    if (x_dtype == AIMB_BF16) return im2col_dispatch<bf16>(x, mean, std_, cols, dtype, B, T, H, W, patch, kpad, s);
    if (x_dtype == AIMB_U8) return im2col_dispatch<uint8_t>(x, mean, std_, cols, dtype, B, T, H, W, patch, kpad, s);
    return AIMB_ERR_ARG;
}

extern "C" int aimb_tail_fwd(const void* x, const void* gamma, const void* beta, float* feat, float* mean, float* rstd,
                             int32_t B, int32_t T, int32_t n, int32_t D, float eps, int32_t dtype, void* stream) {
    if (!x || !gamma || !beta || !feat || !mean || !rstd || B < 0 || T <= 0 || n <= 0 || !vec_ok(D, dtype))
        return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned grid = (unsigned)((B * T + 3) / 4);
    if (dtype == AIMB_BF16)
        launch_k((tail_fwd_kernel<bf16>), dim3(grid), dim3(128), 0, s, (const bf16*)x, (const bf16*)gamma, (const bf16*)beta, feat, mean, rstd,
                                                   B * T, T, n, D, eps);
    else if (dtype == AIMB_F32)
        launch_k((tail_fwd_kernel<float>), dim3(grid), dim3(128), 0, s, (const float*)x, (const float*)gamma, (const float*)beta, feat, mean,
                                                    rstd, B * T, T, n, D, eps);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_tail_bwd(const float* dfeat, const void* x, const float* mean, const float* rstd, const void* gamma,
                             void* dx, float* dgamma, float* dbeta, int32_t B, int32_t T, int32_t n, int32_t D,
                             int32_t dtype, void* stream) {
    if (!dfeat || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || B < 0 || T <= 0 || n <= 0 ||
        !vec_ok(D, dtype))
        return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t esz = dtype == AIMB_BF16 ? 2 : 4;
    if (cudaMemsetAsync(dx, 0, (size_t)B * T * n * D * esz, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (cudaMemsetAsync(dgamma, 0, (size_t)D * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (cudaMemsetAsync(dbeta, 0, (size_t)D * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    unsigned grid = (unsigned)((B * T + 3) / 4);
    if (dtype == AIMB_BF16)
        launch_k((tail_bwd_kernel<bf16>), dim3(grid), dim3(128), 0, s, dfeat, (const bf16*)x, mean, rstd, (const bf16*)gamma, (bf16*)dx, dgamma,
                                                   dbeta, B * T, T, n, D);
    else if (dtype == AIMB_F32)
        launch_k((tail_bwd_kernel<float>), dim3(grid), dim3(128), 0, s, dfeat, (const float*)x, mean, rstd, (const float*)gamma, (float*)dx,
                                                    dgamma, dbeta, B * T, T, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

static int g_colsum_rpb = 0;
extern "C" void aimb_debug_colsum_rpb(int rows_per_block) { g_colsum_rpb = rows_per_block; }

extern "C" int aimb_colsum(const void* x, int64_t ld, const float* row_scale, int32_t row_mod, float alpha, float* out,
                           int64_t R, int32_t C, int32_t accumulate, int32_t dtype, void* stream) {
    if (!x || !out || R < 0 || C <= 0 || ld < C || (row_scale && row_mod <= 0)) return AIMB_ERR_ARG;
    const int V = dtype == AIMB_BF16 ? 8 : 4;
    if (C % V || ld % V || ((uintptr_t)x & 15)) return AIMB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate && cudaMemsetAsync(out, 0, (size_t)C * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (R == 0) return AIMB_OK;
    const int gx = (C / V + 31) / 32;
    // One wave: two 1024-thread blocks fit an SM, so at most 2 * #SMs blocks.  (With a fixed 128 rows per block the
    // M = 12 608, C = 768 reduction of the step was 297 blocks on 296 slots: a second wave for one block doubled its time.)
    // Not more than ~112 row groups either: every group adds into the same C addresses and the atomics serialise in L2.
    int64_t max_groups = (2 * device_sm_count()) / gx;
    if (max_groups > 112) max_groups = 112;
    if (max_groups < 1) max_groups = 1;
    int64_t rpb64 = (R + max_groups - 1) / max_groups;
    if (rpb64 < 32) rpb64 = 32;
    if (g_colsum_rpb > 0) rpb64 = g_colsum_rpb;          // bench_tools only (aimb_debug_colsum_rpb)
    const int rpb = (int)(rpb64 > (1 << 30) ? (1 << 30) : rpb64);
    dim3 grid(gx, (unsigned)((R + rpb - 1) / rpb)), block(32, 32);
    if (dtype == AIMB_BF16)
        launch_k((colsum_kernel<bf16>), dim3(grid), dim3(block), 0, s, (const bf16*)x, ld, row_scale, row_mod, alpha, out, R, C, rpb);
    else if (dtype == AIMB_F32)
        launch_k((colsum_kernel<float>), dim3(grid), dim3(block), 0, s, (const float*)x, ld, row_scale, row_mod, alpha, out, R, C, rpb);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_transpose_batched(const void* src, void* dst, const int64_t* table, int32_t nmat, int32_t dtype,
                                      void* stream) {
    if (!src || !dst || !table || nmat < 0) return AIMB_ERR_ARG;
    if (nmat == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(48, 1, nmat), block(32, 8);
    if (dtype == AIMB_BF16)
        launch_k((transpose_batched_kernel<bf16>), dim3(grid), dim3(block), 0, s, (const bf16*)src, (bf16*)dst, table);
    else if (dtype == AIMB_F32)
        launch_k((transpose_batched_kernel<float>), dim3(grid), dim3(block), 0, s, (const float*)src, (float*)dst, table);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_temb_grad(const void* dz, float* out, int32_t B, int32_t T, int32_t n, int32_t D, int32_t dtype,
                              void* stream) {
    if (!dz || !out || B < 0 || T <= 0 || n <= 0 || D <= 0) return AIMB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(out, 0, (size_t)T * D * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (B == 0) return AIMB_OK;
    dim3 grid((D + 31) / 32, B * T), block(32, 8);
    if (dtype == AIMB_BF16) launch_k((temb_grad_kernel<bf16>), dim3(grid), dim3(block), 0, s, (const bf16*)dz, out, T, n, D);
    else if (dtype == AIMB_F32) launch_k((temb_grad_kernel<float>), dim3(grid), dim3(block), 0, s, (const float*)dz, out, T, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_transpose(const void* in, void* out, int32_t R, int32_t C, int32_t dtype, void* stream) {
    if (!in || !out || R <= 0 || C <= 0) return AIMB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    if (dtype == AIMB_BF16) launch_k((transpose_kernel<bf16>), dim3(grid), dim3(block), 0, s, (const bf16*)in, (bf16*)out, R, C);
    else if (dtype == AIMB_F32) launch_k((transpose_kernel<float>), dim3(grid), dim3(block), 0, s, (const float*)in, (float*)out, R, C);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_fork_combine(const void* x, const void* a_o, const void* s_frame, const float* lam, const float* rs,
                                 void* out, int32_t BT, int32_t n, int32_t D, int32_t dtype, void* stream) {
    if (!x || !a_o || !s_frame || !lam || !out || BT < 0 || n <= 0 || !vec_ok(D, dtype)) return AIMB_ERR_ARG;
    if (BT == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned grid = (unsigned)(((int64_t)BT * n + 3) / 4);
    if (dtype == AIMB_BF16)
        launch_k((fork_combine_kernel<bf16>), dim3(grid), dim3(128), 0, s, (const bf16*)x, (const bf16*)a_o, (const bf16*)s_frame, lam, rs, (bf16*)out,
                                                       BT, n, D);
    else if (dtype == AIMB_F32)
        launch_k((fork_combine_kernel<float>), dim3(grid), dim3(128), 0, s, (const float*)x, (const float*)a_o, (const float*)s_frame, lam, rs,
                                                        (float*)out, BT, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_fork_combine_bwd(const void* dx, const float* lam, const float* rs, void* d_ao, void* d_s, int32_t BT,
                                     int32_t n, int32_t D, int32_t dtype, void* stream) {
    if (!dx || !lam || !d_ao || !d_s || BT < 0 || n <= 0 || !vec_ok(D, dtype)) return AIMB_ERR_ARG;
    if (BT == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int V = dtype == AIMB_BF16 ? 8 : 4;
    dim3 grid((D / V + 31) / 32, BT), block(32, 8);
    if (dtype == AIMB_BF16)
        launch_k((fork_combine_bwd_kernel<bf16>), dim3(grid), dim3(block), 0, s, (const bf16*)dx, lam, rs, (bf16*)d_ao, (bf16*)d_s, n, D);
    else if (dtype == AIMB_F32)
        launch_k((fork_combine_bwd_kernel<float>), dim3(grid), dim3(block), 0, s, (const float*)dx, lam, rs, (float*)d_ao, (float*)d_s, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_adamw_flat(float* p, const float* g, float* m, float* v, const uint8_t* wd_mask, const float* step, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int64_t n, void* stream) {
    if (!p || !g || !m || !v || !step || n < 0) return AIMB_ERR_ARG;
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15 || (wd_mask && ((uintptr_t)wd_mask & 3))) return AIMB_ERR_ARG;
    if (n == 0) return AIMB_OK;
    const unsigned grid = (unsigned)((n + 1023) / 1024);
    launch_k(adamw_flat_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, wd_mask, step, lr, beta1, beta2, eps,
             weight_decay, n);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

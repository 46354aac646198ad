// SIMT attention cores (fp32 math, any storage dtype).
//  * spatial fwd/bwd : fp32 parity mode + cross-check of the tensor-core flash kernels (attn_mma.cu)
//  * temporal fwd/bwd: the production T-Adapter attention.  One warp per (clip, token, head); the
//    `n (b t) d -> t (b n) d` rearrange of vitclip_aim.py:200 is never materialised: the T rows of a
//    sequence are read in place with row stride n*3D (each row-head slice is one 128 B/256 B line).
#include "common.cuh"

namespace aimb {

constexpr int DH = 64;         // head_dim of every CLIP ViT on this path
constexpr int KS = DH + 1;     // padded smem row stride (conflict-free when lanes index rows)
constexpr float ATT_SCALE = 0.125f;  // 1/sqrt(64)   (vit_clip.py:147)

template <typename T>
__device__ __forceinline__ void load_rows_to_smem(float* dst, const T* src, int64_t row_stride, int rows, int tid,
                                                  int nthreads) {
    // rows x 64 elements, coalesced over the 64 contiguous columns
    for (int e = tid; e < rows * DH; e += nthreads) {
        int r = e >> 6, c = e & 63;
        dst[r * KS + c] = ldf<T>(src + (int64_t)r * row_stride + c);
    }
}

// ------------------------------------------------------------------ spatial forward
template <typename T>
__global__ void __launch_bounds__(256) attn_spatial_fwd_simt(const T* __restrict__ qkv, T* __restrict__ o,
                                                             float* __restrict__ lse, int n, int heads) {
    extern __shared__ float sm[];
    const int D = heads * DH, ld = 3 * D;
    const int f = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npad = (n + 31) & ~31;
    float* sK = sm;
    float* sV = sK + n * KS;
    float* sQ = sV + n * KS;           // [8][64]
    float* sP = sQ + 8 * DH;           // [8][npad]
    const T* base = qkv + (int64_t)f * n * ld + h * DH;
    load_rows_to_smem<T>(sK, base + D, ld, n, tid, 256);
    load_rows_to_smem<T>(sV, base + 2 * D, ld, n, tid, 256);
    __syncthreads();
    const int slots = npad / 32;
    for (int i = warp; i < n; i += 8) {
        const T* qrow = base + (int64_t)i * ld;
        sQ[warp * DH + lane] = ldf<T>(qrow + lane);
        sQ[warp * DH + lane + 32] = ldf<T>(qrow + lane + 32);
        __syncwarp();
        float mx = -INFINITY;
        for (int s = 0; s < slots; ++s) {
            int j = s * 32 + lane;
            float acc = -INFINITY;
            if (j < n) {
                acc = 0.f;
#pragma unroll 16
                for (int d = 0; d < DH; ++d) acc = fmaf(sQ[warp * DH + d], sK[j * KS + d], acc);
                acc *= ATT_SCALE;
            }
            sP[warp * npad + j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int s = 0; s < slots; ++s) {
            int j = s * 32 + lane;
            float p = (j < n) ? __expf(sP[warp * npad + j] - mx) : 0.f;
            sP[warp * npad + j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < n; ++j) {
            float p = sP[warp * npad + j];
            o0 = fmaf(p, sV[j * KS + lane], o0);
            o1 = fmaf(p, sV[j * KS + lane + 32], o1);
        }
        float inv = 1.f / sum;
        T* orow = o + ((int64_t)f * n + i) * D + h * DH;
        stf<T>(orow + lane, o0 * inv);
        stf<T>(orow + lane + 32, o1 * inv);
        if (lse && lane == 0) lse[((int64_t)f * heads + h) * n + i] = mx + __logf(sum);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ spatial backward (two phases, no atomics)
template <typename T>
__global__ void __launch_bounds__(256) attn_spatial_bwd_simt(const T* __restrict__ qkv, const T* __restrict__ o,
                                                             const T* __restrict__ d_o, const float* __restrict__ lse,
                                                             T* __restrict__ d_qkv, int n, int heads) {
    extern __shared__ float sm[];
    const int D = heads * DH, ld = 3 * D;
    const int f = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npad = (n + 31) & ~31;
    float* sA = sm;                      // phase 1: K   phase 2: Q
    float* sB = sA + n * KS;             // phase 1: V   phase 2: dO
    float* sLse = sB + n * KS;           // [npad]
    float* sDelta = sLse + npad;         // [npad]
    float* sRowA = sDelta + npad;        // [8][64]
    float* sRowB = sRowA + 8 * DH;       // [8][64]
    float* sP = sRowB + 8 * DH;          // [8][npad]
    float* sDS = sP + 8 * npad;          // [8][npad]
    const T* base = qkv + (int64_t)f * n * ld + h * DH;
    T* dbase = d_qkv + (int64_t)f * n * ld + h * DH;
    const T* obase = o + (int64_t)f * n * D + h * DH;
    const T* dobase = d_o + (int64_t)f * n * D + h * DH;
    const float* lrow = lse + ((int64_t)f * heads + h) * n;
    const int slots = npad / 32;

    load_rows_to_smem<T>(sA, base + D, ld, n, tid, 256);
    load_rows_to_smem<T>(sB, base + 2 * D, ld, n, tid, 256);
    for (int i = tid; i < npad; i += 256) sLse[i] = i < n ? lrow[i] : 0.f;
    __syncthreads();
    // ---- phase 1: warp owns query row i -> dQ_i, delta_i
    for (int i = warp; i < n; i += 8) {
        float q0 = ldf<T>(base + (int64_t)i * ld + lane), q1 = ldf<T>(base + (int64_t)i * ld + lane + 32);
        float g0 = ldf<T>(dobase + (int64_t)i * D + lane), g1 = ldf<T>(dobase + (int64_t)i * D + lane + 32);
        float oo0 = ldf<T>(obase + (int64_t)i * D + lane), oo1 = ldf<T>(obase + (int64_t)i * D + lane + 32);
        float delta = warp_sum(g0 * oo0 + g1 * oo1);
        sRowA[warp * DH + lane] = q0; sRowA[warp * DH + lane + 32] = q1;
        sRowB[warp * DH + lane] = g0; sRowB[warp * DH + lane + 32] = g1;
        if (lane == 0) sDelta[i] = delta;
        __syncwarp();
        float li = sLse[i];
        for (int s = 0; s < slots; ++s) {
            int j = s * 32 + lane;
            float ds = 0.f;
            if (j < n) {
                float sc = 0.f, dp = 0.f;
#pragma unroll 16
                for (int d = 0; d < DH; ++d) {
                    sc = fmaf(sRowA[warp * DH + d], sA[j * KS + d], sc);
                    dp = fmaf(sRowB[warp * DH + d], sB[j * KS + d], dp);
                }
                float p = __expf(sc * ATT_SCALE - li);
                ds = p * (dp - delta) * ATT_SCALE;
            }
            sDS[warp * npad + j] = ds;
        }
        __syncwarp();
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < n; ++j) {
            float ds = sDS[warp * npad + j];
            a0 = fmaf(ds, sA[j * KS + lane], a0);
            a1 = fmaf(ds, sA[j * KS + lane + 32], a1);
        }
        stf<T>(dbase + (int64_t)i * ld + lane, a0);
        stf<T>(dbase + (int64_t)i * ld + lane + 32, a1);
        __syncwarp();
    }
    __syncthreads();
    // ---- phase 2: warp owns key row j -> dK_j, dV_j   (smem now holds Q and dO)
    load_rows_to_smem<T>(sA, base, ld, n, tid, 256);
    load_rows_to_smem<T>(sB, dobase, D, n, tid, 256);
    __syncthreads();
    for (int j = warp; j < n; j += 8) {
        sRowA[warp * DH + lane] = ldf<T>(base + D + (int64_t)j * ld + lane);
        sRowA[warp * DH + lane + 32] = ldf<T>(base + D + (int64_t)j * ld + lane + 32);
        sRowB[warp * DH + lane] = ldf<T>(base + 2 * D + (int64_t)j * ld + lane);
        sRowB[warp * DH + lane + 32] = ldf<T>(base + 2 * D + (int64_t)j * ld + lane + 32);
        __syncwarp();
        for (int s = 0; s < slots; ++s) {
            int i = s * 32 + lane;
            float p = 0.f, ds = 0.f;
            if (i < n) {
                float sc = 0.f, dp = 0.f;
#pragma unroll 16
                for (int d = 0; d < DH; ++d) {
                    sc = fmaf(sA[i * KS + d], sRowA[warp * DH + d], sc);
                    dp = fmaf(sB[i * KS + d], sRowB[warp * DH + d], dp);
                }
                p = __expf(sc * ATT_SCALE - sLse[i]);
                ds = p * (dp - sDelta[i]) * ATT_SCALE;
            }
            sP[warp * npad + i] = p;
            sDS[warp * npad + i] = ds;
        }
        __syncwarp();
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int i = 0; i < n; ++i) {
            float p = sP[warp * npad + i], ds = sDS[warp * npad + i];
            v0 = fmaf(p, sB[i * KS + lane], v0);
            v1 = fmaf(p, sB[i * KS + lane + 32], v1);
            k0 = fmaf(ds, sA[i * KS + lane], k0);
            k1 = fmaf(ds, sA[i * KS + lane + 32], k1);
        }
        stf<T>(dbase + D + (int64_t)j * ld + lane, k0);
        stf<T>(dbase + D + (int64_t)j * ld + lane + 32, k1);
        stf<T>(dbase + 2 * D + (int64_t)j * ld + lane, v0);
        stf<T>(dbase + 2 * D + (int64_t)j * ld + lane + 32, v1);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ temporal forward: warp per (b, token, head)
template <typename T>
__global__ void __launch_bounds__(128) attn_temporal_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o, int B,
                                                                int T_, int n, int heads) {
    extern __shared__ float sm[];
    const int D = heads * DH, ld = 3 * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t prob = (int64_t)blockIdx.x * 4 + warp;
    if (prob >= (int64_t)B * n * heads) return;
    const int h = (int)(prob % heads);
    const int tok = (int)((prob / heads) % n);
    const int b = (int)(prob / ((int64_t)heads * n));
    const int per_warp = 3 * T_ * KS + T_ * (T_ + 1);
    float* sQ = sm + warp * per_warp;
    float* sK = sQ + T_ * KS;
    float* sV = sK + T_ * KS;
    float* sS = sV + T_ * KS;
    const int64_t row0 = (int64_t)b * T_ * n + tok;   // row of frame t: row0 + t*n
    const T* base = qkv + row0 * ld + h * DH;
    const int64_t rs = (int64_t)n * ld;
    for (int t = 0; t < T_; ++t) {
        const T* r = base + t * rs;
        sQ[t * KS + lane] = ldf<T>(r + lane);           sQ[t * KS + lane + 32] = ldf<T>(r + lane + 32);
        sK[t * KS + lane] = ldf<T>(r + D + lane);       sK[t * KS + lane + 32] = ldf<T>(r + D + lane + 32);
        sV[t * KS + lane] = ldf<T>(r + 2 * D + lane);   sV[t * KS + lane + 32] = ldf<T>(r + 2 * D + lane + 32);
    }
    __syncwarp();
    for (int p = lane; p < T_ * T_; p += 32) {
        int i = p / T_, j = p % T_;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) acc = fmaf(sQ[i * KS + d], sK[j * KS + d], acc);
        sS[i * (T_ + 1) + j] = acc * ATT_SCALE;
    }
    __syncwarp();
    if (lane < T_) {
        float* row = sS + lane * (T_ + 1);
        float mx = -INFINITY;
        for (int j = 0; j < T_; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < T_; ++j) { float e = __expf(row[j] - mx); row[j] = e; sum += e; }
        float inv = 1.f / sum;
        for (int j = 0; j < T_; ++j) row[j] *= inv;
    }
    __syncwarp();
    T* obase = o + row0 * D + h * DH;
    for (int i = 0; i < T_; ++i) {
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < T_; ++j) {
            float p = sS[i * (T_ + 1) + j];
            o0 = fmaf(p, sV[j * KS + lane], o0);
            o1 = fmaf(p, sV[j * KS + lane + 32], o1);
        }
        stf<T>(obase + (int64_t)i * n * D + lane, o0);
        stf<T>(obase + (int64_t)i * n * D + lane + 32, o1);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_temporal_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ d_o,
                                                                T* __restrict__ d_qkv, int B, int T_, int n, int heads) {
    extern __shared__ float sm[];
    const int D = heads * DH, ld = 3 * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t prob = (int64_t)blockIdx.x * 4 + warp;
    if (prob >= (int64_t)B * n * heads) return;
    const int h = (int)(prob % heads);
    const int tok = (int)((prob / heads) % n);
    const int b = (int)(prob / ((int64_t)heads * n));
    const int per_warp = 4 * T_ * KS + 2 * T_ * (T_ + 1);
    float* sQ = sm + warp * per_warp;
    float* sK = sQ + T_ * KS;
    float* sV = sK + T_ * KS;
    float* sG = sV + T_ * KS;            // dO
    float* sP = sG + T_ * KS;            // probabilities
    float* sD = sP + T_ * (T_ + 1);      // dP then dS
    const int64_t row0 = (int64_t)b * T_ * n + tok;
    const T* base = qkv + row0 * ld + h * DH;
    const T* gbase = d_o + row0 * D + h * DH;
    const int64_t rs = (int64_t)n * ld;
    for (int t = 0; t < T_; ++t) {
        const T* r = base + t * rs;
        const T* g = gbase + (int64_t)t * n * D;
        sQ[t * KS + lane] = ldf<T>(r + lane);           sQ[t * KS + lane + 32] = ldf<T>(r + lane + 32);
        sK[t * KS + lane] = ldf<T>(r + D + lane);       sK[t * KS + lane + 32] = ldf<T>(r + D + lane + 32);
        sV[t * KS + lane] = ldf<T>(r + 2 * D + lane);   sV[t * KS + lane + 32] = ldf<T>(r + 2 * D + lane + 32);
        sG[t * KS + lane] = ldf<T>(g + lane);           sG[t * KS + lane + 32] = ldf<T>(g + lane + 32);
    }
    __syncwarp();
    const int TS = T_ + 1;
    for (int p = lane; p < T_ * T_; p += 32) {
        int i = p / T_, j = p % T_;
        float sc = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) {
            sc = fmaf(sQ[i * KS + d], sK[j * KS + d], sc);
            dp = fmaf(sG[i * KS + d], sV[j * KS + d], dp);
        }
        sP[i * TS + j] = sc * ATT_SCALE;
        sD[i * TS + j] = dp;
    }
    __syncwarp();
    if (lane < T_) {
        float* row = sP + lane * TS;
        float* drow = sD + lane * TS;
        float mx = -INFINITY;
        for (int j = 0; j < T_; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < T_; ++j) { float e = __expf(row[j] - mx); row[j] = e; sum += e; }
        float inv = 1.f / sum, delta = 0.f;
        for (int j = 0; j < T_; ++j) { row[j] *= inv; delta = fmaf(row[j], drow[j], delta); }
        for (int j = 0; j < T_; ++j) drow[j] = row[j] * (drow[j] - delta) * ATT_SCALE;
    }
    __syncwarp();
    T* dbase = d_qkv + row0 * ld + h * DH;
    for (int i = 0; i < T_; ++i) {
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int j = 0; j < T_; ++j) {
            float ds_ij = sD[i * TS + j];   // row i: dQ_i += dS_ij K_j
            q0 = fmaf(ds_ij, sK[j * KS + lane], q0);
            q1 = fmaf(ds_ij, sK[j * KS + lane + 32], q1);
            float ds_ji = sD[j * TS + i];   // column i: dK_i += dS_ji Q_j ; dV_i += P_ji dO_j
            float p_ji = sP[j * TS + i];
            k0 = fmaf(ds_ji, sQ[j * KS + lane], k0);
            k1 = fmaf(ds_ji, sQ[j * KS + lane + 32], k1);
            v0 = fmaf(p_ji, sG[j * KS + lane], v0);
            v1 = fmaf(p_ji, sG[j * KS + lane + 32], v1);
        }
        T* r = dbase + i * rs;
        stf<T>(r + lane, q0);            stf<T>(r + lane + 32, q1);
        stf<T>(r + D + lane, k0);        stf<T>(r + D + lane + 32, k1);
        stf<T>(r + 2 * D + lane, v0);    stf<T>(r + 2 * D + lane + 32, v1);
    }
}

// ------------------------------------------------------------------ fork block weights (vit_clip.py:147-151,182-186)
// w_o[f] = sum_{i,j} exp( q_i . k_j / 8 ) over the FULL width D (sum over heads of per-head logits),
// w_c[f] = sum_i exp( q_i . kc_f / 8 ).  fp32, no max subtraction (as the reference).
template <typename T>
__global__ void __launch_bounds__(256) fork_weights_kernel(const T* __restrict__ qkv, const T* __restrict__ kc,
                                                           float* __restrict__ w_o, float* __restrict__ w_c, int n,
                                                           int D) {
    extern __shared__ float sm[];
    const int f = blockIdx.x, i0 = blockIdx.y * 8;   // 8 query rows per block (one per warp)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = i0 + warp;
    float* sq = sm + warp * D;
    const int ld = 3 * D;
    const T* fb = qkv + (int64_t)f * n * ld;
    float tot_o = 0.f, tot_c = 0.f;
    if (i < n) {
        for (int d = lane; d < D; d += 32) sq[d] = ldf<T>(fb + (int64_t)i * ld + d);
        __syncwarp();
        for (int j = 0; j < n; ++j) {
            const T* kr = fb + (int64_t)j * ld + D;
            float acc = 0.f;
            for (int d = lane; d < D; d += 32) acc = fmaf(sq[d], ldf<T>(kr + d), acc);
            acc = warp_sum(acc);
            tot_o += __expf(acc * ATT_SCALE);
        }
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(sq[d], ldf<T>(kc + (int64_t)f * D + d), acc);
        acc = warp_sum(acc);
        tot_c = __expf(acc * ATT_SCALE);
        if (lane == 0) { atomicAdd(w_o + f, tot_o); atomicAdd(w_c + f, tot_c); }
    }
}

static size_t spatial_fwd_smem(int n) { int npad = (n + 31) & ~31; return (size_t)(2 * n * KS + 8 * DH + 8 * npad) * 4; }
static size_t spatial_bwd_smem(int n) {
    int npad = (n + 31) & ~31;
    return (size_t)(2 * n * KS + 2 * npad + 16 * DH + 16 * npad) * 4;
}

template <typename T>
int spatial_fwd_simt_launch(const void* qkv, void* o, float* lse, int frames, int n, int heads, cudaStream_t s) {
    size_t smem = spatial_fwd_smem(n);
    if (smem > 227 * 1024) return AIMB_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(attn_spatial_fwd_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return AIMB_ERR_CUDA;
        attr_set = true;
    }
    attn_spatial_fwd_simt<T><<<frames * heads, 256, smem, s>>>((const T*)qkv, (T*)o, lse, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <typename T>
int spatial_bwd_simt_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames,
                            int n, int heads, cudaStream_t s) {
    size_t smem = spatial_bwd_smem(n);
    if (smem > 227 * 1024) return AIMB_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(attn_spatial_bwd_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return AIMB_ERR_CUDA;
        attr_set = true;
    }
    attn_spatial_bwd_simt<T><<<frames * heads, 256, smem, s>>>((const T*)qkv, (const T*)o, (const T*)d_o, lse, (T*)d_qkv, n,
                                                               heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

int attn_spatial_fwd_simt_dispatch(const void* qkv, void* o, float* lse, int frames, int n, int heads, int dtype,
                                   cudaStream_t s) {
    if (dtype == AIMB_BF16) return spatial_fwd_simt_launch<bf16>(qkv, o, lse, frames, n, heads, s);
    if (dtype == AIMB_F32) return spatial_fwd_simt_launch<float>(qkv, o, lse, frames, n, heads, s);
    return AIMB_ERR_ARG;
}
int attn_spatial_bwd_simt_dispatch(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv,
                                   int frames, int n, int heads, int dtype, cudaStream_t s) {
    if (dtype == AIMB_BF16) return spatial_bwd_simt_launch<bf16>(qkv, o, d_o, lse, d_qkv, frames, n, heads, s);
    if (dtype == AIMB_F32) return spatial_bwd_simt_launch<float>(qkv, o, d_o, lse, d_qkv, frames, n, heads, s);
    return AIMB_ERR_ARG;
}

}  // namespace aimb

using namespace aimb;

extern "C" int aimb_attn_temporal_fwd(const void* qkv, void* o, int32_t B, int32_t T, int32_t n, int32_t heads,
                                      int32_t dtype, void* stream) {
    if (!qkv || !o || B < 0 || T <= 0 || T > 32 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = (size_t)4 * (3 * T * KS + T * (T + 1)) * 4;
    int64_t probs = (int64_t)B * n * heads;
    unsigned grid = (unsigned)((probs + 3) / 4);
    cudaError_t e = cudaSuccess;
    (void)e;
    if (dtype == AIMB_BF16) {
        static bool once = false;
        if (!once) {
            e = cudaFuncSetAttribute(attn_temporal_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return AIMB_ERR_CUDA;
            once = true;
        }
        attn_temporal_fwd_kernel<bf16><<<grid, 128, smem, s>>>((const bf16*)qkv, (bf16*)o, B, T, n, heads);
    } else if (dtype == AIMB_F32) {
        static bool once = false;
        if (!once) {
            e = cudaFuncSetAttribute(attn_temporal_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return AIMB_ERR_CUDA;
            once = true;
        }
        attn_temporal_fwd_kernel<float><<<grid, 128, smem, s>>>((const float*)qkv, (float*)o, B, T, n, heads);
    } else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_attn_temporal_bwd(const void* qkv, const void* d_o, void* d_qkv, int32_t B, int32_t T, int32_t n,
                                      int32_t heads, int32_t dtype, void* stream) {
    if (!qkv || !d_o || !d_qkv || B < 0 || T <= 0 || T > 32 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = (size_t)4 * (4 * T * KS + 2 * T * (T + 1)) * 4;
    int64_t probs = (int64_t)B * n * heads;
    unsigned grid = (unsigned)((probs + 3) / 4);
    cudaError_t e = cudaSuccess;
    (void)e;
    if (dtype == AIMB_BF16) {
        static bool once = false;
        if (!once) {
            e = cudaFuncSetAttribute(attn_temporal_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return AIMB_ERR_CUDA;
            once = true;
        }
        attn_temporal_bwd_kernel<bf16><<<grid, 128, smem, s>>>((const bf16*)qkv, (const bf16*)d_o, (bf16*)d_qkv, B, T, n,
                                                               heads);
    } else if (dtype == AIMB_F32) {
        static bool once = false;
        if (!once) {
            e = cudaFuncSetAttribute(attn_temporal_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return AIMB_ERR_CUDA;
            once = true;
        }
        attn_temporal_bwd_kernel<float><<<grid, 128, smem, s>>>((const float*)qkv, (const float*)d_o, (float*)d_qkv, B, T,
                                                                n, heads);
    } else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

extern "C" int aimb_fork_weights(const void* qkv, const void* kc, float* w_o, float* w_c, int32_t frames, int32_t n,
                                 int32_t D, int32_t dtype, void* stream) {
    if (!qkv || !kc || !w_o || !w_c || frames < 0 || n <= 0 || D <= 0 || D > 2048) return AIMB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (frames == 0) return AIMB_OK;
    if (cudaMemsetAsync(w_o, 0, (size_t)frames * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (cudaMemsetAsync(w_c, 0, (size_t)frames * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    dim3 grid(frames, (n + 7) / 8);
    size_t smem = (size_t)8 * D * 4;
    if (dtype == AIMB_BF16)
        fork_weights_kernel<bf16><<<grid, 256, smem, s>>>((const bf16*)qkv, (const bf16*)kc, w_o, w_c, n, D);
    else if (dtype == AIMB_F32)
        fork_weights_kernel<float><<<grid, 256, smem, s>>>((const float*)qkv, (const float*)kc, w_o, w_c, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

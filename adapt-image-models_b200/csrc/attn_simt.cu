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
    pdl_grid_sync();
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
    pdl_grid_sync();
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

// ------------------------------------------------------------------ temporal attention: warp per (b, token, head)
// T_ (frames) is a template parameter (4/8/16/32).  Per warp: Q/K/V (and dO) rows of the sequence are staged in
// shared memory as fp32 [T][68] (16-byte aligned rows, 128-byte coalesced global reads of each row-head slice),
// the T x T scores are produced by lanes owning (i, j) pairs with float4 dot products, and the P.V products by
// lanes owning two head-dim columns.
constexpr int TKS = DH + 4;   // smem row stride (floats): keeps float4 alignment, conflict-free for the patterns used

template <typename T> __device__ __forceinline__ float2 ld_pair(const T* p);
template <> __device__ __forceinline__ float2 ld_pair<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }
template <> __device__ __forceinline__ float2 ld_pair<bf16>(const bf16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <typename T> __device__ __forceinline__ void st_pair(T* p, float a, float b);
template <> __device__ __forceinline__ void st_pair<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void st_pair<bf16>(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float dot64(const float* a, const float* b) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
        float4 x = *reinterpret_cast<const float4*>(a + d), y = *reinterpret_cast<const float4*>(b + d);
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    return acc;
}

template <typename T, int T_>
__global__ void __launch_bounds__(128) attn_temporal_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o, int B, int n,
                                                                int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) float sm[];
    constexpr int PS = T_ + 4;
    constexpr int UI = T_ <= 8 ? T_ : 2;      // outer-loop unroll: full only for short sequences (register pressure)
    constexpr int UP = T_ <= 8 ? (T_ * T_ + 31) / 32 : 1;
    constexpr int PER_WARP = 3 * T_ * TKS + T_ * PS;
    const int D = heads * DH, ld = 3 * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t prob = (int64_t)blockIdx.x * 4 + warp;
    if (prob >= (int64_t)B * n * heads) return;
    const int h = (int)(prob % heads);
    const int tok = (int)((prob / heads) % n);
    const int b = (int)(prob / ((int64_t)heads * n));
    float* sQ = sm + warp * PER_WARP;
    float* sK = sQ + T_ * TKS;
    float* sV = sK + T_ * TKS;
    float* sS = sV + T_ * TKS;
    const int64_t row0 = (int64_t)b * T_ * n + tok;   // row of frame t: row0 + t*n
    const T* base = qkv + row0 * ld + h * DH + 2 * lane;
    const int64_t rs = (int64_t)n * ld;
#pragma unroll
    for (int t = 0; t < T_; ++t) {
        const T* r = base + t * rs;
        *reinterpret_cast<float2*>(sQ + t * TKS + 2 * lane) = ld_pair<T>(r);
        *reinterpret_cast<float2*>(sK + t * TKS + 2 * lane) = ld_pair<T>(r + D);
        *reinterpret_cast<float2*>(sV + t * TKS + 2 * lane) = ld_pair<T>(r + 2 * D);
    }
    __syncwarp();
#pragma unroll UP
    for (int p = lane; p < T_ * T_; p += 32) {
        const int i = p / T_, j = p % T_;
        sS[i * PS + j] = dot64(sQ + i * TKS, sK + j * TKS) * ATT_SCALE;
    }
    __syncwarp();
    if (lane < T_) {
        float* row = sS + lane * PS;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < T_; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < T_; ++j) { float e = __expf(row[j] - mx); row[j] = e; sum += e; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < T_; ++j) row[j] *= inv;
    }
    __syncwarp();
    float2 vr[T_];
#pragma unroll
    for (int j = 0; j < T_; ++j) vr[j] = *reinterpret_cast<const float2*>(sV + j * TKS + 2 * lane);
    T* obase = o + row0 * D + h * DH + 2 * lane;
#pragma unroll UI
    for (int i = 0; i < T_; ++i) {
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int j = 0; j < T_; j += 4) {
            const float4 p = *reinterpret_cast<const float4*>(sS + i * PS + j);
            o0 = fmaf(p.x, vr[j].x, o0); o1 = fmaf(p.x, vr[j].y, o1);
            o0 = fmaf(p.y, vr[j + 1].x, o0); o1 = fmaf(p.y, vr[j + 1].y, o1);
            o0 = fmaf(p.z, vr[j + 2].x, o0); o1 = fmaf(p.z, vr[j + 2].y, o1);
            o0 = fmaf(p.w, vr[j + 3].x, o0); o1 = fmaf(p.w, vr[j + 3].y, o1);
        }
        st_pair<T>(obase + (int64_t)i * n * D, o0, o1);
    }
}

template <typename T, int T_>
__global__ void __launch_bounds__(128) attn_temporal_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ d_o,
                                                                T* __restrict__ d_qkv, int B, int n, int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) float sm[];
    constexpr int PS = T_ + 4;
    constexpr int UI = T_ <= 8 ? T_ : 2;      // outer-loop unroll: full only for short sequences (register pressure)
    constexpr int UP = T_ <= 8 ? (T_ * T_ + 31) / 32 : 1;
    constexpr int PER_WARP = 4 * T_ * TKS + 2 * T_ * PS;
    const int D = heads * DH, ld = 3 * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t prob = (int64_t)blockIdx.x * 4 + warp;
    if (prob >= (int64_t)B * n * heads) return;
    const int h = (int)(prob % heads);
    const int tok = (int)((prob / heads) % n);
    const int b = (int)(prob / ((int64_t)heads * n));
    float* sQ = sm + warp * PER_WARP;
    float* sK = sQ + T_ * TKS;
    float* sV = sK + T_ * TKS;
    float* sG = sV + T_ * TKS;            // dO
    float* sP = sG + T_ * TKS;            // probabilities
    float* sD = sP + T_ * PS;             // dP then dS
    const int64_t row0 = (int64_t)b * T_ * n + tok;
    const T* base = qkv + row0 * ld + h * DH + 2 * lane;
    const T* gbase = d_o + row0 * D + h * DH + 2 * lane;
    const int64_t rs = (int64_t)n * ld;
#pragma unroll
    for (int t = 0; t < T_; ++t) {
        const T* r = base + t * rs;
        *reinterpret_cast<float2*>(sQ + t * TKS + 2 * lane) = ld_pair<T>(r);
        *reinterpret_cast<float2*>(sK + t * TKS + 2 * lane) = ld_pair<T>(r + D);
        *reinterpret_cast<float2*>(sV + t * TKS + 2 * lane) = ld_pair<T>(r + 2 * D);
        *reinterpret_cast<float2*>(sG + t * TKS + 2 * lane) = ld_pair<T>(gbase + (int64_t)t * n * D);
    }
    __syncwarp();
#pragma unroll UP
    for (int p = lane; p < T_ * T_; p += 32) {
        const int i = p / T_, j = p % T_;
        sP[i * PS + j] = dot64(sQ + i * TKS, sK + j * TKS) * ATT_SCALE;
        sD[i * PS + j] = dot64(sG + i * TKS, sV + j * TKS);
    }
    __syncwarp();
    if (lane < T_) {
        float* row = sP + lane * PS;
        float* drow = sD + lane * PS;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < T_; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < T_; ++j) { float e = __expf(row[j] - mx); row[j] = e; sum += e; }
        const float inv = 1.f / sum;
        float delta = 0.f;
#pragma unroll
        for (int j = 0; j < T_; ++j) { row[j] *= inv; delta = fmaf(row[j], drow[j], delta); }
#pragma unroll
        for (int j = 0; j < T_; ++j) drow[j] = row[j] * (drow[j] - delta) * ATT_SCALE;
    }
    __syncwarp();
    T* dbase = d_qkv + row0 * ld + h * DH + 2 * lane;
    float2 reg[T_];
    // dQ_i = sum_j dS_ij K_j     (rows of dS)
#pragma unroll
    for (int j = 0; j < T_; ++j) reg[j] = *reinterpret_cast<const float2*>(sK + j * TKS + 2 * lane);
#pragma unroll UI
    for (int i = 0; i < T_; ++i) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < T_; j += 4) {
            const float4 c = *reinterpret_cast<const float4*>(sD + i * PS + j);
            a0 = fmaf(c.x, reg[j].x, a0); a1 = fmaf(c.x, reg[j].y, a1);
            a0 = fmaf(c.y, reg[j + 1].x, a0); a1 = fmaf(c.y, reg[j + 1].y, a1);
            a0 = fmaf(c.z, reg[j + 2].x, a0); a1 = fmaf(c.z, reg[j + 2].y, a1);
            a0 = fmaf(c.w, reg[j + 3].x, a0); a1 = fmaf(c.w, reg[j + 3].y, a1);
        }
        st_pair<T>(dbase + i * rs, a0, a1);
    }
    // dK_i = sum_j dS_ji Q_j     (columns of dS)
#pragma unroll
    for (int j = 0; j < T_; ++j) reg[j] = *reinterpret_cast<const float2*>(sQ + j * TKS + 2 * lane);
#pragma unroll UI
    for (int i = 0; i < T_; ++i) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < T_; ++j) {
            const float c = sD[j * PS + i];
            a0 = fmaf(c, reg[j].x, a0); a1 = fmaf(c, reg[j].y, a1);
        }
        st_pair<T>(dbase + i * rs + D, a0, a1);
    }
    // dV_i = sum_j P_ji dO_j     (columns of P)
#pragma unroll
    for (int j = 0; j < T_; ++j) reg[j] = *reinterpret_cast<const float2*>(sG + j * TKS + 2 * lane);
#pragma unroll UI
    for (int i = 0; i < T_; ++i) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < T_; ++j) {
            const float c = sP[j * PS + i];
            a0 = fmaf(c, reg[j].x, a0); a1 = fmaf(c, reg[j].y, a1);
        }
        st_pair<T>(dbase + i * rs + 2 * D, a0, a1);
    }
}

template <typename T, int T_>
static int temporal_fwd_launch(const void* qkv, void* o, int B, int n, int heads, cudaStream_t s) {
    constexpr size_t smem = (size_t)4 * (3 * T_ * TKS + T_ * (T_ + 4)) * 4;
    AIMB_SET_SMEM_ATTR((int)smem, attn_temporal_fwd_kernel<T, T_>);
    int64_t probs = (int64_t)B * n * heads;
    launch_k((attn_temporal_fwd_kernel<T, T_>), dim3((unsigned)((probs + 3) / 4)), dim3(128), smem, s, (const T*)qkv, (T*)o, B, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <typename T, int T_>
static int temporal_bwd_launch(const void* qkv, const void* d_o, void* d_qkv, int B, int n, int heads, cudaStream_t s) {
    constexpr size_t smem = (size_t)4 * (4 * T_ * TKS + 2 * T_ * (T_ + 4)) * 4;
    AIMB_SET_SMEM_ATTR((int)smem, attn_temporal_bwd_kernel<T, T_>);
    int64_t probs = (int64_t)B * n * heads;
    launch_k((attn_temporal_bwd_kernel<T, T_>), dim3((unsigned)((probs + 3) / 4)), dim3(128), smem, s, (const T*)qkv, (const T*)d_o, (T*)d_qkv, B, n,
                                                                                   heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <typename T>
static int temporal_fwd_T(const void* qkv, void* o, int B, int T_, int n, int heads, cudaStream_t s) {
    switch (T_) {
        case 4: return temporal_fwd_launch<T, 4>(qkv, o, B, n, heads, s);
        case 8: return temporal_fwd_launch<T, 8>(qkv, o, B, n, heads, s);
        case 16: return temporal_fwd_launch<T, 16>(qkv, o, B, n, heads, s);
        case 32: return temporal_fwd_launch<T, 32>(qkv, o, B, n, heads, s);
    }
    return AIMB_ERR_UNSUPPORTED;
}
template <typename T>
static int temporal_bwd_T(const void* qkv, const void* d_o, void* d_qkv, int B, int T_, int n, int heads, cudaStream_t s) {
    switch (T_) {
        case 4: return temporal_bwd_launch<T, 4>(qkv, d_o, d_qkv, B, n, heads, s);
        case 8: return temporal_bwd_launch<T, 8>(qkv, d_o, d_qkv, B, n, heads, s);
        case 16: return temporal_bwd_launch<T, 16>(qkv, d_o, d_qkv, B, n, heads, s);
        case 32: return temporal_bwd_launch<T, 32>(qkv, d_o, d_qkv, B, n, heads, s);
    }
    return AIMB_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------ fork block weights (vit_clip.py:147-151,182-186)
// w_o[f] = sum_{i,j} exp( q_i . k_j / 8 ) over the FULL width D (sum over heads of per-head logits),
// w_c[f] = sum_i exp( q_i . kc_f / 8 ).  fp32, no max subtraction (as the reference).
template <typename T>
__global__ void __launch_bounds__(256) fork_weights_kernel(const T* __restrict__ qkv, const T* __restrict__ kc,
                                                           float* __restrict__ w_o, float* __restrict__ w_c, int n,
                                                           int D) {
    pdl_grid_sync();
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
    AIMB_SET_SMEM_ATTR(227 * 1024, attn_spatial_fwd_simt<T>);
    launch_k((attn_spatial_fwd_simt<T>), dim3(frames * heads), dim3(256), smem, s, (const T*)qkv, (T*)o, lse, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <typename T>
int spatial_bwd_simt_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames,
                            int n, int heads, cudaStream_t s) {
    size_t smem = spatial_bwd_smem(n);
    if (smem > 227 * 1024) return AIMB_ERR_UNSUPPORTED;
    AIMB_SET_SMEM_ATTR(227 * 1024, attn_spatial_bwd_simt<T>);
    launch_k((attn_spatial_bwd_simt<T>), dim3(frames * heads), dim3(256), smem, s, (const T*)qkv, (const T*)o, (const T*)d_o, lse, (T*)d_qkv, n,
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

namespace aimb {
int attn_temporal_fwd_mma(const void* qkv, void* o, int B, int T, int n, int heads, cudaStream_t s);
int attn_temporal_bwd_mma(const void* qkv, const void* d_o, void* d_qkv, int B, int T, int n, int heads, cudaStream_t s);
}
static int g_temporal_simt = 0;   // tests / bench_tools: 1 forces the SIMT kernels for T = 8 too
extern "C" void aimb_debug_temporal_simt(int on) { g_temporal_simt = on; }

extern "C" int aimb_attn_temporal_fwd(const void* qkv, void* o, int32_t B, int32_t T, int32_t n, int32_t heads,
                                      int32_t dtype, void* stream) {
    if (!qkv || !o || B < 0 || T <= 0 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && ((T == 8 && heads % 2 == 0) || T == 16 || T == 32) && !g_temporal_simt) return attn_temporal_fwd_mma(qkv, o, B, T, n, heads, s);
    if (dtype == AIMB_BF16) return temporal_fwd_T<bf16>(qkv, o, B, T, n, heads, s);
    if (dtype == AIMB_F32) return temporal_fwd_T<float>(qkv, o, B, T, n, heads, s);
    return AIMB_ERR_ARG;
}

extern "C" int aimb_attn_temporal_bwd(const void* qkv, const void* d_o, void* d_qkv, int32_t B, int32_t T, int32_t n,
                                      int32_t heads, int32_t dtype, void* stream) {
    if (!qkv || !d_o || !d_qkv || B < 0 || T <= 0 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (B == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && ((T == 8 && heads % 2 == 0) || T == 16 || T == 32) && !g_temporal_simt) return attn_temporal_bwd_mma(qkv, d_o, d_qkv, B, T, n, heads, s);
    if (dtype == AIMB_BF16) return temporal_bwd_T<bf16>(qkv, d_o, d_qkv, B, T, n, heads, s);
    if (dtype == AIMB_F32) return temporal_bwd_T<float>(qkv, d_o, d_qkv, B, T, n, heads, s);
    return AIMB_ERR_ARG;
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
        launch_k((fork_weights_kernel<bf16>), dim3(grid), dim3(256), smem, s, (const bf16*)qkv, (const bf16*)kc, w_o, w_c, n, D);
    else if (dtype == AIMB_F32)
        launch_k((fork_weights_kernel<float>), dim3(grid), dim3(256), smem, s, (const float*)qkv, (const float*)kc, w_o, w_c, n, D);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

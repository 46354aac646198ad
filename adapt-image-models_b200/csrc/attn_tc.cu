// Spatial attention (vit_clip.py:140-156) on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma with
// the score matrix, the probabilities and the output accumulator all resident in tensor memory (TMEM).
//
// One problem = one (frame, head): Q, K, V are [n, 64] slices of the fused QKV buffer [frames * n, 3D] and are fetched
// in place by a 3-D tensor map {3D columns, n tokens, frames} (rows >= n of a box are zero-filled by the TMA unit, so
// padded keys / values are exact zeros).  n <= 256 (ViT-B/16: 197) runs here; larger n keeps the mma.sync kernel.
//
// Forward (attn_fwd_tc_kernel): persistent CTAs, 12 warps.
//   warp 0      TMA producer: Q (MT x 128 rows), K and V (NKP rows) of the next problem into a 2-stage ring
//   warp 1      one thread issues every tcgen05.mma:  S = Q_tile K^T (SS, 128 x NKP x 64) and O = P V (TS: the A
//               operand P is read from TMEM, B = V is an MN-major shared-memory operand)
//   warps 4-7   softmax warpgroup 0, warps 8-11 softmax warpgroup 1: thread = one query row (TMEM lane), so the row
//               max / row sum need no shuffles.  Pass 1 reads S for the max, pass 2 reads it again, exponentiates and
//               writes P (bf16, two keys per 32-bit column) over the S columns it has already consumed; O lands in
//               the (by then free) columns 128..191 of the same region, is normalised by 1 / rowsum and stored
//               head-major into o [M, D]; lse = max * scale + ln(rowsum) is kept for backward.
//   Units (problem, 128-row query tile) alternate between the two warpgroups, each owning one 256-column TMEM region:
//   while one group exponentiates, the other group's MMAs and output stores proceed.
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "ptx.cuh"

namespace aimb {
namespace atc {

constexpr int HD = 64;
constexpr int BOXR = 64;                  // TMA box: 64 rows x 64 bf16 columns (128-byte rows, 128B swizzle)
constexpr int BOXB = BOXR * 128;          // 8 KB
constexpr int QTILE_B = 128 * 128;        // one 128-row query tile
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
constexpr float SCALE = 0.125f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int FWD_THREADS = 384;

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// barrier block of the forward kernel
struct FwdBars {
    uint64_t qk_full[2], v_full[2], qk_empty[2], v_empty[2];
    uint64_t s_full[2], p_full[2], o_full[2], o_empty[2];
    uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, bf16* __restrict__ o, float* __restrict__ lse, const int n,
                   const int heads, const int D, const int nprob, const int MT, const int NKP, const int NST, long long* __restrict__ tl) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KB = (NKP + BOXR - 1) / BOXR;                 // 64-row boxes of K (and of V)
    const int STAGE_B = MT * QTILE_B + 2 * KB * BOXB;
    // bench_tools only: event timeline of CTA 0 (clock64 per unit and event), see bench_tools/attn_tc_timeline.py
    auto mark = [&](int u, int ev) { if (tl && blockIdx.x == 0 && (threadIdx.x & 31) == 0) tl[u * 16 + ev] = clock64(); };
    FwdBars* bars = reinterpret_cast<FwdBars*>(smem + NST * STAGE_B);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
    const int nloc = (nprob - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // problems of this CTA
    const int U = nloc * MT;                                                             // units of this CTA
    // TMEM columns: S / P of warpgroup g at g * RS; O in its own 64 columns when both score regions leave room for it
    // (then the next S may be issued without waiting for the output to be read), else in columns 128..191 of the region
    const bool OSEP = 2 * NKP + HD <= 512;
    const int RS = OSEP ? NKP : 256;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tm);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&bars->qk_full[i], 1); ptx::mbar_init(&bars->v_full[i], 1);
            ptx::mbar_init(&bars->qk_empty[i], 1); ptx::mbar_init(&bars->v_empty[i], 1);
            ptx::mbar_init(&bars->s_full[i], 1); ptx::mbar_init(&bars->p_full[i], 4);
            ptx::mbar_init(&bars->o_full[i], 1); ptx::mbar_init(&bars->o_empty[i], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<512>(&bars->tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_ptr, 0);      // warp-uniform for the MMA issue path
    pdl_wait();
    if (warp == 0) {
        // ---- TMA producer (all lanes run the loop, one elected lane issues: see ptx.cuh "warp-uniform issue")
        for (int k = 0; k < nloc; ++k) {
            const int pi = blockIdx.x + k * gridDim.x;
            const int f = pi / heads, h = pi - f * heads;
            const int st = k % NST;
            const uint32_t r = (uint32_t)(k / NST) & 1u;
            uint8_t* sq = smem + st * STAGE_B;
            uint8_t* sk = sq + MT * QTILE_B;
            uint8_t* sv = sk + KB * BOXB;
            ptx::mbar_wait(&bars->qk_empty[st], r ^ 1u);
            ptx::mbar_arrive_expect_tx_e(&bars->qk_full[st], (uint32_t)((2 * MT + KB) * BOXB));
            for (int b = 0; b < KB; ++b) ptx::tma_load_3d_e(sk + b * BOXB, &tm, &bars->qk_full[st], D + h * HD, b * BOXR, f);
            for (int b = 0; b < 2 * MT; ++b) ptx::tma_load_3d_e(sq + b * BOXB, &tm, &bars->qk_full[st], h * HD, b * BOXR, f);
            ptx::mbar_wait(&bars->v_empty[st], r ^ 1u);
            ptx::mbar_arrive_expect_tx_e(&bars->v_full[st], (uint32_t)(KB * BOXB));
            for (int b = 0; b < KB; ++b) ptx::tma_load_3d_e(sv + b * BOXB, &tm, &bars->v_full[st], 2 * D + h * HD, b * BOXR, f);
        }
    } else if (warp == 1) {
        // ---- MMA issuer (warp-uniform)
        const uint32_t idesc_s = ptx::umma_idesc_bf16(128, NKP);
        const uint32_t idesc_o = ptx::umma_idesc_bf16(128, HD, 0, 1);
        const int KS = NKP / 16;
        const uint32_t smem0 = ptx::smem_u32(smem);
        auto issue_s = [&](int u) {
            const int k = u / MT, mt = u - k * MT, st = k % NST, g = u & 1, it = u >> 1;
            if (!OSEP && it > 0) ptx::mbar_wait(&bars->o_empty[g], (uint32_t)(it - 1) & 1u);   // O(u-2) lives inside this region
            if (mt == 0) ptx::mbar_wait(&bars->qk_full[st], (uint32_t)(k / NST) & 1u);
            ptx::tc_fence_after();
            const uint32_t sq = smem0 + st * STAGE_B;
            const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sq + mt * QTILE_B);
            const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sq + MT * QTILE_B);
            const uint32_t d_tmem = tmem_base + g * RS;
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) ptx::umma_bf16_e(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc_s, kk ? 1u : 0u);
            ptx::umma_commit_e(&bars->s_full[g]);
            mark(u, 10);
            if (mt == MT - 1) ptx::umma_commit_e(&bars->qk_empty[st]);
        };
        auto issue_pv = [&](int u) {
            const int k = u / MT, mt = u - k * MT, st = k % NST, g = u & 1, it = u >> 1;
            ptx::mbar_wait(&bars->p_full[g], (uint32_t)it & 1u);
            mark(u, 8);
            if (OSEP && u > 0) ptx::mbar_wait(&bars->o_empty[g ^ 1], (uint32_t)((u - 1) >> 1) & 1u);   // shared O columns drained
            if (mt == 0) ptx::mbar_wait(&bars->v_full[st], (uint32_t)(k / NST) & 1u);
            ptx::tc_fence_after();
            const uint64_t vdesc = ptx::umma_desc_mnmajor_sw128(smem0 + st * STAGE_B + MT * QTILE_B + KB * BOXB, BOXB);
            const uint32_t p_tmem = tmem_base + g * RS;
            const uint32_t d_tmem = OSEP ? tmem_base + 448 : p_tmem + 128;
#pragma unroll 4
            for (int j = 0; j < KS; ++j)       // 16 keys per MMA: +8 TMEM columns of P, +2048 B of V (two 8-row swizzle atoms)
                ptx::umma_bf16_ts_e(d_tmem, p_tmem + j * 8, vdesc + (uint64_t)(j * 128), idesc_o, j ? 1u : 0u);
            ptx::umma_commit_e(&bars->o_full[g]);
            mark(u, 9);
            if (mt == MT - 1) ptx::umma_commit_e(&bars->v_empty[st]);
        };
        if (U > 0) issue_s(0);
        if (U > 1) issue_s(1);
        for (int u = 0; u < U; ++u) {
            issue_pv(u);
            if (u + 2 < U) issue_s(u + 2);
        }
    } else if (warp >= 4) {
        const int g = (warp - 4) >> 2, wq = warp & 3;
        const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + g * RS;
        const uint32_t orow = tmem_base + ((uint32_t)(wq * 32) << 16) + (OSEP ? 448 : g * RS + 128);
        const int NCH = (n + 31) >> 5;
        for (int it = 0;; ++it) {
            const int u = 2 * it + g;
            if (u >= U) break;
            const int k = u / MT, mt = u - k * MT;
            const int pi = blockIdx.x + k * gridDim.x;
            const int f = pi / heads, h = pi - f * heads;
            const int row = mt * 128 + wq * 32 + lane;
            const bool active = (mt * 128 + wq * 32) < n;        // warp-uniform: does this warp own any real query row
            ptx::mbar_wait(&bars->s_full[g], (uint32_t)it & 1u);
            ptx::tc_fence_after();
            if (wq == 0) mark(u, 0);
            float mx = -INFINITY, sum = 0.f;
            uint32_t va[32], vb[32];
            if (active) {
                // ---- pass 1: row maximum (tail chunk: only the 8-column groups that hold real keys)
                auto rmax = [&](const uint32_t (&v)[32], int c) {
                    if ((c + 1) * 32 <= n) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (c * 32 + q * 8 < n) {
#pragma unroll
                                for (int j = q * 8; j < q * 8 + 8; ++j) if (c * 32 + j < n) mx = fmaxf(mx, __uint_as_float(v[j]));
                            }
                        }
                    }
                };
                ptx::tmem_ld_32x32b_x32(trow, va);
                for (int c = 0; c < NCH; c += 2) {
                    ptx::tmem_wait_ld();
                    if (c + 1 < NCH) ptx::tmem_ld_32x32b_x32(trow + (c + 1) * 32, vb);
                    rmax(va, c);
                    if (c + 1 < NCH) {
                        ptx::tmem_wait_ld();
                        if (c + 2 < NCH) ptx::tmem_ld_32x32b_x32(trow + (c + 2) * 32, va);
                        rmax(vb, c + 1);
                    }
                }
                if (wq == 0) mark(u, 1);
                // ---- pass 2: p = exp2((s - max) * scale * log2 e); P (bf16) overwrites the consumed S columns
                const float mb = mx * SCALE_LOG2;
                uint32_t pk[16];
                auto expo = [&](const uint32_t (&v)[32], int c) {
                    const bool tail = (c + 1) * 32 > n;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (tail && c * 32 + q * 8 >= n) {           // warp-uniform: nothing but padding in this group
#pragma unroll
                            for (int j = 0; j < 4; ++j) pk[q * 4 + j] = 0u;
                            continue;
                        }
#pragma unroll
                        for (int j = q * 4; j < q * 4 + 4; ++j) {
                            float p0 = ex2_ftz(fmaf(__uint_as_float(v[2 * j]), SCALE_LOG2, -mb));
                            float p1 = ex2_ftz(fmaf(__uint_as_float(v[2 * j + 1]), SCALE_LOG2, -mb));
                            if (tail) {
                                if (c * 32 + 2 * j >= n) p0 = 0.f;
                                if (c * 32 + 2 * j + 1 >= n) p1 = 0.f;
                            }
                            sum += p0 + p1;
                            pk[j] = pack_bf16(p0, p1);
                        }
                    }
                    ptx::tmem_st_32x32b_x16(trow + c * 16, pk);
                };
                ptx::tmem_ld_32x32b_x32(trow, va);
                for (int c = 0; c < NCH; c += 2) {
                    ptx::tmem_wait_ld();
                    if (c + 1 < NCH) ptx::tmem_ld_32x32b_x32(trow + (c + 1) * 32, vb);
                    expo(va, c);
                    if (c + 1 < NCH) {
                        ptx::tmem_wait_ld();
                        if (c + 2 < NCH) ptx::tmem_ld_32x32b_x32(trow + (c + 2) * 32, va);
                        expo(vb, c + 1);
                    }
                }
                ptx::tmem_wait_st();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->p_full[g]);
            if (wq == 0) mark(u, 2);
            // ---- O = P V
            ptx::mbar_wait(&bars->o_full[g], (uint32_t)it & 1u);
            ptx::tc_fence_after();
            if (wq == 0) mark(u, 3);
            if (active) {
                ptx::tmem_ld_32x32b_x32(orow, va);
                ptx::tmem_ld_32x32b_x32(orow + 32, vb);
                ptx::tmem_wait_ld();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->o_empty[g]);      // the output columns may be rewritten while we store
            if (wq == 0) mark(u, 4);
            if (active && row < n) {
                const float inv = __fdividef(1.f, sum);
                uint32_t ob[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    ob[j] = pack_bf16(__uint_as_float(va[2 * j]) * inv, __uint_as_float(va[2 * j + 1]) * inv);
                    ob[16 + j] = pack_bf16(__uint_as_float(vb[2 * j]) * inv, __uint_as_float(vb[2 * j + 1]) * inv);
                }
                bf16* op = o + ((int64_t)f * n + row) * D + h * HD;
#pragma unroll
                for (int q = 0; q < 4; ++q) st_global_v8(op + q * 16, ob + q * 8);
                if (lse) lse[((int64_t)f * heads + h) * n + row] = mx * SCALE + __logf(sum);
            }
            if (wq == 0) mark(u, 5);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct Key3 {
    const void* ptr; int64_t slabs, rows, cols;
    bool operator==(const Key3& o) const { return ptr == o.ptr && slabs == o.slabs && rows == o.rows && cols == o.cols; }
};
struct Key3Hash {
    size_t operator()(const Key3& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<int64_t>()(k.slabs);
        h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
        h = h * 1000003u ^ std::hash<int64_t>()(k.cols);
        return h;
    }
};

// bf16 [slabs, rows, cols] contiguous; box = 64 columns x 64 rows x 1 slab, 128B swizzle, zero fill out of bounds
static int make_tmap3(CUtensorMap* out, const void* ptr, int64_t slabs, int64_t rows, int64_t cols) {
    static std::mutex mu;
    static std::unordered_map<Key3, CUtensorMap, Key3Hash> cache;
    Key3 key{ptr, slabs, rows, cols};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return AIMB_OK; }
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return AIMB_ERR_DRIVER;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
    cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)BOXR, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap tmap;
    CUresult r = fn(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return AIMB_ERR_DRIVER;
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 4096) cache.clear();
        cache[key] = tmap;
    }
    *out = tmap;
    return AIMB_OK;
}

static int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    static int cached[64] = {0};
    if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

}  // namespace atc

static long long* g_attn_timeline = nullptr;
void attn_tc_set_timeline(long long* p) { g_attn_timeline = p; }

bool attn_spatial_tc_supported(int n, int heads) { return n >= 1 && n <= 256 && heads >= 1; }

int attn_spatial_fwd_tc(const void* qkv, void* o, float* lse, int frames, int n, int heads, cudaStream_t s) {
    using namespace atc;
    if (!attn_spatial_tc_supported(n, heads)) return AIMB_ERR_UNSUPPORTED;
    const int D = heads * HD;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)o & 31)) return AIMB_ERR_ARG;
    CUtensorMap tm;
    int rc = make_tmap3(&tm, qkv, frames, n, 3 * (int64_t)D);
    if (rc) return rc;
    const int MT = (n + 127) / 128;
    const int NKP = (n + 15) & ~15;
    const int KB = (NKP + BOXR - 1) / BOXR;
    const int STAGE_B = MT * QTILE_B + 2 * KB * BOXB;
    const int NST = (2 * STAGE_B + 2048 <= 227 * 1024) ? 2 : 1;
    const int smem = NST * STAGE_B + 1024 + 256;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return AIMB_ERR_UNSUPPORTED;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return AIMB_ERR_CUDA;
        attr_set[dev] = true;
    }
    const int nprob = frames * heads;
    const int sms = sm_count();
    const int waves = (nprob + sms - 1) / sms;
    const int grid = (nprob + waves - 1) / waves;
    launch_k(attn_fwd_tc_kernel, dim3(grid), dim3(FWD_THREADS), (size_t)smem, s, tm, (bf16*)o, lse, n, heads, D, nprob, MT, NKP, NST, g_attn_timeline);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

}  // namespace aimb

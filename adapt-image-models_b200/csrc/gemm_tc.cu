// bf16 GEMM  C[M,N] = epilogue(A[M,K] * W[N,K]^T)  on the 5th-gen tensor cores:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring -> tcgen05.mma (one issuing thread,
//   fp32 accumulators in TMEM, double-buffered) -> tcgen05.ld -> fused epilogue -> global.
// Persistent: one CTA per SM loops over 128 x BN output tiles.  Warp roles:
//   warp 0 : TMA producer          warp 1 : TMEM alloc + MMA issuer
//   warps 2-9 : epilogue (warp%4 = the 32-lane TMEM quadrant it may read; two warps per quadrant split the columns)
// This one kernel carries every nn.Linear of the block (vit_clip.py:93-97, 132-138, 157, 60-69)
// in forward and the dgrad GEMMs in backward; `nn.Linear` weights [out,in] are already the
// K-major B operand, so no transposes are needed in forward.
#include <mutex>
#include <unordered_map>
#include "common.cuh"
#include "ptx.cuh"

namespace aimb {

int gemm_simt_launch(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sn, int64_t b_sk,
                     const EpiParams& epi, int64_t M, int N, int K, int dtype, cudaStream_t s);

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 192;          // wgrad kernel: 1 producer + 1 MMA + 4 epilogue warps
constexpr int GEMM_THREADS = 320;        // GEMM: 1 producer + 1 MMA + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int STG_BYTES = EPI_WARPS * 32 * 33 * 4;   // per-warp 32x32 fp32 transpose buffers (padded)
constexpr int A_STAGE_BYTES = BM * BK * 2;

template <int BN> struct TileCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES_RAW = (232448 - STG_BYTES - 1024 - 256) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---- vectorised bf16 epilogue: 8 consecutive columns (16 bytes) of one row -----------------------------
__device__ __forceinline__ void ld8_bf16(const bf16* p, float* v) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}
__device__ __forceinline__ void st8_bf16(bf16* p, const float* v) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = t;
}

// v[8] = accumulators of row m, columns n0..n0+7; bias8 = bias of those columns (already loaded).
__device__ __forceinline__ void epilogue_vec8(const EpiParams& e, int64_t m, int n0, float* v, const float* bias8) {
    float rs = 1.f;
    if (e.row_scale) rs = e.row_scale[m % e.row_mod];
    const int64_t off = m * e.ldo + n0;
    float t[8];
    if (e.bias) {
        const float bs = e.bias_rowscaled ? rs : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(bias8[j], bs, v[j]);
    }
    if (e.out_pre) {
        st8_bf16((bf16*)e.out_pre + off, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = roundT<bf16>(v[j]);
    }
    if (e.act != AIMB_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act(e.act, v[j]);
    }
    if (e.dact_src) {
        ld8_bf16((const bf16*)e.dact_src + off, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= apply_act_grad(e.dact, t[j]);
    }
    const float sc = e.alpha * ((e.row_scale && !e.bias_rowscaled) ? rs : 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= sc;
    if (e.res1) {
        ld8_bf16((const bf16*)e.res1 + off, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if (e.res2) {
        ld8_bf16((const bf16*)e.res2 + off, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    st8_bf16((bf16*)e.out + off, v);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams epi,
               const int M, const int N, const int K) {
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + STG_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = N / BN;
    const int m_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * m_tiles;
    const int KB = K / BK;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], EPI_WARPS); }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sa);
                    const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128 B swizzle row: +2 in 16-byte units
                        ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (kb == KB - 1) ptx::umma_commit(&tfull_bar[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // 8 epilogue warps: quad = TMEM lane quadrant this warp may read, half = which half of the BN columns.
        // Accumulators are transposed through a private 32x33 fp32 smem tile so that global traffic
        // (residual / saved-activation loads, output stores) is 16 B per lane, 8 rows x 64 B per instruction.
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = stg_all + (warp - 2) * (32 * 33);
        const int row_l = lane >> 2, c0 = (lane & 3) * 8;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            const int64_t row_base = (int64_t)m_blk * BM + quad * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr + c, r);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                const int n0 = n_blk * BN + c + c0;
                float bias8[8];
                if (epi.bias) ld8_bf16((const bf16*)epi.bias + n0, bias8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = i * 8 + row_l;
                    const int64_t row = row_base + rl;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = stg[rl * 33 + c0 + j];
                    if (row < M) epilogue_vec8(epi, row, n0, v, bias8);
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

struct TmKey {
    const void* ptr; int64_t rows, cols, ld; int box_rows;
    bool operator==(const TmKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
        h = h * 1000003u ^ std::hash<int64_t>()(k.cols);
        h = h * 1000003u ^ std::hash<int64_t>()(k.ld);
        h = h * 1000003u ^ std::hash<int>()(k.box_rows);
        return h;
    }
};

// 2-D bf16 row-major [rows, cols] (row stride ld elements); box = 64 columns x box_rows rows, 128B swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> cache;
    TmKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return AIMB_OK; }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return AIMB_ERR_DRIVER;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return AIMB_ERR_DRIVER;
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 65536) cache.clear();
        cache[key] = tm;
    }
    *out = tm;
    return AIMB_OK;
}

static int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, int M, int N, int K, cudaStream_t s) {
    using Cfg = TileCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return AIMB_ERR_CUDA;
        attr_set = true;
    }
    int total = (N / BN) * ((M + BM - 1) / BM);
    int grid = total < num_sms() ? total : num_sms();
    gemm_tc_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, s>>>(ta, tb, p, M, N, K);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

// Pick the N tile that minimises (waves x per-tile cost) over the persistent grid.
static int pick_bn(int64_t M, int N) {
    const int cand[4] = {256, 192, 128, 64};
    int best = 0; double best_cost = 1e30;
    int64_t mt = (M + BM - 1) / BM;
    for (int i = 0; i < 4; ++i) {
        int bn = cand[i];
        if (N % bn) continue;
        int64_t tiles = mt * (N / bn);
        int64_t waves = (tiles + num_sms() - 1) / num_sms();
        double per_tile = (double)(bn < 128 ? 128 : bn) + 16.0;   // below N=128 the A-operand traffic dominates
        double cost = (double)waves * per_tile;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

int gemm_tc_launch(const void* A, int64_t lda, const void* W, int64_t ldw, const EpiParams& p, int64_t M, int N, int K,
                   int force_bn, cudaStream_t s) {
    if (K % BK || N % 64 || (lda % 8) || (ldw % 8) || ((uintptr_t)A & 15) || ((uintptr_t)W & 15)) return AIMB_ERR_ARG;
    if (p.ldo % 8 || ((uintptr_t)p.out & 15)) return AIMB_ERR_ARG;
    if (M >= (1ll << 31)) return AIMB_ERR_ARG;
    int bn = force_bn > 0 ? force_bn : pick_bn(M, N);
    if (bn == 0 || N % bn) return AIMB_ERR_ARG;
    CUtensorMap ta, tb;
    int rc = make_tmap_bf16(&ta, A, M, K, lda, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb, W, N, K, ldw, bn);
    if (rc) return rc;
    switch (bn) {
        case 256: return launch_tc<256>(ta, tb, p, (int)M, N, K, s);
        case 192: return launch_tc<192>(ta, tb, p, (int)M, N, K, s);
        case 128: return launch_tc<128>(ta, tb, p, (int)M, N, K, s);
        case 64: return launch_tc<64>(ta, tb, p, (int)M, N, K, s);
    }
    return AIMB_ERR_ARG;
}


// ---------------------------------------------------------------------------------------- adapter wgrad
// out[128-tile of P's columns, NS] (+)= alpha * P[R, CB]^T * Q[R, NS]   (contraction over the R rows).
// Both operands are MN-major for the tensor core (rows = contraction index): TMA drops [64 rows][64 cols]
// boxes straight from the row-major activations, no transposes.  grid = (CB/128, splits): each CTA reduces
// its slice of R into TMEM and adds the 128 x NS partial into the fp32 gradient with red.global.
template <int NS>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, float* __restrict__ out,
                const int64_t ldo, const int transposed_out, const float alpha, const int KB, const int kb_per_split) {
    constexpr int BOX = 64 * 64 * 2;                 // one [64 r][64 c] bf16 box
    constexpr int A_BYTES = 2 * BOX, B_BYTES = (NS / 64) * BOX, STG = A_BYTES + B_BYTES;
    constexpr int STAGES = (200 * 1024) / STG > 6 ? 6 : (200 * 1024) / STG;
    constexpr int TCOLS = NS <= 64 ? 64 : NS <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STG);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x;
    const int kb0 = blockIdx.y * kb_per_split;
    const int kb1 = (kb0 + kb_per_split < KB) ? kb0 + kb_per_split : KB;
    if (kb0 >= kb1) return;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmP);
        ptx::prefetch_tmap(&tmQ);
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(done_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<TCOLS>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx(&full_bar[stage], STG);
                uint8_t* sa = smem + stage * STG;
                ptx::tma_load_2d(sa, &tmP, &full_bar[stage], mt * 128, kb * 64);
                ptx::tma_load_2d(sa + BOX, &tmP, &full_bar[stage], mt * 128 + 64, kb * 64);
#pragma unroll
                for (int j = 0; j < NS / 64; ++j)
                    ptx::tma_load_2d(sa + A_BYTES + j * BOX, &tmQ, &full_bar[stage], j * 64, kb * 64);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, NS, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * STG);
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 16 contraction rows per MMA = two 8-row swizzle atoms = 2048 B
                    const uint64_t adesc = ptx::umma_desc_mnmajor_sw128(sa + k * 2048, BOX);
                    const uint64_t bdesc = ptx::umma_desc_mnmajor_sw128(sa + A_BYTES + k * 2048, BOX);
                    ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit(&empty_bar[stage]);
                if (kb == kb1 - 1) ptx::umma_commit(done_bar);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        const int m = mt * 128 + quad * 32 + lane;          // column of P == row of the (untransposed) result
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < NS; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + c, r);
            ptx::tmem_wait_ld();
            if (transposed_out) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(out + (int64_t)(c + j) * ldo + m, alpha * __uint_as_float(r[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(out + (int64_t)m * ldo + c + j, alpha * __uint_as_float(r[j]));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<TCOLS>(tmem_base);
    }
}

template <int NS>
static int launch_wgrad(const CUtensorMap& tp, const CUtensorMap& tq, float* out, int64_t ldo, int transposed, float alpha,
                        int CB, int R, cudaStream_t s) {
    constexpr int BOX = 64 * 64 * 2;
    constexpr int STG = 2 * BOX + (NS / 64) * BOX;
    constexpr int STAGES = (200 * 1024) / STG > 6 ? 6 : (200 * 1024) / STG;
    constexpr int SMEM = STAGES * STG + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
            return AIMB_ERR_CUDA;
        attr_set = true;
    }
    int KB = (R + 63) / 64;
    int mtiles = CB / 128;
    int splits = num_sms() / mtiles;
    if (splits < 1) splits = 1;
    if (splits > KB) splits = KB;
    int per = (KB + splits - 1) / splits;
    splits = (KB + per - 1) / per;
    dim3 grid(mtiles, splits);
    wgrad_tc_kernel<NS><<<grid, TC_THREADS, SMEM, s>>>(tp, tq, out, ldo, transposed, alpha, KB, per);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

// dW[N, K] (+)= alpha * dY[R, N]^T X[R, K].  The wide operand (a multiple of 128 columns) becomes the UMMA M
// side, the narrow one (<= 256 columns, multiple of 64) the N side; the result is stored transposed if needed.
int wgrad_tc_launch(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t R, int N, int K, float alpha,
                    cudaStream_t s) {
    const void *P, *Q; int64_t ldp, ldq; int CB, NS, transposed;
    if (N % 128 == 0 && K % 64 == 0 && K <= 256) { P = dY; ldp = ldy; CB = N; Q = X; ldq = ldx; NS = K; transposed = 0; }
    else if (K % 128 == 0 && N % 64 == 0 && N <= 256) { P = X; ldp = ldx; CB = K; Q = dY; ldq = ldy; NS = N; transposed = 1; }
    else return AIMB_ERR_UNSUPPORTED;
    if ((ldp % 8) || (ldq % 8) || ((uintptr_t)P & 15) || ((uintptr_t)Q & 15) || R >= (1ll << 31)) return AIMB_ERR_UNSUPPORTED;
    CUtensorMap tp, tq;
    int rc = make_tmap_bf16(&tp, P, R, CB, ldp, 64);
    if (rc) return rc;
    rc = make_tmap_bf16(&tq, Q, R, NS, ldq, 64);
    if (rc) return rc;
    const int64_t ldo = K;
    switch (NS) {
        case 64: return launch_wgrad<64>(tp, tq, dW, ldo, transposed, alpha, CB, (int)R, s);
        case 128: return launch_wgrad<128>(tp, tq, dW, ldo, transposed, alpha, CB, (int)R, s);
        case 192: return launch_wgrad<192>(tp, tq, dW, ldo, transposed, alpha, CB, (int)R, s);
        case 256: return launch_wgrad<256>(tp, tq, dW, ldo, transposed, alpha, CB, (int)R, s);
    }
    return AIMB_ERR_UNSUPPORTED;
}

}  // namespace aimb

using namespace aimb;

static int g_force_bn = 0;
extern "C" void aimb_debug_force_bn(int bn) { g_force_bn = bn; }

extern "C" int aimb_gemm_nt(const void* A, int64_t lda, const void* W, int64_t ldw, const aimb_epilogue_t* epi, int64_t M,
                            int32_t N, int32_t K, int32_t dtype, int32_t impl, void* stream) {
    if (!A || !W || !epi || !epi->out || M < 0 || N <= 0 || K <= 0 || lda < K || ldw < K) return AIMB_ERR_ARG;
    if (epi->accumulate && !epi->out_f32) return AIMB_ERR_ARG;
    if (epi->bias_rowscaled && !epi->row_scale) return AIMB_ERR_ARG;
    if (M == 0) return AIMB_OK;
    EpiParams p = make_epi(epi, N);
    cudaStream_t s = (cudaStream_t)stream;
    // tcgen05 kernel for every shape it tiles (all ViT-B/16 and ViT-L/14 GEMMs); shapes it cannot tile
    // (N or K not a multiple of 64 — toy widths only) run on the SIMT kernel, still on the GPU.
    const bool tc_ok = (K % BK == 0) && (N % 64 == 0) && (lda % 8 == 0) && (ldw % 8 == 0) && (p.ldo % 8 == 0);
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && !p.out_f32 && tc_ok)
        return gemm_tc_launch(A, lda, W, ldw, p, M, N, K, g_force_bn, s);
    if (dtype != AIMB_BF16 && dtype != AIMB_F32) return AIMB_ERR_ARG;
    return gemm_simt_launch(A, lda, 1, W, ldw, 1, p, M, N, K, dtype, s);
}

extern "C" int aimb_gemm_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t R, int32_t N,
                               int32_t K, float alpha, int32_t accumulate, int32_t dtype, int32_t impl, void* stream) {
    if (!dY || !X || !dW || R < 0 || N <= 0 || K <= 0 || ldy < N || ldx < K) return AIMB_ERR_ARG;
    cudaStream_t s0 = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && R > 0) {
        const bool shape_ok = (N % 128 == 0 && K % 64 == 0 && K <= 256) || (K % 128 == 0 && N % 64 == 0 && N <= 256);
        if (shape_ok && ldy % 8 == 0 && ldx % 8 == 0) {
            if (!accumulate && cudaMemsetAsync(dW, 0, (size_t)N * K * 4, s0) != cudaSuccess) return AIMB_ERR_CUDA;
            return wgrad_tc_launch(dY, ldy, X, ldx, dW, R, N, K, alpha, s0);
        }
    }
    EpiParams p{};
    p.out = dW; p.alpha = alpha; p.row_mod = 1; p.out_f32 = 1; p.accumulate = accumulate; p.ldo = K;
    cudaStream_t s = (cudaStream_t)stream;
    if (R == 0) {
        if (!accumulate && cudaMemsetAsync(dW, 0, (size_t)N * K * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
        return AIMB_OK;
    }
    // C[n, k] = sum_r dY[r, n] * X[r, k]:  "M" = N rows (stride 1 over n, ldy over r), "N" = K
    return gemm_simt_launch(dY, 1, ldy, X, 1, ldx, p, N, K, (int)R, dtype, s);
}

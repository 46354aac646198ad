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
#include <type_traits>
#include <unordered_map>
#include "common.cuh"
#include "ptx.cuh"
#include "gemm_shared.cuh"

namespace aimb {

int gemm_simt_launch(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sn, int64_t b_sk,
                     const EpiParams& epi, int64_t M, int N, int K, int dtype, cudaStream_t s);


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

// Global operands of the epilogue (saved pre-activation for dact, residuals) for one 32x32 chunk, as seen by
// one lane: 4 row-groups x 16 bytes each.  They do not depend on the accumulator, so they are fetched one
// chunk AHEAD (and before the accumulator-ready wait for the first chunk): their latency never sits on the
// epilogue's critical path.
#ifdef AIMB_DEBUG_EPILOGUE            // bench_tools stage attribution only (nvcc -DAIMB_DEBUG_EPILOGUE); release kernels carry no switch
__device__ int g_dbg_skip_epilogue = 0;   // 1 = drain TMEM but skip the epilogue math / global traffic, 4-6: see gemm_tc4_kernel
#else
constexpr int g_dbg_skip_epilogue = 0;
#endif

// Column-sum flush of one tile: `ncols` per-CTA partials from shared memory into the global fp32 vector.  Every row tile of
// the GEMM adds into the same addresses (99 CTAs at M = 12 608) and same-address atomics serialise in L2, so 4 columns go
// out as ONE 16-byte vector atomic (scalar fallback for an unaligned destination).  Called by the first `ncols` epilogue threads.
__device__ __forceinline__ void colsum_flush(float* __restrict__ dst, float* scol, int te, int ncols) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        if (te < ncols / 4) {
            const float4 v = *reinterpret_cast<const float4*>(scol + 4 * te);
            atomicAdd(reinterpret_cast<float4*>(dst) + te, v);
            *reinterpret_cast<float4*>(scol + 4 * te) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (te < ncols) {
        atomicAdd(dst + te, scol[te]);
        scol[te] = 0.f;
    }
}

struct EpiExt {
    uint4 d[4], r1[4], r2[4];
};

// EXT bit mask (compile time): 1 = dact_src, 2 = res1, 4 = res2 — unused slots cost no registers.
template <int EXT>
__device__ __forceinline__ void epi_prefetch(const EpiParams& e, EpiExt& x, int64_t row_base, int row_l, int n0, int M) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row_base + i * 8 + row_l;
        if (row < M) {
            const int64_t off = row * e.ldo + n0;
            if ((EXT & 1) && (EXT != 7 || e.dact_src)) x.d[i] = *reinterpret_cast<const uint4*>((const bf16*)e.dact_src + off);
            if ((EXT & 2) && (EXT != 7 || e.res1)) x.r1[i] = *reinterpret_cast<const uint4*>((const bf16*)e.res1 + off);
            if ((EXT & 4) && (EXT != 7 || e.res2)) x.r2[i] = *reinterpret_cast<const uint4*>((const bf16*)e.res2 + off);
        }
    }
}

// v[8] = accumulators of row m, columns n0..n0+7; bias8 = bias of those columns; x/i = prefetched operands.
// ACT / DACT are compile-time: only the activation actually used is in the instruction stream.
template <int ACT, int DACT, int EXT, typename XT>
__device__ __forceinline__ void epilogue_vec8(const EpiParams& e, int64_t m, int n0, float* v, const float* bias8,
                                              const XT& x, int i) {
    float rs = 1.f;
    if (e.row_scale) rs = e.row_scale[m % e.row_mod];
    const int64_t off = m * e.ldo + n0;
    float t[8];
    if (e.ln_mean) {       // folded LayerNorm (generic kernel only; the row-layout kernel has its own variant 8)
        const float mu = e.ln_mean[m], rsd = e.ln_rstd[m];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (v[j] - mu * e.ln_wsum[n0 + j]) * rsd;
    }
    if (e.bias) {
        const float bs = e.bias_rowscaled ? rs : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(bias8[j], bs, v[j]);
    }
    if (e.out_pre) {     // the activation sees the bf16-rounded pre-activation, exactly what backward will read
        uint4 pk;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            pw[j] = *reinterpret_cast<uint32_t*>(&h2);
            v[2 * j] = __uint_as_float(pw[j] << 16);
            v[2 * j + 1] = __uint_as_float(pw[j] & 0xffff0000u);
        }
        *reinterpret_cast<uint4*>((bf16*)e.out_pre + off) = pk;
    }
    if (ACT != AIMB_ACT_NONE && (ACT > 0 || e.act != AIMB_ACT_NONE)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_fn<ACT>(v[j], e.act);
    }
    if ((EXT & 1) && DACT != AIMB_ACT_NONE && (EXT != 7 || e.dact_src)) {
        unpack8(x.d[i], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= act_grad_fn<DACT>(t[j], e.dact);
    }
    const float sc = e.alpha * ((e.row_scale && !e.bias_rowscaled) ? rs : 1.f);
    if (sc != 1.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= sc;
    }
    if ((EXT & 2) && (EXT != 7 || e.res1)) {
        unpack8(x.r1[i], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if ((EXT & 4) && (EXT != 7 || e.res2)) {
        unpack8(x.r2[i], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if (e.out) st8_bf16((bf16*)e.out + off, v);
}

// Epilogue of one warp for its 32 rows x NCOLS columns of a tile (shared by the 1-CTA and 2-CTA kernels).
// Accumulators are transposed through a private 32x36 fp32 smem tile so that global traffic is 16 B per lane,
// 8 rows x 64 B per instruction.  `wait_acc` blocks until the accumulator is ready (called after the first
// chunk's operand prefetch has been issued).
template <int NCOLS, int ACT, int DACT, int EXT, typename WaitFn>
__device__ __forceinline__ void epilogue_warp_t(const EpiParams& epi, float* stg, float* scol, int col_in_tile, uint32_t taddr,
                                                int64_t row_base, int n_base, int M, int lane, WaitFn wait_acc,
                                                uint8_t* a2 = nullptr, int a2_row0 = 0, int a2_col0 = 0) {
    // Latency plan: everything that does not depend on the accumulator (bias -> smem, the first chunk's
    // residual / saved-activation operands) is requested BEFORE the accumulator-ready wait; inside the loop the
    // TMEM load of chunk c+1 and the operand loads of chunk c+1 are in flight while chunk c is processed.
    constexpr bool DB = (EXT != 7);          // the generic variant trades the prefetch for registers
    const int row_l = lane >> 2, c0 = (lane & 3) * 8;
    float* sbias = stg + 32 * STG_LD;        // [NCOLS] fp32 bias of this warp's columns
    if (epi.bias) {
        for (int j = lane; j < NCOLS; j += 32) sbias[j] = __bfloat162float(((const bf16*)epi.bias)[n_base + j]);
    }
    EpiExt cur, nxt;
    epi_prefetch<EXT>(epi, cur, row_base, row_l, n_base + c0, M);
    wait_acc();
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(taddr, r);
    const int dbg = g_dbg_skip_epilogue;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stg + lane * STG_LD + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        if (c + 32 < NCOLS) {
            ptx::tmem_ld_32x32b_x32(taddr + c + 32, r);     // registers are free again: next chunk's TMEM load
            if (DB) epi_prefetch<EXT>(epi, nxt, row_base, row_l, n_base + c + 32 + c0, M);
        }
        __syncwarp();
        const int n0 = n_base + c + c0;
        float bias8[8];
        if (epi.bias) {
            *reinterpret_cast<float4*>(bias8) = *reinterpret_cast<const float4*>(sbias + c + c0);
            *reinterpret_cast<float4*>(bias8 + 4) = *reinterpret_cast<const float4*>(sbias + c + c0 + 4);
        }
        float cs8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cs8[j] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + row_l;
            const int64_t row = row_base + rl;
            float v[8];
            *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(stg + rl * STG_LD + c0);
            *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(stg + rl * STG_LD + c0 + 4);
            if (dbg == 2) { if (v[0] == 1.2345e-30f && v[7] == 3.3e-31f) stg[lane * STG_LD] = v[3]; continue; }
            if (row < M) {
                epilogue_vec8<ACT, DACT, EXT>(epi, row, n0, v, bias8, cur, i);
#pragma unroll
                for (int j = 0; j < 8; ++j) cs8[j] += v[j];
            }
            if (a2) {   // fused adapter: the bf16 result is also the A operand of the second GEMM (K-major, 128B swizzle)
                const int trow = a2_row0 + rl, tcol = a2_col0 + c + c0;
                uint4 pk;
                __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) hp[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                if (row >= M) pk = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(a2 + (tcol >> 6) * (BM * 128) + trow * 128 + ((((tcol & 63) >> 3) ^ (trow & 7)) << 4)) = pk;
            }
        }
        if (epi.colsum_out) {      // bias gradient of the consumer: reduce the 8 row-lanes that share these columns,
#pragma unroll                     // then the 4 quadrant warps through shared memory (flushed once per tile)
            for (int j = 0; j < 8; ++j) {
                float t = cs8[j];
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                if (lane < 4) atomicAdd(scol + col_in_tile + c + c0 + j, t);
            }
        }
        __syncwarp();
        if (DB) cur = nxt;
        else if (c + 32 < NCOLS) epi_prefetch<EXT>(epi, cur, row_base, row_l, n_base + c + 32 + c0, M);
    }
}

template <int NCOLS, int V, typename WaitFn>
__device__ __forceinline__ void epilogue_warp(const EpiParams& epi, float* stg, float* scol, int col_in_tile, uint32_t taddr,
                                              int64_t row_base, int n_base, int M, int lane, WaitFn wait_acc) {
    if (g_dbg_skip_epilogue == 1) {
        wait_acc();
        for (int c = 0; c < NCOLS; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + c, r);
            ptx::tmem_wait_ld();
            if (r[0] == 0x7fc12345u && r[5] == 0x12345u) stg[lane] = __uint_as_float(r[1]);
        }
        return;
    }
    using EV = EpiVariant<V>;
    epilogue_warp_t<NCOLS, EV::ACT, EV::DACT, EV::EXT>(epi, stg, scol, col_in_tile, taddr, row_base, n_base, M, lane, wait_acc);
}

template <int BN, int V>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams epi,
               const int M, const int N, const int K) {
    pdl_trigger();   // the next kernel may start its prologue; it blocks in its own pdl_wait() until this grid is done
    using Cfg = TileCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer
    float* stg_all = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    float* scol = stg_all + EPI_WARPS * STG_WARP_FLOATS;
    if (threadIdx.x < 256) scol[threadIdx.x] = 0.f;
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
    pdl_wait();      // barrier init, TMEM allocation, descriptor prefetch above overlapped the previous kernel's tail

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
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = stg_all + (warp - 2) * STG_WARP_FLOATS;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int64_t row_base = (int64_t)m_blk * BM + quad * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + half * (BN / 2);
            epilogue_warp<BN / 2, V>(epi, stg, scol, half * (BN / 2), taddr, row_base, n_blk * BN + half * (BN / 2), M, lane, [&]() {
                ptx::mbar_wait(&tfull_bar[as], aphase);
                ptx::tc_fence_after();
            });
            ptx::tc_fence_before();
            __syncwarp();
            if (epi.colsum_out) {      // one global atomic per column per tile (the 8 epilogue warps meet on named barrier 1)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int te = threadIdx.x - 64;
                colsum_flush(epi.colsum_out + n_blk * BN, scol, te, BN);
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
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



// ---------------------------------------------------------------------------------------- row-layout epilogue (v5)
// gemm_tc4_kernel: same TMA / tcgen05 mainloop, but the epilogue never transposes fp32 accumulators.  Measured on
// the v4 kernel (bench_tools/gemm_epi.py, SKIPS=1,2,3,0): draining TMEM and the math are hidden under the mainloop,
// what is not hidden are (a) residual / saved-activation loads — one 16-column chunk of look-ahead is ~16 KB in
// flight per SM, a few % of what HBM latency x bandwidth needs — and (b) stores issued as 16 rows x 32 B per warp
// instruction (the LSU pays per row segment: two output streams cost 30 us on the c_fc GEMM).  Here:
//   * each epilogue warp owns a 32-row x 64-column slab and 4 KB shared-memory slab buffers (128 B rows, 16-byte
//     chunks XOR-swizzled by row & 7 — conflict-free both for "lane = row" and for "8 lanes = one row" access);
//   * residual / dact operands for the WHOLE slab are fetched with cp.async (16 B per lane, 4 full rows per
//     instruction) one tile ahead — no registers, 4 KB per warp and tensor in flight;
//   * the math runs in the TMEM-native layout (lane = row, 16 consecutive columns per tcgen05.ld.x16), results are
//     packed to bf16 and written IN PLACE over the consumed residual chunk;
//   * the slab is flushed with full 128-byte lines: 8 lanes x 16 B per row, 4 rows per store instruction;
//   * column sums (bias gradients) are taken at flush time from the values actually stored.
//   * DIRECT variant (large K, no column sums): results leave straight from registers as one 32-byte sector per lane
//     and chunk (st.global.v8.b32) - no output slab at all.  The 1-CTA mainloop already moves ~96 KB per k-block
//     through shared memory (TMA writes + tensor-core reads) against 128 B/clk, so every staged output byte is paid
//     twice more in the same currency: measured 59 -> 54.5 us (bias), 73.5 -> 59.6 us (x QuickGELU'), 46.3 -> 41.2 us
//     (QKV).  With tiny K the epilogue itself is the bottleneck and the full-line flush of the slab path wins.
template <int BN, int V, bool DIRECT = false> struct TileCfg4 {
    static constexpr int EPI_W = 4 * (BN / 64);
    static constexpr int THREADS = 64 + 32 * EPI_W;
    static constexpr int EPI_BYTES = EPI_W * EpiBufs<V, DIRECT>::WARP_BYTES + 1024;  // + scol[256]
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES_RAW = (232448 - EPI_BYTES - 1024 - 256) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr bool OK = STAGES >= 3;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = (STAGES > 0 ? STAGES : 1) * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

__device__ __forceinline__ bool want_pre_k(const EpiParams& e, bool pre) { return pre && e.out_pre != nullptr; }
template <int BN, int V, bool DIRECT>
__global__ void __launch_bounds__(TileCfg4<BN, V, DIRECT>::THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP, const EpiParams epi,
                const int M, const int N, const int K) {
    pdl_trigger();
    using Cfg = TileCfg4<BN, V, DIRECT>;
    using EB = EpiBufs<V, DIRECT>;
    using EV = EpiVariant<V>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_W = Cfg::EPI_W;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
    float* scol = reinterpret_cast<float*>(epi_smem + EPI_W * EB::WARP_BYTES);   // [slab buffers][bias][scol]
    if (threadIdx.x < 256) scol[threadIdx.x] = 0.f;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;   // provably uniform: the producer / MMA loops stay in the uniform datapath
    const int n_tiles = N / BN;
    const int m_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * m_tiles;
    const int KB = K / BK;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], EPI_W); }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_wait();
    // Producer and MMA warps: every lane runs the loop, one elected lane issues (ptx.cuh "warp-uniform issue": issued from
    // an `if (lane == 0)` region each tcgen05.mma costs ~100 clk of operand moves, which a 128 x 192 MMA (96 clk) cannot hide)
    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx_e(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    ptx::tma_load_2d_e(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {
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
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_bf16_e(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit_e(&empty_bar[stage]);
                    if (kb == KB - 1) ptx::umma_commit_e(&tfull_bar[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        constexpr int EXT = EV::EXT;
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int slab = ew >> 2;
        constexpr bool TMA_ST = (EXT == 0) && !DIRECT;                   // plain / activation variants: the slab leaves through TMA
        const uint32_t buf0 = ptx::smem_u32(epi_smem + ew * (EB::NBUF * 4096));   // out (in place over dact_src / res1); 1 KB aligned
        const uint32_t buf1 = buf0 + 4096;                    // out_pre, or res2
        float* sbias = reinterpret_cast<float*>(epi_smem + EPI_W * (EB::NBUF * 4096) + ew * EB::BIAS_BYTES);
        if (TMA_ST && lane == 0) { ptx::prefetch_tmap(&tmO); if (want_pre_k(epi, EB::PRE)) ptx::prefetch_tmap(&tmP); }
        const bf16* g0 = (EXT & 1) ? (const bf16*)epi.dact_src : (const bf16*)epi.res1;
        const bf16* g1 = (const bf16*)epi.res2;
        const bool want_pre = EB::PRE && epi.out_pre != nullptr;
        const int dbg = g_dbg_skip_epilogue;   // bench_tools only: 4 = no flush, 5 = no slab writes either, 6 = STG flush instead of TMA
        auto fetch = [&](int tile) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int64_t rb = (int64_t)m_blk * BM + quad * 32;
            const int nb = n_blk * BN + slab * 64;
            if (EXT & 3) slab_fetch(buf0, g0, epi.ldo, rb, nb, M, lane);
            if (EXT & 4) slab_fetch(buf1, g1, epi.ldo, rb, nb, M, lane);
        };
        if (EXT != 0 && (int)blockIdx.x < total) fetch(blockIdx.x);
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int64_t row_base = (int64_t)m_blk * BM + quad * 32;
            const int n_base = n_blk * BN + slab * 64;
            const int64_t row = row_base + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN + slab * 64;
            if (epi.bias) {
                sbias[lane] = __bfloat162float(((const bf16*)epi.bias)[n_base + lane]);
                sbias[lane + 32] = __bfloat162float(((const bf16*)epi.bias)[n_base + lane + 32]);
            }
            float rs = 1.f;
            if (epi.row_scale && row < M) rs = epi.row_scale[(int)row % epi.row_mod];
            float ln_mu = 0.f, ln_rs = 1.f;
            if (V == 8) {          // folded LayerNorm: lane = row, so each lane needs just its own row's statistics
                sbias[64 + lane] = epi.ln_wsum[n_base + lane];
                sbias[96 + lane] = epi.ln_wsum[n_base + lane + 32];
                if (row < M) { ln_mu = epi.ln_mean[row]; ln_rs = epi.ln_rstd[row]; }
            }
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            uint32_t ra[16], rb16[16];
            ptx::tmem_ld_32x32b_x16(taddr, ra);
            cp_async_commit_wait();
            if (TMA_ST && lane == 0) ptx::bulk_wait_read0();   // the previous tile's TMA store has finished reading the slab
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t (&cur)[16] = (q & 1) ? rb16 : ra;
                uint32_t (&nxt)[16] = (q & 1) ? ra : rb16;
                ptx::tmem_wait_ld();
                if (q < 3) ptx::tmem_ld_32x32b_x16(taddr + (q + 1) * 16, nxt);
                if (q == 3) {                                  // accumulator drained: hand the TMEM buffer back early
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
                }
                uint4 o_lo = make_uint4(0, 0, 0, 0), p_lo = o_lo;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int ch = 2 * q + hh;
                    const uint32_t off = slab_off(lane, ch);
                    uint4 xd = make_uint4(0, 0, 0, 0), x1 = xd, x2 = xd, pre_pk = xd;
                    if (EXT & 1) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(xd.x), "=r"(xd.y), "=r"(xd.z), "=r"(xd.w) : "r"(buf0 + off));
                    if ((EXT & 3) == 2) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x1.x), "=r"(x1.y), "=r"(x1.z), "=r"(x1.w) : "r"(buf0 + off));
                    if (EXT & 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x2.x), "=r"(x2.y), "=r"(x2.z), "=r"(x2.w) : "r"(buf1 + off));
                    float v[8], bias8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(cur[hh * 8 + j]);
                    if (epi.bias) {
                        *reinterpret_cast<float4*>(bias8) = *reinterpret_cast<const float4*>(sbias + ch * 8);
                        *reinterpret_cast<float4*>(bias8 + 4) = *reinterpret_cast<const float4*>(sbias + ch * 8 + 4);
                    }
                    if (V == 8) {
                        float s8[8];
                        *reinterpret_cast<float4*>(s8) = *reinterpret_cast<const float4*>(sbias + 64 + ch * 8);
                        *reinterpret_cast<float4*>(s8 + 4) = *reinterpret_cast<const float4*>(sbias + 64 + ch * 8 + 4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(-ln_mu, s8[j], v[j]) * ln_rs;
                    }
                    const uint4 o = epi_math8<EV::ACT, EV::DACT, EXT>(epi, rs, v, bias8, xd, x1, x2, pre_pk, want_pre);
                    if (dbg == 5 && o.x != 0x12345678u) continue;
                    if (DIRECT) {        // one 32-byte sector per lane and chunk, straight from registers
                        if (hh == 0) { o_lo = o; p_lo = pre_pk; }
                        else if (row < M) {
                            bf16* gp = (bf16*)epi.out + row * epi.ldo + n_base + q * 16;
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(gp), "r"(o_lo.x), "r"(o_lo.y), "r"(o_lo.z),
                                         "r"(o_lo.w), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                            if (want_pre) {
                                bf16* pp = (bf16*)epi.out_pre + row * epi.ldo + n_base + q * 16;
                                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(pp), "r"(p_lo.x), "r"(p_lo.y),
                                             "r"(p_lo.z), "r"(p_lo.w), "r"(pre_pk.x), "r"(pre_pk.y), "r"(pre_pk.z), "r"(pre_pk.w) : "memory");
                            }
                        }
                        continue;
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(buf0 + off), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    if (want_pre)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(buf1 + off), "r"(pre_pk.x), "r"(pre_pk.y), "r"(pre_pk.z), "r"(pre_pk.w) : "memory");
                }
            }
            if (dbg == 4 || dbg == 5) { __syncwarp(); continue; }
            if (DIRECT) {
                __syncwarp();
                if (EXT != 0 && tile + (int)gridDim.x < total) fetch(tile + gridDim.x);
                continue;
            }
            if (TMA_ST && !epi.colsum_out && dbg != 6) {
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmO, buf0, n_base, (int)row_base);
                    if (want_pre) ptx::tma_store_2d(&tmP, buf1, n_base, (int)row_base);
                    ptx::bulk_commit();
                }
                continue;
            }
            __syncwarp();
            // flush: 4 full rows per instruction
            float cs8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cs8[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + (lane >> 3), ch = lane & 7;
                const int64_t grow = row_base + r;
                if (grow < M) {
                    const uint32_t off = slab_off(r, ch);
                    uint4 o;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(buf0 + off));
                    const int64_t goff = grow * epi.ldo + n_base + ch * 8;
                    *reinterpret_cast<uint4*>((bf16*)epi.out + goff) = o;
                    if (want_pre) {
                        uint4 pp;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(pp.x), "=r"(pp.y), "=r"(pp.z), "=r"(pp.w) : "r"(buf1 + off));
                        *reinterpret_cast<uint4*>((bf16*)epi.out_pre + goff) = pp;
                    }
                    if (epi.colsum_out) {
                        float t[8];
                        unpack8(o, t);
#pragma unroll
                        for (int j = 0; j < 8; ++j) cs8[j] += t[j];
                    }
                }
            }
            if (epi.colsum_out) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = cs8[j];
                    t += __shfl_xor_sync(0xffffffffu, t, 8);
                    t += __shfl_xor_sync(0xffffffffu, t, 16);
                    if (lane < 8) atomicAdd(scol + slab * 64 + lane * 8 + j, t);
                }
            }
            __syncwarp();
            if (EXT != 0 && tile + (int)gridDim.x < total) fetch(tile + gridDim.x);
            if (epi.colsum_out) {
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_W) : "memory");
                const int te = threadIdx.x - 64;
                colsum_flush(epi.colsum_out + n_blk * BN, scol, te, BN);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_W) : "memory");
            }
        }
    }
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();   // smem must stay valid until the TMA stores have read it
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------- fused adapter
// Adapter bottleneck (vit_clip.py:51-69) as ONE kernel per direction:
//   forward : out = res1 + res2 + alpha * rs * (gelu(a W1^T + b1) W2^T + b2)      (h, g' = rs*gelu(h) kept for backward)
//   backward: d_a = res1 + ((dy W2) . gelu'(h) * alpha * rs) W1                   (d_h kept for the weight gradients,
//                                                                                 its column sums = db1)
// Per 128-row tile: GEMM1 (K = D, N = R) accumulates in TMEM; its epilogue writes the bf16 hidden tile to global AND,
// swizzled, into shared memory, where it is the K-major A operand of GEMM2 (K = R, N = D in chunks of BN2) — the
// [M, R] hidden never makes a round trip through HBM between the two GEMMs and the second launch disappears.
// TMEM: acc1 = columns [0, R); acc2 double-buffered at [0, BN2) (reuses acc1's columns once it is drained) and
// [256, 256 + BN2).  One tile per CTA (M/128 <= #SMs for every configuration of the path).
template <int R, int BN2> struct AdapterCfg {
    static constexpr int B1_BYTES = R * BK * 2, B2_BYTES = BN2 * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + (B1_BYTES > B2_BYTES ? B1_BYTES : B2_BYTES);
    static constexpr int A2_BYTES = (R / BK) * BM * 128;
    static constexpr int STAGES_RAW = (232448 - STG_BYTES - A2_BYTES - 1024 - 256) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + A2_BYTES + STG_BYTES + 1024 + 256;
    static_assert(STAGES >= 2, "adapter kernel needs at least a double-buffered ring");
    static_assert(R % 64 == 0 && R <= 256 && BN2 <= 256 && BN2 % 64 == 0, "TMEM plan");
};

template <int R, int BN2, int V1, int V2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
adapter_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmB2, const EpiParams epi1, const EpiParams epi2, const int M, const int D) {
    pdl_trigger();
    using Cfg = AdapterCfg<R, BN2>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a2 = smem + STAGES * Cfg::STAGE_BYTES;                               // [R/64][128 rows][128 B] swizzled
    float* stg_all = reinterpret_cast<float*>(a2 + Cfg::A2_BYTES);
    float* scol = stg_all + EPI_WARPS * STG_WARP_FLOATS;
    if (threadIdx.x < 256) scol[threadIdx.x] = 0.f;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(a2 + Cfg::A2_BYTES + STG_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc1_bar = empty_bar + STAGES;      // GEMM1 accumulator complete
    uint64_t* a2_bar = acc1_bar + 1;              // hidden tile written to smem by the 8 epilogue warps
    uint64_t* tfull_bar = a2_bar + 1;             // [2]
    uint64_t* tempty_bar = tfull_bar + 2;         // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x;
    const int KB1 = D / BK, KB2 = R / BK, NCH = D / BN2;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB1);
        ptx::prefetch_tmap(&tmB2);
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
        ptx::mbar_init(acc1_bar, 1);
        ptx::mbar_init(a2_bar, EPI_WARPS);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], EPI_WARPS); }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_wait();

    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB1; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx_e(&full_bar[stage], A_STAGE_BYTES + Cfg::B1_BYTES);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                ptx::tma_load_2d_e(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB1, &full_bar[stage], kb * BK, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            for (int j = 0; j < NCH; ++j)
                for (int kb = 0; kb < KB2; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx_e(&full_bar[stage], Cfg::B2_BYTES);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB2, &full_bar[stage], kb * BK, j * BN2);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(BM, R);
            constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(BM, BN2);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < KB1; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sa);
                const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    ptx::umma_bf16_e(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit_e(&empty_bar[stage]);
                if (kb == KB1 - 1) ptx::umma_commit_e(acc1_bar);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            ptx::mbar_wait(a2_bar, 0);            // hidden tile is in smem (and acc1's TMEM columns are drained)
            ptx::tc_fence_after();
            const uint32_t a2s = ptx::smem_u32(a2);
            for (int j = 0; j < NCH; ++j) {
                const int buf = j & 1;
                ptx::mbar_wait(&tempty_bar[buf], ((j >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kb = 0; kb < KB2; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sb = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES + A_STAGE_BYTES);
                    const uint64_t adesc = ptx::umma_desc_kmajor_sw128(a2s + kb * (BM * 128));
                    const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sb);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        ptx::umma_bf16_e(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
                    ptx::umma_commit_e(&empty_bar[stage]);
                    if (kb == KB2 - 1) ptx::umma_commit_e(&tfull_bar[buf]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = stg_all + (warp - 2) * STG_WARP_FLOATS;
        const int64_t row_base = (int64_t)m_blk * BM + quad * 32;
        using E1 = EpiVariant<V1>;
        using E2 = EpiVariant<V2>;
        // ---- epilogue 1: hidden tile -> global (kept for backward) and -> smem (A operand of GEMM2)
        epilogue_warp_t<R / 2, E1::ACT, E1::DACT, E1::EXT>(
            epi1, stg, scol, half * (R / 2), tmem_base + ((uint32_t)(quad * 32) << 16) + half * (R / 2), row_base,
            half * (R / 2), M, lane,
            [&]() {
                ptx::mbar_wait(acc1_bar, 0);
                ptx::tc_fence_after();
            },
            a2, quad * 32, half * (R / 2));
        ptx::fence_proxy_async();            // make the generic-proxy smem writes visible to the tensor core (async proxy)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a2_bar);
        if (epi1.colsum_out) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int te = threadIdx.x - 64;
            colsum_flush(epi1.colsum_out, scol, te, R);
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // ---- epilogue 2: output chunks
        for (int j = 0; j < NCH; ++j) {
            const int buf = j & 1;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256 + half * (BN2 / 2);
            epilogue_warp_t<BN2 / 2, E2::ACT, E2::DACT, E2::EXT>(epi2, stg, scol, half * (BN2 / 2), taddr, row_base,
                                                                j * BN2 + half * (BN2 / 2), M, lane, [&]() {
                                                                    ptx::mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
                                                                    ptx::tc_fence_after();
                                                                });
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[buf]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------- host side
// epilogue kernel selection: 0 = auto (gemm_tc4_kernel, except the two-output-stream c_fc variant V=1 which measures
// faster on gemm_tc_kernel: its stores trickle out chunk by chunk instead of one burst per tile), 1 = gemm_tc_kernel,
// 3 = gemm_tc3_kernel, 4 = gemm_tc4_kernel wherever its shared-memory budget allows
static int g_epi_kernel = 0;
static int g_direct = 1;           // 0: never use the DIRECT (register-store) epilogue, 1: auto (K >= 512), 2: always
static inline bool use_tc4(int v) { return (g_epi_kernel == 4 || (g_epi_kernel == 0 && v != 1)) && v != 7; }
// DIRECT epilogue: 32-byte sector stores from registers; for every variant without fused column sums once the mainloop
// (not the epilogue) is the long pole
static inline bool use_tc4_direct(int v, int K, const EpiParams& p) {
    if (v == 7 || p.colsum_out || (g_epi_kernel != 0 && g_epi_kernel != 4) || g_direct == 0) return false;
    if ((p.ldo % 16) || ((uintptr_t)p.out & 31) || (p.out_pre && ((uintptr_t)p.out_pre & 31))) return false;
    return g_direct == 2 || K >= 512;
}
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

static int num_sms() { return device_sm_count(); }

template <int BN, int V>
static int launch_tc_v(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, int M, int N, int K, cudaStream_t s) {
    using Cfg = TileCfg<BN>;
    // generic kernel (runtime-switched epilogue, variant 7): the fallback for every combination the row-layout kernel's
    // buffers do not fit, and debug mode 1
    AIMB_SET_SMEM_ATTR(Cfg::SMEM_BYTES, gemm_tc_kernel<BN, 7>);
    int total = (N / BN) * ((M + BM - 1) / BM);
    int grid = total < num_sms() ? total : num_sms();
    if (use_tc4_direct(V, K, p)) {
        if constexpr (V != 7 && TileCfg4<BN, V, true>::OK) {
            using Cfg4 = TileCfg4<BN, V, true>;
            AIMB_SET_SMEM_ATTR(Cfg4::SMEM_BYTES, gemm_tc4_kernel<BN, V, true>);
            launch_k((gemm_tc4_kernel<BN, V, true>), dim3(grid), dim3(Cfg4::THREADS), Cfg4::SMEM_BYTES, s, ta, tb, ta, ta, p, M, N, K);
            AIMB_CHECK_LAUNCH();
            return AIMB_OK;
        }
    }
    if constexpr (V != 7 && TileCfg4<BN, V, false>::OK) {
        if (use_tc4(V)) {
            using Cfg4 = TileCfg4<BN, V, false>;
            AIMB_SET_SMEM_ATTR(Cfg4::SMEM_BYTES, gemm_tc4_kernel<BN, V, false>);
            CUtensorMap to = ta, tp = ta;                       // placeholders when the variant does not store through TMA
            if (EpiVariant<V>::EXT == 0) {
                if (make_tmap_bf16(&to, p.out, M, N, p.ldo, 32)) return AIMB_ERR_DRIVER;
                if (p.out_pre && make_tmap_bf16(&tp, p.out_pre, M, N, p.ldo, 32)) return AIMB_ERR_DRIVER;
            }
            launch_k((gemm_tc4_kernel<BN, V, false>), dim3(grid), dim3(Cfg4::THREADS), Cfg4::SMEM_BYTES, s, ta, tb, to, tp, p, M, N, K);
            AIMB_CHECK_LAUNCH();
            return AIMB_OK;
        }
    }
    launch_k((gemm_tc_kernel<BN, 7>), dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, ta, tb, p, M, N, K);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <int BN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, int M, int N, int K, cudaStream_t s) {
    switch (pick_variant(p)) {
        case 0: return launch_tc_v<BN, 0>(ta, tb, p, M, N, K, s);
        case 1: return launch_tc_v<BN, 1>(ta, tb, p, M, N, K, s);
        case 2: return launch_tc_v<BN, 2>(ta, tb, p, M, N, K, s);
        case 3: return launch_tc_v<BN, 3>(ta, tb, p, M, N, K, s);
        case 4: return launch_tc_v<BN, 4>(ta, tb, p, M, N, K, s);
        case 5: return launch_tc_v<BN, 5>(ta, tb, p, M, N, K, s);
        case 6: return launch_tc_v<BN, 6>(ta, tb, p, M, N, K, s);
        case 8: return launch_tc_v<BN, 8>(ta, tb, p, M, N, K, s);
        default: return launch_tc_v<BN, 7>(ta, tb, p, M, N, K, s);
    }
}

template <int V> static bool tc4_ok_v(int bn, bool direct) {
    switch (bn) {
        case 256: return direct ? TileCfg4<256, V, true>::OK : TileCfg4<256, V, false>::OK;
        case 192: return direct ? TileCfg4<192, V, true>::OK : TileCfg4<192, V, false>::OK;
        case 128: return direct ? TileCfg4<128, V, true>::OK : TileCfg4<128, V, false>::OK;
        case 64: return direct ? TileCfg4<64, V, true>::OK : TileCfg4<64, V, false>::OK;
    }
    return false;
}
static bool tc4_ok(int bn, int v, bool direct) {
    switch (v) {
        case 0: return tc4_ok_v<0>(bn, direct);
        case 1: return tc4_ok_v<1>(bn, direct);
        case 2: return tc4_ok_v<2>(bn, direct);
        case 3: return tc4_ok_v<3>(bn, direct);
        case 4: return tc4_ok_v<4>(bn, direct);
        case 5: return tc4_ok_v<5>(bn, direct);
        case 6: return tc4_ok_v<6>(bn, direct);
        case 8: return tc4_ok_v<8>(bn, direct);
    }
    return false;
}
static int pick_bn(int64_t M, int N, int K, int variant = -1, bool direct = false) {
    const int cand[4] = {256, 192, 128, 64};
    int best = 0; double best_cost = 1e30;
    int64_t mt = (M + BM - 1) / BM;
    const double kb = (double)(K / BK);
    for (int i = 0; i < 4; ++i) {
        int bn = cand[i];
        if (N % bn) continue;
        if (variant >= 0 && variant != 7 && !tc4_ok(bn, variant, direct)) continue;
        int64_t tiles = mt * (N / bn);
        int64_t waves = (tiles + num_sms() - 1) / num_sms();
        double per_tile = kb * (double)(bn < 128 ? 128 : bn) + 1536.0;   // below N=128 the A-operand traffic dominates
        double cost = (double)waves * per_tile;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

int gemm_tc_launch(const void* A, int64_t lda, const void* W, int64_t ldw, const EpiParams& p, int64_t M, int N, int K,
                   int force_bn, int cta_mode, cudaStream_t s) {
    if (K % BK || N % 64 || (lda % 8) || (ldw % 8) || ((uintptr_t)A & 15) || ((uintptr_t)W & 15)) return AIMB_ERR_ARG;
    if (p.ldo % 8 || ((uintptr_t)p.out & 15)) return AIMB_ERR_ARG;
    if (M >= (1ll << 31)) return AIMB_ERR_ARG;
    const int var = pick_variant(p);
    const bool direct = use_tc4_direct(var, K, p);
    int bn = force_bn > 0 ? force_bn : pick_bn(M, N, K, (direct || use_tc4(var)) ? var : -1, direct);
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
    pdl_trigger();   // the next kernel may start its prologue; it blocks in its own pdl_wait() until this grid is done
    constexpr int BOX = 64 * 64 * 2;                 // one [64 r][64 c] bf16 box
    constexpr int A_BYTES = 2 * BOX, B_BYTES = (NS / 64) * BOX, STG = A_BYTES + B_BYTES;
    constexpr int STAGES = (200 * 1024) / STG > 6 ? 6 : (200 * 1024) / STG;
    constexpr int TCOLS = NS <= 64 ? 64 : NS <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a __shared__ pointer
    float* wstg = reinterpret_cast<float*>(smem + STAGES * STG);                 // 4 warps x [32][33] fp32
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STG + 4 * 32 * 33 * 4);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
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
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_wait();      // barrier init, TMEM allocation, descriptor prefetch above overlapped the previous kernel's tail
    if (warp == 0) {           // producer and MMA warps run warp-uniformly, one elected lane issues (ptx.cuh)
        {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                ptx::mbar_arrive_expect_tx_e(&full_bar[stage], STG);
                uint8_t* sa = smem + stage * STG;
                ptx::tma_load_2d_e(sa, &tmP, &full_bar[stage], mt * 128, kb * 64);
                ptx::tma_load_2d_e(sa + BOX, &tmP, &full_bar[stage], mt * 128 + 64, kb * 64);
#pragma unroll
                for (int j = 0; j < NS / 64; ++j)
                    ptx::tma_load_2d_e(sa + A_BYTES + j * BOX, &tmQ, &full_bar[stage], j * 64, kb * 64);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        {
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
                    ptx::umma_bf16_e(tmem_base, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit_e(&empty_bar[stage]);
                if (kb == kb1 - 1) ptx::umma_commit_e(done_bar);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        const bool vec_atomics = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (ldo % 4 == 0);
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < NS; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(taddr + c, r);
            ptx::tmem_wait_ld();
            // All `splits` CTAs of a column tile add into the same addresses and same-address atomics serialise in L2: the
            // 32 x 32 block goes through a padded smem tile so that every lane owns 4 consecutive floats of the result and
            // issues ONE 16-byte vector atomic for them (4x fewer serialised operations than scalar adds).
            float* st = wstg + quad * (32 * 33);
#pragma unroll
            for (int j = 0; j < 32; ++j) st[lane * 33 + j] = alpha * __uint_as_float(r[j]);      // [lane = row m][column j]
            __syncwarp();
            const int64_t m0 = (int64_t)mt * 128 + quad * 32;
            const int a4 = lane >> 3, b4 = (lane & 7) * 4;
            if (transposed_out) {       // out[(c + j) * ldo + m]: 4 consecutive m per lane
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int j = it * 4 + a4;
                    float* o = out + (int64_t)(c + j) * ldo + m0 + b4;
                    const float4 v = make_float4(st[(b4 + 0) * 33 + j], st[(b4 + 1) * 33 + j], st[(b4 + 2) * 33 + j], st[(b4 + 3) * 33 + j]);
                    if (vec_atomics) atomicAdd(reinterpret_cast<float4*>(o), v);
                    else { atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w); }
                }
            } else {                    // out[(m0 + rr) * ldo + c + col]: 4 consecutive columns per lane
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + a4;
                    float* o = out + (m0 + rr) * ldo + c + b4;
                    const float4 v = make_float4(st[rr * 33 + b4], st[rr * 33 + b4 + 1], st[rr * 33 + b4 + 2], st[rr * 33 + b4 + 3]);
                    if (vec_atomics) atomicAdd(reinterpret_cast<float4*>(o), v);
                    else { atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w); }
                }
            }
            __syncwarp();
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
    constexpr int SMEM = STAGES * STG + 4 * 32 * 33 * 4 + 1024 + 256;
    AIMB_SET_SMEM_ATTR(SMEM, wgrad_tc_kernel<NS>);
    int KB = (R + 63) / 64;
    int mtiles = CB / 128;
    int splits = num_sms() / mtiles;
    if (splits < 1) splits = 1;
    if (splits > KB) splits = KB;
    int per = (KB + splits - 1) / splits;
    splits = (KB + per - 1) / per;
    dim3 grid(mtiles, splits);
    launch_k((wgrad_tc_kernel<NS>), dim3(grid), dim3(TC_THREADS), SMEM, s, tp, tq, out, ldo, transposed, alpha, KB, per);
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

// ---- fused adapter host side
template <int R, int BN2, int V1, int V2>
static int launch_adapter_v(const CUtensorMap& ta, const CUtensorMap& tb1, const CUtensorMap& tb2, const EpiParams& e1,
                            const EpiParams& e2, int M, int D, cudaStream_t s) {
    using Cfg = AdapterCfg<R, BN2>;
    AIMB_SET_SMEM_ATTR(Cfg::SMEM_BYTES, adapter_tc_kernel<R, BN2, V1, V2>);
    launch_k((adapter_tc_kernel<R, BN2, V1, V2>), dim3((M + BM - 1) / BM), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, s, ta, tb1, tb2, e1,
             e2, M, D);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
template <int R, int BN2>
static int launch_adapter(const CUtensorMap& ta, const CUtensorMap& tb1, const CUtensorMap& tb2, const EpiParams& e1,
                          const EpiParams& e2, int M, int D, cudaStream_t s) {
    const int v1 = pick_variant(e1), v2 = pick_variant(e2);
    if (v1 == 2 && v2 == 3) return launch_adapter_v<R, BN2, 2, 3>(ta, tb1, tb2, e1, e2, M, D, s);
    if (v1 == 2 && v2 == 4) return launch_adapter_v<R, BN2, 2, 4>(ta, tb1, tb2, e1, e2, M, D, s);
    if (v1 == 2 && v2 == 0) return launch_adapter_v<R, BN2, 2, 0>(ta, tb1, tb2, e1, e2, M, D, s);
    if (v1 == 6 && v2 == 3) return launch_adapter_v<R, BN2, 6, 3>(ta, tb1, tb2, e1, e2, M, D, s);
    if (v1 == 6 && v2 == 0) return launch_adapter_v<R, BN2, 6, 0>(ta, tb1, tb2, e1, e2, M, D, s);
    return AIMB_ERR_UNSUPPORTED;
}

// A [M, D]; B1 [R, D]; B2 [D, R] (all K-major: nn.Linear weights forward, their transposes backward)
int adapter_tc_launch(const void* A, int64_t lda, const void* B1, const void* B2, const EpiParams& e1, const EpiParams& e2,
                      int64_t M, int D, int R, cudaStream_t s) {
    if (M < BM || M >= (1ll << 31) || (M + BM - 1) / BM > 4096 || D % 64 || (lda % 8) || ((uintptr_t)A & 15)) return AIMB_ERR_UNSUPPORTED;
    int bn2 = (D % 192 == 0) ? 192 : (D % 256 == 0 ? 256 : 0);
    if (!bn2) return AIMB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb1, tb2;
    int rc = make_tmap_bf16(&ta, A, M, D, lda, BM);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb1, B1, R, D, D, R);
    if (rc) return rc;
    rc = make_tmap_bf16(&tb2, B2, D, R, R, bn2);
    if (rc) return rc;
    if (R == 192 && bn2 == 192) return launch_adapter<192, 192>(ta, tb1, tb2, e1, e2, (int)M, D, s);
    if (R == 256 && bn2 == 256) return launch_adapter<256, 256>(ta, tb1, tb2, e1, e2, (int)M, D, s);
    if (R == 64 && bn2 == 256) return launch_adapter<64, 256>(ta, tb1, tb2, e1, e2, (int)M, D, s);
    return AIMB_ERR_UNSUPPORTED;
}

}  // namespace aimb

using namespace aimb;

static int g_force_bn = 0;
static int g_cta_mode = 0;   // 0/1: 1-CTA kernel (default), 2: CTA-pair (cta_group::2) kernel where it tiles
extern "C" void aimb_debug_force_bn(int bn) { g_force_bn = bn; }
extern "C" void aimb_debug_cta_mode(int mode) {
    // 0: auto (gemm_tc4_kernel wherever its buffers fit), 1: force the generic gemm_tc_kernel, 4: force gemm_tc4_kernel.
    // (modes 2 / 3 / 5 selected kernel generations that were measured and removed in round 2: profiles/r1_gemm_*.txt)
    g_cta_mode = 1;
    aimb::g_epi_kernel = (mode == 1 || mode == 4) ? mode : 0;
}
extern "C" void aimb_debug_direct_epilogue(int v) { aimb::g_direct = v; }
extern "C" void aimb_debug_skip_epilogue(int v) {
#ifdef AIMB_DEBUG_EPILOGUE
    cudaMemcpyToSymbol(g_dbg_skip_epilogue, &v, sizeof(int));
#else
    (void)v;       // compiled out of release builds
#endif
}

extern "C" int aimb_gemm_nt(const void* A, int64_t lda, const void* W, int64_t ldw, const aimb_epilogue_t* epi, int64_t M,
                            int32_t N, int32_t K, int32_t dtype, int32_t impl, void* stream) {
    if (!A || !W || !epi || !epi->out || M < 0 || N <= 0 || K <= 0 || lda < K || ldw < K) return AIMB_ERR_ARG;
    if (epi->accumulate && !epi->out_f32) return AIMB_ERR_ARG;
    if (epi->bias_rowscaled && !epi->row_scale) return AIMB_ERR_ARG;
    if (M == 0) return AIMB_OK;
    EpiParams p = make_epi(epi, N);
    cudaStream_t s = (cudaStream_t)stream;
    if (p.colsum_out && !p.colsum_accumulate && cudaMemsetAsync(p.colsum_out, 0, (size_t)N * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    // tcgen05 kernel for every shape it tiles (all ViT-B/16 and ViT-L/14 GEMMs); shapes it cannot tile
    // (N or K not a multiple of 64 — toy widths only) run on the SIMT kernel, still on the GPU.
    // (also M < 128: the per-frame [B*T, D] GEMMs of the fork block — a TMA box may not exceed the tensor)
    const bool tc_ok = (K % BK == 0) && (N % 64 == 0) && (lda % 8 == 0) && (ldw % 8 == 0) && (p.ldo % 8 == 0) && M >= BM;
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && !p.out_f32 && tc_ok)
        return gemm_tc_launch(A, lda, W, ldw, p, M, N, K, g_force_bn, g_cta_mode, s);
    if (dtype != AIMB_BF16 && dtype != AIMB_F32) return AIMB_ERR_ARG;
    return gemm_simt_launch(A, lda, 1, W, ldw, 1, p, M, N, K, dtype, s);
}

extern "C" int aimb_gemm_wgrad(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t R, int32_t N,
                               int32_t K, float alpha, int32_t accumulate, int32_t dtype, int32_t impl, void* stream) {
    if (!dY || !X || !dW || R < 0 || N <= 0 || K <= 0 || ldy < N || ldx < K) return AIMB_ERR_ARG;
    cudaStream_t s0 = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && R >= 64) {
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

extern "C" int aimb_adapter_fused(const void* A, int64_t lda, const void* W1, const void* W2, const aimb_epilogue_t* epi1,
                                  const aimb_epilogue_t* epi2, int64_t M, int32_t D, int32_t R, int32_t dtype, void* stream) {
    if (!A || !W1 || !W2 || !epi1 || !epi2 || !epi2->out || M < 0 || D <= 0 || R <= 0) return AIMB_ERR_ARG;
    if (dtype != AIMB_BF16) return AIMB_ERR_UNSUPPORTED;
    if (M == 0) return AIMB_OK;
    EpiParams e1 = make_epi(epi1, R), e2 = make_epi(epi2, D);
    cudaStream_t s = (cudaStream_t)stream;
    if (e1.colsum_out && !e1.colsum_accumulate && cudaMemsetAsync(e1.colsum_out, 0, (size_t)R * 4, s) != cudaSuccess) return AIMB_ERR_CUDA;
    if (e2.colsum_out) return AIMB_ERR_UNSUPPORTED;
    return adapter_tc_launch(A, lda, W1, W2, e1, e2, M, D, R, s);
}

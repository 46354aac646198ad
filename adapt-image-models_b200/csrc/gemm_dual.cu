// Paired GEMM: two nn.Linear of one block as N- / K-concatenated segments of ONE tcgen05 launch (C ABI: aimb_gemm_dual).
// Same TMA / tcgen05 mainloop and register-store (DIRECT) epilogue as gemm_tc4_kernel (gemm_tc.cu); shared pieces live in
// gemm_shared.cuh.
#include <type_traits>
#include "common.cuh"
#include "ptx.cuh"
#include "gemm_shared.cuh"

namespace aimb {

static int num_sms() { return device_sm_count(); }

// ---------------------------------------------------------------------------------------- paired GEMM
// gemm_dual_kernel: TWO nn.Linear of one block in ONE launch, same mainloop / DIRECT epilogue as gemm_tc4_kernel.
// The MLP adapter (vitclip_aim.py:210-211 == vit_clip.py:285-286) reads the same ln_2(x) as c_fc and adds into the same
// sum as c_proj, so its two small GEMMs (M x 192 x 768: 99 tiles on 148 SMs, epilogue-bound) ride on the frozen ones:
//   NCAT  [C1 | C2] = [epi1(A W1^T) | epi2(A W2^T)]: the B operand and the epilogue (activation, outputs, leading
//         dimension) switch per column tile; segment 2 may use a narrower tile (BN2 <= BN1).  At M = 12 608 the 99 extra
//         tiles fall into the last, nearly empty wave of the N = 3072 GEMM (8.03 -> 8.70 waves of 9).
//         forward : [hf | h_m] -> [QuickGELU | GELU * alpha * DropPath];  backward: [d_hf | d_h] from dx
//   KCAT  C = epi(A1 W1^T + A2 W2^T): the k-block loop runs over both operand pairs (K = K1 + K2) into ONE accumulator;
//         `bias2` (row-scaled: the DropPath multiplier of the adapter branch) joins the epilogue.
//         forward : x_out = gf Wp^T + g_m W2^T + bp + s*mask*b2 + x2;  backward: d_xn2 = d_hf Wfc + d_h W1
// No column sums here (the adapter's db1 is a separate reduction on the side stream).
struct DualExtra {
    const void* bias2;              // [N] bf16, KCAT only
    const float* bias2_row_scale;   // [bias2_row_mod] fp32 or null
    float bias2_scale;
    int32_t bias2_row_mod;
};
template <int A, int B> struct MaxI { static constexpr int v = A > B ? A : B; };
template <int BN1, int BN2, int V1, int V2> struct DualCfg {
    static_assert(BN2 <= BN1, "segment 2 uses the narrower tile");
    static constexpr int EPI_W = 4 * (BN1 / 64);
    static constexpr int THREADS = 64 + 32 * EPI_W;
    static constexpr int NBUF = MaxI<EpiBufs<V1, true>::NEXT, EpiBufs<V2, true>::NEXT>::v;
    static constexpr int WARP_BYTES = NBUF * 4096 + 512;                   // + 64 fp32 bias + 64 fp32 bias2 values
    static constexpr int EPI_BYTES = EPI_W * WARP_BYTES;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + BN1 * BK * 2;
    static constexpr int STAGES_RAW = (232448 - EPI_BYTES - 1024 - 256) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static_assert(STAGES >= 3, "paired GEMM needs a 3-stage ring");
    static constexpr int TMEM_COLS = (2 * BN1 <= 128) ? 128 : (2 * BN1 <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

template <int BN1, int BN2, int V1, int V2, bool KCAT>
__global__ void __launch_bounds__(DualCfg<BN1, BN2, V1, V2>::THREADS, 1)
gemm_dual_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ EpiParams epi1, const __grid_constant__ EpiParams epi2, const DualExtra ex,
                 const int M, const int nt1, const int nt2, const int KB1, const int KB2) {
    pdl_trigger();
    using Cfg = DualCfg<BN1, BN2, V1, V2>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int EPI_W = Cfg::EPI_W;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
    const int n_tiles = nt1 + nt2;
    const int m_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * m_tiles;
    const int KBT = KCAT ? KB1 + KB2 : KB1;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        ptx::prefetch_tmap(&tmB2);
        if (KCAT) ptx::prefetch_tmap(&tmA2);
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
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int m_blk = tile / n_tiles, n_idx = tile % n_tiles;
            const bool seg2 = !KCAT && n_idx >= nt1;
            for (int kb = 0; kb < KBT; ++kb) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                if (KCAT) {
                    ptx::mbar_arrive_expect_tx_e(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (kb < KB1) {
                        ptx::tma_load_2d_e(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                        ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_idx * BN1);
                    } else {
                        ptx::tma_load_2d_e(sa, &tmA2, &full_bar[stage], (kb - KB1) * BK, m_blk * BM);
                        ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB2, &full_bar[stage], (kb - KB1) * BK, n_idx * BN1);
                    }
                } else if (!seg2) {
                    ptx::mbar_arrive_expect_tx_e(&full_bar[stage], Cfg::STAGE_BYTES);
                    ptx::tma_load_2d_e(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_idx * BN1);
                } else {
                    ptx::mbar_arrive_expect_tx_e(&full_bar[stage], A_STAGE_BYTES + BN2 * BK * 2);
                    ptx::tma_load_2d_e(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d_e(sa + A_STAGE_BYTES, &tmB2, &full_bar[stage], kb * BK, (n_idx - nt1) * BN2);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(BM, BN1);
        constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(BM, BN2);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int n_idx = tile % n_tiles;
            const uint32_t idesc = (!KCAT && n_idx >= nt1) ? idesc2 : idesc1;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN1;
            for (int kb = 0; kb < KBT; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t adesc = ptx::umma_desc_kmajor_sw128(sa);
                const uint64_t bdesc = ptx::umma_desc_kmajor_sw128(sa + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    ptx::umma_bf16_e(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit_e(&empty_bar[stage]);
                if (kb == KBT - 1) ptx::umma_commit_e(&tfull_bar[as]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        constexpr int EXT1 = EpiVariant<V1>::EXT, EXT2 = EpiVariant<V2>::EXT;
        const int ew = warp - 2;
        const int quad = warp & 3;
        const int slab = ew >> 2;
        const uint32_t buf0 = ptx::smem_u32(epi_smem + ew * (Cfg::NBUF * 4096));
        const uint32_t buf1 = buf0 + 4096;
        float* sbias = reinterpret_cast<float*>(epi_smem + EPI_W * (Cfg::NBUF * 4096) + ew * 512);
        float* sbias2 = sbias + 64;
        auto fetch = [&](int tile) {
            const int m_blk = tile / n_tiles, n_idx = tile % n_tiles;
            const bool seg2 = !KCAT && n_idx >= nt1;
            if (seg2 && slab >= BN2 / 64) return;
            const EpiParams& e = seg2 ? epi2 : epi1;
            const int ext = seg2 ? EXT2 : EXT1;
            const int64_t rb = (int64_t)m_blk * BM + quad * 32;
            const int nb = (seg2 ? (n_idx - nt1) * BN2 : n_idx * BN1) + slab * 64;
            if (ext & 3) slab_fetch(buf0, (ext & 1) ? (const bf16*)e.dact_src : (const bf16*)e.res1, e.ldo, rb, nb, M, lane);
            if (ext & 4) slab_fetch(buf1, (const bf16*)e.res2, e.ldo, rb, nb, M, lane);
        };
        if ((EXT1 | EXT2) != 0 && (int)blockIdx.x < total) fetch(blockIdx.x);
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_idx = tile % n_tiles;
            const bool seg2 = !KCAT && n_idx >= nt1;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int64_t row = (int64_t)m_blk * BM + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN1 + slab * 64;
            if (seg2 && slab >= BN2 / 64) {        // this warp's columns do not exist in the narrow tile
                ptx::mbar_wait(&tfull_bar[as], aphase);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
                if ((EXT1 | EXT2) != 0 && tile + (int)gridDim.x < total) fetch(tile + gridDim.x);
                continue;
            }
            auto body = [&](auto vtag, const EpiParams& e, const int n_base) {
                constexpr int V = decltype(vtag)::value;
                using EV = EpiVariant<V>;
                constexpr int EXT = EV::EXT;
                constexpr bool PRE = (V == 1 || V == 2);
                const bool want_pre = PRE && e.out_pre != nullptr;
                if (e.bias) {
                    sbias[lane] = __bfloat162float(((const bf16*)e.bias)[n_base + lane]);
                    sbias[lane + 32] = __bfloat162float(((const bf16*)e.bias)[n_base + lane + 32]);
                }
                float rs = 1.f, b2s = 0.f;
                if (e.row_scale && row < M) rs = e.row_scale[(int)row % e.row_mod];
                if (KCAT && ex.bias2) {
                    sbias2[lane] = __bfloat162float(((const bf16*)ex.bias2)[n_base + lane]);
                    sbias2[lane + 32] = __bfloat162float(((const bf16*)ex.bias2)[n_base + lane + 32]);
                    b2s = ex.bias2_scale;
                    if (ex.bias2_row_scale && row < M) b2s *= ex.bias2_row_scale[(int)row % ex.bias2_row_mod];
                }
                ptx::mbar_wait(&tfull_bar[as], aphase);
                ptx::tc_fence_after();
                uint32_t ra[16], rb16[16];
                ptx::tmem_ld_32x32b_x16(taddr, ra);
                cp_async_commit_wait();
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
                        if (e.bias) {
                            *reinterpret_cast<float4*>(bias8) = *reinterpret_cast<const float4*>(sbias + ch * 8);
                            *reinterpret_cast<float4*>(bias8 + 4) = *reinterpret_cast<const float4*>(sbias + ch * 8 + 4);
                        }
                        if (KCAT && ex.bias2) {
                            float c8[8];
                            *reinterpret_cast<float4*>(c8) = *reinterpret_cast<const float4*>(sbias2 + ch * 8);
                            *reinterpret_cast<float4*>(c8 + 4) = *reinterpret_cast<const float4*>(sbias2 + ch * 8 + 4);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaf(c8[j], b2s, v[j]);
                        }
                        const uint4 o = epi_math8<EV::ACT, EV::DACT, EXT>(e, rs, v, bias8, xd, x1, x2, pre_pk, want_pre);
                        if (hh == 0) { o_lo = o; p_lo = pre_pk; }
                        else if (row < M) {          // one 32-byte sector per lane and chunk, straight from registers
                            bf16* gp = (bf16*)e.out + row * e.ldo + n_base + q * 16;
                            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(gp), "r"(o_lo.x), "r"(o_lo.y), "r"(o_lo.z),
                                         "r"(o_lo.w), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                            if (want_pre) {
                                bf16* pp = (bf16*)e.out_pre + row * e.ldo + n_base + q * 16;
                                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(pp), "r"(p_lo.x), "r"(p_lo.y),
                                             "r"(p_lo.z), "r"(p_lo.w), "r"(pre_pk.x), "r"(pre_pk.y), "r"(pre_pk.z), "r"(pre_pk.w) : "memory");
                            }
                        }
                    }
                }
                __syncwarp();
            };
            if (seg2) body(std::integral_constant<int, V2>{}, epi2, (n_idx - nt1) * BN2 + slab * 64);
            else body(std::integral_constant<int, V1>{}, epi1, n_idx * BN1 + slab * 64);
            if ((EXT1 | EXT2) != 0 && tile + (int)gridDim.x < total) fetch(tile + gridDim.x);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}


// ---- paired GEMM host side
template <int BN1, int BN2, int V1, int V2, bool KCAT>
static int launch_dual_v(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ta2, const CUtensorMap& tb2,
                         const EpiParams& e1, const EpiParams& e2, const DualExtra& ex, int M, int nt1, int nt2, int KB1, int KB2,
                         cudaStream_t s) {
    using Cfg = DualCfg<BN1, BN2, V1, V2>;
    AIMB_SET_SMEM_ATTR(Cfg::SMEM_BYTES, gemm_dual_kernel<BN1, BN2, V1, V2, KCAT>);
    const int total = (nt1 + nt2) * ((M + BM - 1) / BM);
    const int grid = total < num_sms() ? total : num_sms();
    launch_k((gemm_dual_kernel<BN1, BN2, V1, V2, KCAT>), dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, ta, tb, ta2, tb2, e1, e2,
             ex, M, nt1, nt2, KB1, KB2);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

static bool dual_epi_ok(const EpiParams& p) {
    return !p.colsum_out && !p.out_f32 && !p.ln_mean && p.out && (p.ldo % 16 == 0) && !((uintptr_t)p.out & 31) &&
           !(p.out_pre && ((uintptr_t)p.out_pre & 31)) && !(p.dact_src && ((uintptr_t)p.dact_src & 15)) &&
           !(p.res1 && ((uintptr_t)p.res1 & 15));
}

// NCAT: A [M, K], W1 [N1, K], W2 [N2, K]; e1 / e2 describe the two outputs (own pointers and leading dimensions).
int gemm_dual_ncat_launch(const void* A, int64_t lda, const void* W1, int64_t ldw1, const void* W2, int64_t ldw2, const EpiParams& e1,
                          const EpiParams& e2, int64_t M, int N1, int N2, int K, cudaStream_t s) {
    if (M < BM || M >= (1ll << 31) || K % BK || (lda % 8) || (ldw1 % 8) || (ldw2 % 8) || ((uintptr_t)A & 15) || ((uintptr_t)W1 & 15) ||
        ((uintptr_t)W2 & 15) || !dual_epi_ok(e1) || !dual_epi_ok(e2))
        return AIMB_ERR_UNSUPPORTED;
    const int v1 = pick_variant(e1), v2 = pick_variant(e2);
    const int bn2 = (N2 == 192) ? 192 : (N2 == 256 ? 256 : 0);
    if (N1 % 256 || !bn2) return AIMB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb, tb2;
    int rc = make_tmap_bf16(&ta, A, M, K, lda, BM);
    if (rc) return rc;
    if ((rc = make_tmap_bf16(&tb, W1, N1, K, ldw1, 256))) return rc;
    if ((rc = make_tmap_bf16(&tb2, W2, N2, K, ldw2, bn2))) return rc;
    DualExtra ex{};
    const int nt1 = N1 / 256, nt2 = 1, KB = K / BK;
    if (v1 == 1 && v2 == 2) {
        if (bn2 == 192) return launch_dual_v<256, 192, 1, 2, false>(ta, tb, ta, tb2, e1, e2, ex, (int)M, nt1, nt2, KB, 0, s);
        return launch_dual_v<256, 256, 1, 2, false>(ta, tb, ta, tb2, e1, e2, ex, (int)M, nt1, nt2, KB, 0, s);
    }
    if (v1 == 5 && v2 == 6) {
        if (bn2 == 192) return launch_dual_v<256, 192, 5, 6, false>(ta, tb, ta, tb2, e1, e2, ex, (int)M, nt1, nt2, KB, 0, s);
        return launch_dual_v<256, 256, 5, 6, false>(ta, tb, ta, tb2, e1, e2, ex, (int)M, nt1, nt2, KB, 0, s);
    }
    return AIMB_ERR_UNSUPPORTED;
}

// KCAT: A1 [M, K1], W1 [N, K1], A2 [M, K2], W2 [N, K2] -> one [M, N] output.
int gemm_dual_kcat_launch(const void* A1, int64_t lda1, const void* W1, int64_t ldw1, const void* A2, int64_t lda2, const void* W2,
                          int64_t ldw2, const EpiParams& e, const DualExtra& ex, int64_t M, int N, int K1, int K2, cudaStream_t s) {
    if (M < BM || M >= (1ll << 31) || K1 % BK || K2 % BK || (lda1 % 8) || (lda2 % 8) || (ldw1 % 8) || (ldw2 % 8) ||
        ((uintptr_t)A1 & 15) || ((uintptr_t)A2 & 15) || ((uintptr_t)W1 & 15) || ((uintptr_t)W2 & 15) || !dual_epi_ok(e))
        return AIMB_ERR_UNSUPPORTED;
    const int v = pick_variant(e);
    if (v != 0 && v != 3) return AIMB_ERR_UNSUPPORTED;
    // tile width: waves x per-tile cost as pick_bn (N = 768 -> 192: 2.68 waves of 3; N = 1024 -> 256)
    int bn = 0; double best = 1e30;
    const int64_t mt = (M + BM - 1) / BM;
    for (int c : {256, 192}) {
        if (N % c) continue;
        const int64_t tiles = mt * (N / c), waves = (tiles + num_sms() - 1) / num_sms();
        const double cost = (double)waves * ((double)((K1 + K2) / BK) * c + 1536.0);
        if (cost < best - 1e-9) { best = cost; bn = c; }
    }
    if (!bn) return AIMB_ERR_UNSUPPORTED;
    CUtensorMap ta, tb, ta2, tb2;
    int rc = make_tmap_bf16(&ta, A1, M, K1, lda1, BM);
    if (rc) return rc;
    if ((rc = make_tmap_bf16(&tb, W1, N, K1, ldw1, bn))) return rc;
    if ((rc = make_tmap_bf16(&ta2, A2, M, K2, lda2, BM))) return rc;
    if ((rc = make_tmap_bf16(&tb2, W2, N, K2, ldw2, bn))) return rc;
    const int nt = N / bn, KB1 = K1 / BK, KB2 = K2 / BK;
    if (bn == 256) {
        if (v == 0) return launch_dual_v<256, 256, 0, 0, true>(ta, tb, ta2, tb2, e, e, ex, (int)M, nt, 0, KB1, KB2, s);
        return launch_dual_v<256, 256, 3, 3, true>(ta, tb, ta2, tb2, e, e, ex, (int)M, nt, 0, KB1, KB2, s);
    }
    if (v == 0) return launch_dual_v<192, 192, 0, 0, true>(ta, tb, ta2, tb2, e, e, ex, (int)M, nt, 0, KB1, KB2, s);
    return launch_dual_v<192, 192, 3, 3, true>(ta, tb, ta2, tb2, e, e, ex, (int)M, nt, 0, KB1, KB2, s);
}

}  // namespace aimb

using namespace aimb;

extern "C" int aimb_gemm_dual(int32_t mode, const void* A1, int64_t lda1, const void* W1, int64_t ldw1, const void* A2, int64_t lda2,
                              const void* W2, int64_t ldw2, const aimb_epilogue_t* epi1, const aimb_epilogue_t* epi2,
                              const void* bias2, const float* bias2_row_scale, int32_t bias2_row_mod, float bias2_scale, int64_t M,
                              int32_t N1, int32_t N2, int32_t K1, int32_t K2, int32_t dtype, void* stream) {
    if (!A1 || !W1 || !W2 || !epi1 || !epi1->out || M < 0 || N1 <= 0 || K1 <= 0 || lda1 < K1 || ldw1 < K1) return AIMB_ERR_ARG;
    if (dtype != AIMB_BF16) return AIMB_ERR_UNSUPPORTED;
    if (M == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == AIMB_DUAL_NCAT) {
        if (!epi2 || !epi2->out || N2 <= 0 || ldw2 < K1 || bias2) return AIMB_ERR_ARG;
        return gemm_dual_ncat_launch(A1, lda1, W1, ldw1, W2, ldw2, make_epi(epi1, N1), make_epi(epi2, N2), M, N1, N2, K1, s);
    }
    if (mode == AIMB_DUAL_KCAT) {
        if (!A2 || K2 <= 0 || lda2 < K2 || ldw2 < K2) return AIMB_ERR_ARG;
        DualExtra ex{};
        ex.bias2 = bias2; ex.bias2_row_scale = bias2 ? bias2_row_scale : nullptr; ex.bias2_scale = bias2_scale;
        ex.bias2_row_mod = bias2_row_mod > 0 ? bias2_row_mod : 1;
        return gemm_dual_kcat_launch(A1, lda1, W1, ldw1, A2, lda2, W2, ldw2, make_epi(epi1, N1), ex, M, N1, K1, K2, s);
    }
    return AIMB_ERR_ARG;
}

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
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, const int n,
                   const int heads, const int D, const int nprob, const int MT, const int NKP, const int NST, long long* __restrict__ tl) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int KB = (NKP + BOXR - 1) / BOXR;                 // 64-row boxes of K (and of V)
    const int STAGE_B = MT * QTILE_B + 2 * KB * BOXB;
    // bench_tools only: event timeline of CTA 0 (clock64 per unit and event), see bench_tools/attn_tc_timeline.py
    auto mark = [&](int u, int ev) { if (tl && blockIdx.x == 0 && (threadIdx.x & 31) == 0) tl[u * 16 + ev] = clock64(); };
    uint8_t* sO = smem + NST * STAGE_B;                     // output staging: one [128 rows][64] box per softmax warpgroup
    FwdBars* bars = reinterpret_cast<FwdBars*>(sO + 2 * QTILE_B);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
    const int nloc = (nprob - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // problems of this CTA
    const int U = nloc * MT;                                                             // units of this CTA
    // TMEM columns: S / P of warpgroup g at g * RS; O in its own 64 columns when both score regions leave room for it
    // (then the next S may be issued without waiting for the output to be read), else in columns 128..191 of the region
    const bool OSEP = 2 * NKP + HD <= 512;
    const int RS = OSEP ? NKP : 256;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tm);
        ptx::prefetch_tmap(&tmO);
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
            // the tile leaves through shared memory: a thread writes its 128-byte row as swizzled 16-byte chunks and one TMA
            // store per unit writes the [128 rows][64] box; the tensor map clips the rows >= n.  (32-byte stores straight from
            // the registers measured ~1000 clk per unit: every warp instruction touched 32 different lines.)
            const bool storer = wq == 0 && lane == 0;
            if (it > 0) {                                         // the previous unit's store has finished reading the box
                if (storer) ptx::bulk_wait_read0();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            }
            if (active) {
                const float inv = __fdividef(1.f, sum);
                const int rr = wq * 32 + lane;
                const uint32_t base = ptx::smem_u32(sO + g * QTILE_B) + rr * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t (&src)[32] = i < 4 ? va : vb;
                    const int j0 = (i & 3) * 8;
                    const uint32_t w0 = pack_bf16(__uint_as_float(src[j0]) * inv, __uint_as_float(src[j0 + 1]) * inv);
                    const uint32_t w1 = pack_bf16(__uint_as_float(src[j0 + 2]) * inv, __uint_as_float(src[j0 + 3]) * inv);
                    const uint32_t w2 = pack_bf16(__uint_as_float(src[j0 + 4]) * inv, __uint_as_float(src[j0 + 5]) * inv);
                    const uint32_t w3 = pack_bf16(__uint_as_float(src[j0 + 6]) * inv, __uint_as_float(src[j0 + 7]) * inv);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + ((i ^ (rr & 7)) << 4)), "r"(w0), "r"(w1), "r"(w2),
                                 "r"(w3) : "memory");
                }
                if (lse && row < n) lse[((int64_t)f * heads + h) * n + row] = mx * SCALE + __logf(sum);
            }
            ptx::fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            if (storer) {
                ptx::tma_store_3d(&tmO, ptx::smem_u32(sO + g * QTILE_B), h * HD, mt * 128, f);
                ptx::bulk_commit();
            }
            if (wq == 0) mark(u, 5);
        }
    }
    if (warp >= 4 && (warp & 3) == 0 && lane == 0) ptx::bulk_wait0();      // shared memory must outlive the last TMA stores
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------ backward
// attn_bwd_tc_kernel: one (frame, head) problem at a time per persistent CTA, keys on the TMEM lanes ("transposed"
// score layout), so that both dV = P^T dO and dK = dS^T Q take their left operand without a transpose:
//   S^T  = K_kb Q_h^T       (SS, 128 keys x Nh queries x 64)      TMEM columns   0..127
//   dP^T = V_kb dO_h^T      (SS)                                  TMEM columns 128..255
//   math : p = exp2(s * scale * log2e - lse2[q]),  ds = p * (dp - delta[q])      (thread = one key lane)
//          P^T (bf16) -> TMEM columns 0..63 (over S^T);  dS^T (bf16) -> shared memory, 128B-swizzled [key][64 q] boxes
//   dV_kb += P^T dO_h       (TS: A = P^T from TMEM, B = dO MN-major)               columns 256..319
//   dK_kb += dS^T Q_h       (SS: A = the dS^T boxes read K-major,  B = Q MN-major)  columns 320..383
//   dQ_h  += dS K_kb        (SS: A = the SAME dS^T boxes read MN-major, B = K MN-major)   columns 384 + 64 h
// for the key blocks kb (128 lanes each) and query halves h (Nh = 128, then the rest) of the problem: all 512 TMEM
// columns are in use, nothing is reduced through global memory and there are no atomics.  delta = rowsum(dO o O) and
// lse * log2e of the NEXT problem are prepared by two otherwise idle warps while the current problem computes; K, Q
// and dO are double-buffered in shared memory (V single): 7 x 26 KB + 32 KB of dS^T at n = 197.
constexpr int BWD_THREADS = 384;
constexpr float LOG2E = 1.4426950408889634f;

struct BwdBars {
    uint64_t in_full[2], in_empty[2], v_full[2], v_empty[2], dl_full[2];
    uint64_t sdp_full[2], pds_full[2], dvk_full, acc_empty, dq_full, dq_empty;
    uint32_t tmem_ptr;
};

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 u = __bfloat1622float2(x[j]), v = __bfloat1622float2(y[j]);
        acc = fmaf(u.x, v.x, acc);
        acc = fmaf(u.y, v.y, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmV1,
                   const __grid_constant__ CUtensorMap tmOUT, const bf16* __restrict__ o, const float* __restrict__ lse, const int n,
                   const int heads, const int D, const int nprob, const int NQP, const int NSTG, long long* __restrict__ tl) {
    pdl_trigger();
    auto mark = [&](uint32_t s, int ev) { if (tl && blockIdx.x == 0 && (threadIdx.x & 31) == 0) tl[s * 16 + ev] = clock64(); };
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int MATB = NQP * 128;                       // one [NQP rows][64] bf16 operand
    const int BR = NQP >> 1;                          // TMA box rows (two boxes per operand)
    const int KBL = (n + 127) >> 7;                   // key blocks (TMEM lane blocks)
    const int NH = (NQP + 127) >> 7;                  // query halves
    // shared memory: K[NSTG] Q[NSTG] dO[NSTG] V dS^T(32 KB) dl[2][2][256] barriers   (NSTG = 2 while it fits: n <= 208)
    uint8_t* sK = smem;
    uint8_t* sQ = smem + NSTG * MATB;
    uint8_t* sG = smem + 2 * NSTG * MATB;
    uint8_t* sV = smem + 3 * NSTG * MATB;
    uint8_t* sS = smem + (3 * NSTG + 1) * MATB;
    auto stage_of = [&](int k) { return NSTG == 2 ? (k & 1) : 0; };
    auto phase_of = [&](int k) { return (uint32_t)(NSTG == 2 ? (k >> 1) : k) & 1u; };
    float* dl = reinterpret_cast<float*>(sS + 32768);
    BwdBars* bars = reinterpret_cast<BwdBars*>(sS + 32768 + 4096);
    const int warp = ptx::warp_id_uniform(), lane = threadIdx.x & 31;
    const int nloc = (nprob - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmQKV);
        ptx::prefetch_tmap(&tmDO);
        ptx::prefetch_tmap(&tmOUT);
        ptx::prefetch_tmap(&tmV0);
        for (int i = 0; i < 2; ++i) {
            // in_empty: the MMA warp's commit + the storer thread, once the dQ TMA store has read the boxes staged in K / Q
            ptx::mbar_init(&bars->in_full[i], 1); ptx::mbar_init(&bars->in_empty[i], 2); ptx::mbar_init(&bars->dl_full[i], 2);
        }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars->v_full[i], 1); ptx::mbar_init(&bars->v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&bars->sdp_full[i], 1); ptx::mbar_init(&bars->pds_full[i], 4); }
        ptx::mbar_init(&bars->dvk_full, 1); ptx::mbar_init(&bars->acc_empty, 8);
        ptx::mbar_init(&bars->dq_full, 1); ptx::mbar_init(&bars->dq_empty, 8);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<512>(&bars->tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, bars->tmem_ptr, 0);
    pdl_wait();
    if (warp == 0) {
        // ---- TMA producer
        for (int k = 0; k < nloc; ++k) {
            const int pi = blockIdx.x + k * gridDim.x;
            const int f = pi / heads, h = pi - f * heads;
            const int st = stage_of(k);
            ptx::mbar_wait(&bars->in_empty[st], phase_of(k) ^ 1u);
            ptx::mbar_arrive_expect_tx_e(&bars->in_full[st], (uint32_t)(3 * MATB));
            for (int b = 0; b < 2; ++b) {
                ptx::tma_load_3d_e(sK + st * MATB + b * BR * 128, &tmQKV, &bars->in_full[st], D + h * HD, b * BR, f);
                ptx::tma_load_3d_e(sQ + st * MATB + b * BR * 128, &tmQKV, &bars->in_full[st], h * HD, b * BR, f);
                ptx::tma_load_3d_e(sG + st * MATB + b * BR * 128, &tmDO, &bars->in_full[st], h * HD, b * BR, f);
            }
            // V is single-buffered, in two halves (the 128-row key blocks): half kb is dead after the last dP^T of key block
            // kb of the previous problem, so its successor streams in one to three steps before it is needed
            ptx::mbar_wait(&bars->v_empty[0], ((uint32_t)k & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx_e(&bars->v_full[0], (uint32_t)(min(NQP, 128) * 128));
            ptx::tma_load_3d_e(sV, &tmV0, &bars->v_full[0], 2 * D + h * HD, 0, f);
            if (KBL > 1) {
                ptx::mbar_wait(&bars->v_empty[1], ((uint32_t)k & 1u) ^ 1u);
                ptx::mbar_arrive_expect_tx_e(&bars->v_full[1], (uint32_t)((NQP - 128) * 128));
                ptx::tma_load_3d_e(sV + 16384, &tmV1, &bars->v_full[1], 2 * D + h * HD, 128, f);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (warp-uniform)
        const uint32_t idesc_acc = ptx::umma_idesc_bf16(128, HD, 0, 1);     // dV, dK: A K-major (TMEM / smem), B MN-major
        const uint32_t idesc_dq = ptx::umma_idesc_bf16(128, HD, 1, 1);      // dQ: A and B MN-major
        const uint32_t aK = ptx::smem_u32(sK), aQ = ptx::smem_u32(sQ), aG = ptx::smem_u32(sG), aV = ptx::smem_u32(sV),
                       aS = ptx::smem_u32(sS);
        uint32_t s = 0, s1 = 0, c = 0;                                      // step / second-half-step / key-block counters (barrier phases)
        for (int k = 0; k < nloc; ++k) {
            const int st = stage_of(k);
            ptx::mbar_wait(&bars->in_full[st], phase_of(k));
            for (int kb = 0; kb < KBL; ++kb) {
                const int KSkb = (min(128, NQP - kb * 128)) >> 4;           // 16-key steps of this key block
                for (int h = 0; h < NH; ++h, ++s) {
                    const int Nh = min(128, NQP - h * 128);
                    const int KSh = Nh >> 4;
                    // the query chunks of a step are produced, consumed and fed back in two halves (one per math warpgroup):
                    // the second half's S^T / dP^T MMAs run while the first half is already in the math warps, and the first
                    // half's dV / dK MMAs run while the second half is still there
                    const int cs = (KSh + 1) >> 1;                          // 16-query chunks of the first half
                    const int NA = cs << 4, NB = Nh - NA;
                    ptx::tc_fence_after();
                    {   // S^T = K_kb Q_h^T, dP^T = V_kb dO_h^T
                        const uint64_t ka = ptx::umma_desc_kmajor_sw128(aK + st * MATB + kb * 16384);
                        const uint64_t qb = ptx::umma_desc_kmajor_sw128(aQ + st * MATB + h * 16384);
                        const uint64_t va = ptx::umma_desc_kmajor_sw128(aV + kb * 16384);
                        const uint64_t gb = ptx::umma_desc_kmajor_sw128(aG + st * MATB + h * 16384);
                        const uint32_t idesc_a = ptx::umma_idesc_bf16(128, NA);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16_e(tb, ka + 2 * kk, qb + 2 * kk, idesc_a, kk ? 1u : 0u);
                        if (h == 0) { ptx::mbar_wait(&bars->v_full[kb], (uint32_t)k & 1u); ptx::tc_fence_after(); }
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16_e(tb + 128, va + 2 * kk, gb + 2 * kk, idesc_a, kk ? 1u : 0u);
                        ptx::umma_commit_e(&bars->sdp_full[0]);
                        if (NB > 0) {
                            const uint32_t idesc_b = ptx::umma_idesc_bf16(128, NB);
                            const uint64_t ro = (uint64_t)(NA * 8);                // NA query rows further: NA * 128 B, in 16-byte units
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16_e(tb + NA, ka + 2 * kk, qb + ro + 2 * kk, idesc_b, kk ? 1u : 0u);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16_e(tb + 128 + NA, va + 2 * kk, gb + ro + 2 * kk, idesc_b, kk ? 1u : 0u);
                            ptx::umma_commit_e(&bars->sdp_full[1]);
                        }
                        if (h == NH - 1) ptx::umma_commit_e(&bars->v_empty[kb]);    // last use of this half of V
                        mark(s, 8);
                    }
                    const uint64_t g_mn = ptx::umma_desc_mnmajor_sw128(aG + st * MATB + h * 16384, 16384);
                    const uint64_t q_mn = ptx::umma_desc_mnmajor_sw128(aQ + st * MATB + h * 16384, 16384);
                    const uint64_t k_mn = ptx::umma_desc_mnmajor_sw128(aK + st * MATB + kb * 16384, 16384);
                    const uint64_t s_k = ptx::umma_desc_kmajor_sw128(aS);
                    const uint64_t s_mn = ptx::umma_desc_mnmajor_sw128(aS, 16384);
                    auto dv_dk = [&](int j0, int j1) {
#pragma unroll 4
                        for (int j = j0; j < j1; ++j) {     // dV_kb += P^T dO_h : 16 queries per MMA; P^T chunk j sits where its math warp put it
                            const uint32_t p_col = j < cs ? 8 * j : 16 * cs + 8 * (j - cs);
                            ptx::umma_bf16_ts_e(tb + 256, tb + p_col, g_mn + (uint64_t)(j * 128), idesc_acc, (h | j) ? 1u : 0u);
                        }
#pragma unroll 4
                        for (int j = j0; j < j1; ++j)       // dK_kb += dS^T Q_h : A = dS^T boxes, K-major (64 queries per box)
                            ptx::umma_bf16_e(tb + 320, s_k + (uint64_t)((j >> 2) * 1024 + (j & 3) * 2), q_mn + (uint64_t)(j * 128),
                                             idesc_acc, (h | j) ? 1u : 0u);
                    };
                    ptx::mbar_wait(&bars->pds_full[0], s & 1u);             // first half: P^T in TMEM, dS^T in shared memory
                    mark(s, 9);
                    if (h == 0 && c > 0) ptx::mbar_wait(&bars->acc_empty, (c - 1) & 1u);      // dV / dK columns drained
                    if (kb == 0 && h == 0 && k > 0) ptx::mbar_wait(&bars->dq_empty, (uint32_t)(k - 1) & 1u);
                    ptx::tc_fence_after();
                    dv_dk(0, cs);
                    if (NB > 0) {
                        ptx::mbar_wait(&bars->pds_full[1], s1 & 1u);
                        ++s1;
                        ptx::tc_fence_after();
                        dv_dk(cs, KSh);
                    }
#pragma unroll 4
                    for (int j = 0; j < KSkb; ++j)       // dQ_h += dS K_kb : A = the same boxes, MN-major (16 keys per MMA)
                        ptx::umma_bf16_e(tb + 384 + h * 64, s_mn + (uint64_t)(j * 128), k_mn + (uint64_t)(j * 128), idesc_dq,
                                         (kb | j) ? 1u : 0u);
                    mark(s, 10);
                    if (h == NH - 1) { ptx::umma_commit_e(&bars->dvk_full); ++c; }
                }
            }
            ptx::umma_commit_e(&bars->dq_full);
            ptx::umma_commit_e(&bars->in_empty[st]);
        }
    } else if (warp < 4) {
        // ---- delta / lse warps: delta[q] = sum_d dO[q,d] O[q,d] and lse2[q] = lse[q] * log2(e) of problem k (one ahead)
        const int t = (warp - 2) * 32 + lane;
        for (int k = 0; k < nloc; ++k) {
            const int pi = blockIdx.x + k * gridDim.x;
            const int f = pi / heads, h = pi - f * heads;
            const int st = stage_of(k);
            ptx::mbar_wait(&bars->in_full[st], phase_of(k));
            const uint8_t* g = sG + st * MATB;
            float* l2 = dl + (k & 1) * 512;
            for (int q = t; q < NQP; q += 64) {
                float acc = 0.f, lv = INFINITY;           // padded queries: lse2 = +inf -> p = 0
                if (q < n) {
                    const uint4* op = reinterpret_cast<const uint4*>(o + ((int64_t)f * n + q) * D + h * HD);
                    uint4 ov[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) ov[i] = op[i];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        acc += dot8(*reinterpret_cast<const uint4*>(g + q * 128 + ((i ^ (q & 7)) << 4)), ov[i]);
                    lv = lse[((int64_t)f * heads + h) * n + q] * LOG2E;
                }
                l2[q] = lv;
                l2[256 + q] = acc;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->dl_full[k & 1]);
        }
    } else {
        // ---- math warps: warp%4 = TMEM lane quadrant, (warp-4)/4 = which half of the query chunks
        const int wq = warp & 3, ch = (warp - 4) >> 2;
        const uint32_t trow = tb + ((uint32_t)(wq * 32) << 16);
        const int r = wq * 32 + lane;                                     // key row inside the key block
        const uint32_t sS_row = ptx::smem_u32(sS) + r * 128;
        const bool storer = threadIdx.x == 128;                           // issues (and drains) the TMA stores
        bool st_pending = false;                                          // a TMA store may still be reading the dS^T buffer
        uint32_t s = 0, s1 = 0, c = 0;
        // results leave through the (then idle) dS^T buffer: a thread writes its row as swizzled 16-byte chunks and the two
        // [128 rows][64] boxes go out as TMA stores whose tensor map clips rows >= n (padded keys / queries are never
        // written).  (Plain st.global from 256 threads measured 3-4x slower here: the CTAs run in lockstep and their store
        // bursts collide; the asynchronous stores are absorbed.)
        auto stage_row = [&](int box, const uint32_t (&v)[16]) {          // 32 columns (ch-th half) of row r of box `box`
            const uint32_t base = sS_row + (box << 14);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((ch * 4 + i) ^ (r & 7)) << 4)), "r"(v[4 * i]),
                             "r"(v[4 * i + 1]), "r"(v[4 * i + 2]), "r"(v[4 * i + 3]) : "memory");
        };
        int release_stage = -1;                                           // operand stage whose K / Q buffers carry a dQ store
        auto drain_store = [&]() {                                        // before anything rewrites the dS^T buffer
            if (st_pending) {
                if (storer) {
                    ptx::bulk_wait_read0();
                    if (release_stage >= 0) ptx::mbar_arrive(&bars->in_empty[release_stage]);
                }
                release_stage = -1;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                st_pending = false;
            }
        };
        for (int k = 0; k < nloc; ++k) {
            const int pi = blockIdx.x + k * gridDim.x;
            const int f = pi / heads, h_ = pi - f * heads;
            const float* l2 = dl + (k & 1) * 512;
            ptx::mbar_wait(&bars->dl_full[k & 1], (uint32_t)(k >> 1) & 1u);
            for (int kb = 0; kb < KBL; ++kb) {
                const uint32_t kmask = kb * 128 + r < n ? 0xffffffffu : 0u;
                for (int h = 0; h < NH; ++h, ++s) {
                    const int Nh = min(128, NQP - h * 128);
                    const int NC = Nh >> 4;                               // 16-query chunks of this half
                    const int c0 = ch ? (NC + 1) >> 1 : 0, c1 = ch ? NC : (NC + 1) >> 1;
                    if (ch == 0) {
                        ptx::mbar_wait(&bars->sdp_full[0], s & 1u);
                    } else if (c0 < c1) {                              // the second half exists only when there are >= 2 chunks
                        ptx::mbar_wait(&bars->sdp_full[1], s1 & 1u);
                        ++s1;
                    }
                    ptx::tc_fence_after();
                    if (warp == 4) mark(s, 0);
                    uint32_t sv[2][16], dv[2][16];
                    if (c0 < c1) {
                        ptx::tmem_ld_32x32b_x16(trow + c0 * 16, sv[0]);
                        ptx::tmem_ld_32x32b_x16(trow + 128 + c0 * 16, dv[0]);
                    }
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int cc = c0 + ci;
                        if (cc < c1) {
                            const uint32_t (&sc)[16] = sv[ci & 1];
                            const uint32_t (&dc)[16] = dv[ci & 1];
                            const float* lq = l2 + h * 128 + cc * 16;
                            float lv[16], dd[16];                       // per-query constants: broadcast reads, issued under the TMEM loads
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
                                *reinterpret_cast<float4*>(lv + 4 * m) = *reinterpret_cast<const float4*>(lq + 4 * m);
                                *reinterpret_cast<float4*>(dd + 4 * m) = *reinterpret_cast<const float4*>(lq + 256 + 4 * m);
                            }
                            ptx::tmem_wait_ld();
                            if (ci < 3 && cc + 1 < c1) {                  // next chunk's scores are in flight while this one computes
                                ptx::tmem_ld_32x32b_x16(trow + (cc + 1) * 16, sv[(ci + 1) & 1]);
                                ptx::tmem_ld_32x32b_x16(trow + 128 + (cc + 1) * 16, dv[(ci + 1) & 1]);
                            }
                            uint32_t dsp[8], pkc[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float p0 = ex2_ftz(fmaf(__uint_as_float(sc[2 * j]), SCALE_LOG2, -lv[2 * j]));
                                float p1 = ex2_ftz(fmaf(__uint_as_float(sc[2 * j + 1]), SCALE_LOG2, -lv[2 * j + 1]));
                                float d0 = p0 * (__uint_as_float(dc[2 * j]) - dd[2 * j]);
                                float d1 = p1 * (__uint_as_float(dc[2 * j + 1]) - dd[2 * j + 1]);
                                pkc[j] = pack_bf16(p0, p1) & kmask;                  // padded keys contribute nothing
                                dsp[j] = pack_bf16(d0, d1) & kmask;
                            }
                            if (ci == 0) drain_store();     // the previous results' TMA store has finished reading the buffer
                            // dS^T -> shared memory: box (cc*16)/64 of 64 queries, row r, two 16-byte chunks (128B swizzle)
                            const uint32_t base = sS_row + ((cc >> 2) << 14);
                            const int i0 = (cc & 3) * 2;
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((i0) ^ (r & 7)) << 4)), "r"(dsp[0]),
                                         "r"(dsp[1]), "r"(dsp[2]), "r"(dsp[3]) : "memory");
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((i0 + 1) ^ (r & 7)) << 4)), "r"(dsp[4]),
                                         "r"(dsp[5]), "r"(dsp[6]), "r"(dsp[7]) : "memory");
                            // P^T (bf16) over S^T columns that THIS warp has already consumed: the first half of the chunks
                            // packs from column 0, the second half from the first column of its own range (see p_col)
                            ptx::tmem_st_32x32b_x8(trow + (ch ? 16 * c0 + 8 * (cc - c0) : 8 * cc), pkc);
                        } else if (ci == 0) {
                            drain_store();                  // warps without a chunk (tiny n) still take part in the barrier
                        }
                    }
                    if (warp == 4) mark(s, 1);
                    ptx::tmem_wait_st();
                    ptx::fence_proxy_async();
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0 && (ch == 0 || c0 < c1)) ptx::mbar_arrive(&bars->pds_full[ch]);
                    if (warp == 4) mark(s, 2);
                    if (h == NH - 1) {
                        // ---- dV_kb, dK_kb complete (and the MMAs that read the dS^T buffer with them)
                        ptx::mbar_wait(&bars->dvk_full, c & 1u);
                        ++c;
                        ptx::tc_fence_after();
                        if (warp == 4) mark(s, 3);
                        uint32_t a[32], b[32];
                        ptx::tmem_ld_32x32b_x32(trow + 256 + ch * 32, a);
                        ptx::tmem_ld_32x32b_x32(trow + 320 + ch * 32, b);
                        ptx::tmem_wait_ld();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&bars->acc_empty);
                        uint32_t ov[16], ok[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            ov[j] = pack_bf16(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
                            ok[j] = pack_bf16(__uint_as_float(b[2 * j]) * SCALE, __uint_as_float(b[2 * j + 1]) * SCALE);
                        }
                        stage_row(0, ok);
                        stage_row(1, ov);
                        ptx::fence_proxy_async();
                        asm volatile("bar.sync 2, 256;" ::: "memory");
                        if (storer) {
                            ptx::tma_store_3d(&tmOUT, ptx::smem_u32(sS), D + h_ * HD, kb * 128, f);
                            ptx::tma_store_3d(&tmOUT, ptx::smem_u32(sS) + 16384, 2 * D + h_ * HD, kb * 128, f);
                            ptx::bulk_commit();
                        }
                        st_pending = true;
                        if (warp == 4) mark(s, 4);
                    }
                }
            }
            // ---- dQ of the whole problem
            ptx::mbar_wait(&bars->dq_full, (uint32_t)k & 1u);
            ptx::tc_fence_after();
            if (warp == 4) mark(s - 1, 5);
            uint32_t qa[32], qb2[32];
            ptx::tmem_ld_32x32b_x32(trow + 384 + ch * 32, qa);
            if (NH > 1) ptx::tmem_ld_32x32b_x32(trow + 448 + ch * 32, qb2);
            ptx::tmem_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars->dq_empty);
            // dQ is staged in the K and Q buffers of this problem's operand stage (dead once dq_full has arrived) so that it
            // does not wait for the dK / dV store that has just been issued from the dS^T buffer; the stage goes back to the
            // producer when the store has read it (deferred to the next drain, or at once when there is a single stage)
            const int st = stage_of(k);
            const uint32_t sq0 = ptx::smem_u32(sK + st * MATB) + r * 128, sq1 = ptx::smem_u32(sQ + st * MATB) + r * 128;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                if (hb < NH) {
                    uint32_t oq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t x0 = hb ? qb2[2 * j] : qa[2 * j], x1 = hb ? qb2[2 * j + 1] : qa[2 * j + 1];
                        oq[j] = pack_bf16(__uint_as_float(x0) * SCALE, __uint_as_float(x1) * SCALE);
                    }
                    const uint32_t base = hb ? sq1 : sq0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(base + (((ch * 4 + i) ^ (r & 7)) << 4)), "r"(oq[4 * i]),
                                     "r"(oq[4 * i + 1]), "r"(oq[4 * i + 2]), "r"(oq[4 * i + 3]) : "memory");
                }
            }
            ptx::fence_proxy_async();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (storer) {
                ptx::tma_store_3d(&tmOUT, ptx::smem_u32(sK + st * MATB), h_ * HD, 0, f);
                if (NH > 1) ptx::tma_store_3d(&tmOUT, ptx::smem_u32(sQ + st * MATB), h_ * HD, 128, f);
                ptx::bulk_commit();
                if (NSTG == 1) { ptx::bulk_wait_read0(); ptx::mbar_arrive(&bars->in_empty[st]); }
            }
            st_pending = true;
            release_stage = NSTG == 1 ? -1 : st;
        }
        if (storer) ptx::bulk_wait0();              // shared memory must outlive the last TMA store
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tb);
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
    const void* ptr; int64_t slabs, rows, cols; int box_rows;
    bool operator==(const Key3& o) const {
        return ptr == o.ptr && slabs == o.slabs && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
    }
};
struct Key3Hash {
    size_t operator()(const Key3& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ std::hash<int64_t>()(k.slabs);
        h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
        h = h * 1000003u ^ std::hash<int64_t>()(k.cols);
        h = h * 1000003u ^ std::hash<int>()(k.box_rows);
        return h;
    }
};

// bf16 [slabs, rows, cols] contiguous; box = 64 columns x box_rows rows x 1 slab, 128B swizzle, zero fill out of bounds
static int make_tmap3(CUtensorMap* out, const void* ptr, int64_t slabs, int64_t rows, int64_t cols, int box_rows = BOXR) {
    static std::mutex mu;
    static std::unordered_map<Key3, CUtensorMap, Key3Hash> cache;
    Key3 key{ptr, slabs, rows, cols, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return AIMB_OK; }
    }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return AIMB_ERR_DRIVER;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
    cuuint64_t gstr[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
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

static int sm_count() { return device_sm_count(); }

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
    const int NST = (2 * STAGE_B + 2 * QTILE_B + 2048 <= 227 * 1024) ? 2 : 1;
    const int smem = NST * STAGE_B + 2 * QTILE_B + 1024 + 256;
    CUtensorMap to;
    rc = make_tmap3(&to, o, frames, n, D, 128);
    if (rc) return rc;
    AIMB_SET_SMEM_ATTR(227 * 1024, attn_fwd_tc_kernel);
    const int nprob = frames * heads;
    const int sms = sm_count();
    const int waves = (nprob + sms - 1) / sms;
    const int grid = (nprob + waves - 1) / waves;
    launch_k(attn_fwd_tc_kernel, dim3(grid), dim3(FWD_THREADS), (size_t)smem, s, tm, to, lse, n, heads, D, nprob, MT, NKP, NST, g_attn_timeline);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}


int attn_spatial_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames, int n,
                        int heads, cudaStream_t s) {
    using namespace atc;
    if (!attn_spatial_tc_supported(n, heads)) return AIMB_ERR_UNSUPPORTED;
    const int D = heads * HD;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)o & 15) || ((uintptr_t)d_o & 15) || ((uintptr_t)d_qkv & 31)) return AIMB_ERR_ARG;
    const int NQP = (n + 15) & ~15;
    CUtensorMap tq, tg;
    int rc = make_tmap3(&tq, qkv, frames, n, 3 * (int64_t)D, NQP / 2);
    if (rc) return rc;
    rc = make_tmap3(&tg, d_o, frames, n, D, NQP / 2);
    if (rc) return rc;
    CUtensorMap to, tv0, tv1;
    rc = make_tmap3(&to, d_qkv, frames, n, 3 * (int64_t)D, 128);
    if (rc) return rc;
    rc = make_tmap3(&tv0, qkv, frames, n, 3 * (int64_t)D, NQP < 128 ? NQP : 128);
    if (rc) return rc;
    tv1 = tv0;
    if (NQP > 128) {
        rc = make_tmap3(&tv1, qkv, frames, n, 3 * (int64_t)D, NQP - 128);
        if (rc) return rc;
    }
    const int NSTG = (7 * NQP * 128 + 32768 + 4096 + 256 + 1024 <= 227 * 1024) ? 2 : 1;
    const int smem = (3 * NSTG + 1) * NQP * 128 + 32768 + 4096 + 256 + 1024;
    AIMB_SET_SMEM_ATTR(227 * 1024, attn_bwd_tc_kernel);
    const int nprob = frames * heads;
    const int sms = sm_count();
    const int waves = (nprob + sms - 1) / sms;
    const int grid = (nprob + waves - 1) / waves;
    launch_k(attn_bwd_tc_kernel, dim3(grid), dim3(BWD_THREADS), (size_t)smem, s, tq, tg, tv0, tv1, to, (const bf16*)o, lse, n, heads, D,
             nprob, NQP, NSTG, g_attn_timeline);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

}  // namespace aimb

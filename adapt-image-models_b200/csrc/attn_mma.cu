// Fused flash-style spatial attention for bf16 (vit_clip.py:140-156).  One (frame, head) problem = n tokens
// (197 / 257) x head_dim 64; the score matrix never leaves registers.  Tensor-core path: mma.sync m16n8k16 bf16,
// fp32 accumulation and fp32 softmax statistics.  Q/K/V are read in place from the fused QKV buffer [M, 3D]; O is
// written head-major into [M, D] (== permute(2,0,1,3).flatten(2) of the reference) so out_proj consumes it directly.
//
// Work decomposition (round-1 tuning, see profiles/):
//   forward : a CTA owns up to 8 row tiles (16 query rows each, one warp per tile) of one (frame, head); only K and V
//             live in shared memory (64.5 KB for n = 197), the warp's own 16 query rows are loaded straight from
//             global into MMA A-fragments -> 2 CTAs resident per SM, one CTA's load phase overlaps another's math.
//   backward: one CTA per (frame, head) with Q, K, V, dO in shared memory; phase 1 (warp = 16 query rows) recomputes P
//             from the saved LSE and produces dQ, phase 2 (warp = 16 key rows) produces dK, dV.  No atomics, no score
//             matrix in HBM.  (Splitting the phases into separate 2-matrix CTAs was measured slower: 2x smem fills.)
#include <cstdlib>
#include "common.cuh"

namespace aimb {

constexpr int HD = 64;
constexpr int LDS = 72;  // smem row stride in bf16 (144 B): conflict-free ldmatrix
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(s_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(s_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// B fragments for TWO k-steps (k0..k0+31) of an [n][k] smem tile (8 n rows): r0,r1 = k-step 0; r2,r3 = k-step 1
__device__ __forceinline__ void ldb_frag_nk(uint32_t (&r)[4], const bf16* p, int lane) {
    ldsm_x4(r, p + (lane & 7) * LDS + (lane >> 3) * 8);
}
// B fragments for TWO n-tiles (n0..n0+15) of a [k][n] smem tile (16 k rows): r0,r1 = n-tile 0; r2,r3 = n-tile 1
__device__ __forceinline__ void ldb_frag_kn(uint32_t (&r)[4], const bf16* p, int lane) {
    ldsm_x4_t(r, p + ((lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (lane >> 4) * 8);
}

// A fragments (4 k-steps covering the 64 head dims) of 16 rows read straight from global memory.
// p -> (row0, col 0) of the warp's tile, `ld` = row stride in elements; rows >= rows_valid read as zero.
__device__ __forceinline__ void lda_frags_global(uint32_t (&a)[4][4], const bf16* p, int64_t ld, int rows_valid, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const bool v0 = g < rows_valid, v1 = g + 8 < rows_valid;
    const bf16* r0 = p + (int64_t)g * ld + 2 * t;
    const bf16* r1 = r0 + 8 * ld;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = v0 ? *reinterpret_cast<const uint32_t*>(r0 + ks * 16) : 0u;
        a[ks][1] = v1 ? *reinterpret_cast<const uint32_t*>(r1 + ks * 16) : 0u;
        a[ks][2] = v0 ? *reinterpret_cast<const uint32_t*>(r0 + ks * 16 + 8) : 0u;
        a[ks][3] = v1 ? *reinterpret_cast<const uint32_t*>(r1 + ks * 16 + 8) : 0u;
    }
}

// Load `n` rows x 64 bf16 (row stride `ld` elements in global) into smem [npad][LDS], zero the pad rows.
__device__ __forceinline__ void load_tile_async(bf16* dst, const bf16* src, int64_t ld, int n, int npad, int tid, int nthr) {
    for (int e = tid; e < npad * 8; e += nthr) {
        int r = e >> 3, c = (e & 7) * 8;
        if (r < n) cp_async16(dst + r * LDS + c, src + (int64_t)r * ld + c);
        else *reinterpret_cast<uint4*>(dst + r * LDS + c) = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------ forward
template <int MINB>   // resident CTAs per SM the register budget is sized for: 3 when K,V of one head fit three times (n <= 224)
__global__ void __launch_bounds__(256, MINB) attn_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                              float* __restrict__ lse, int n, int heads, int wpc) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int D = heads * HD, ld = 3 * D;
    const int f = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npad = (n + 31) & ~31;
    bf16* sK = reinterpret_cast<bf16*>(smraw);
    bf16* sV = sK + npad * LDS;
    const bf16* base = qkv + (int64_t)f * n * ld + h * HD;
    load_tile_async(sK, base + D, ld, n, npad, tid, blockDim.x);
    load_tile_async(sV, base + 2 * D, ld, n, npad, tid, blockDim.x);
    const int q0 = (blockIdx.y * wpc + warp) * 16;
    uint32_t qa[4][4];
    if (q0 < n) lda_frags_global(qa, base + (int64_t)q0 * ld, ld, n - q0, lane);
    cp_async_wait_all();
    __syncthreads();
    if (q0 >= n) return;
    const int g = lane >> 2, t = lane & 3;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float oacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;

    for (int kc = 0; kc < npad; kc += 32) {
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
                uint32_t kb[4];
                ldb_frag_nk(kb, sK + (kc + nt * 8) * LDS + kp * 32, lane);
                mma16816(s[nt], qa[2 * kp], kb[0], kb[1]);
                mma16816(s[nt], qa[2 * kp + 1], kb[2], kb[3]);
            }
        }
        float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            int key = kc + nt * 8 + 2 * t;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = s[nt][j] * SCALE_LOG2;
                if (key + (j & 1) >= n) v = -INFINITY;
                s[nt][j] = v;
            }
            cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
            cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        // every 32-key chunk that is processed holds at least one valid key (npad - n < 32), so the max is finite
        const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
        const float c0 = ex2_ftz(m0 - mn0), c1 = ex2_ftz(m1 - mn1);
        m0 = mn0; m1 = mn1;
        l0 *= c0; l1 *= c1;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) { oacc[dt][0] *= c0; oacc[dt][1] *= c0; oacc[dt][2] *= c1; oacc[dt][3] *= c1; }
        uint32_t pa[2][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            float p0 = ex2_ftz(s[nt][0] - m0), p1 = ex2_ftz(s[nt][1] - m0);
            float p2 = ex2_ftz(s[nt][2] - m1), p3 = ex2_ftz(s[nt][3] - m1);
            l0 += p0 + p1; l1 += p2 + p3;
            pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
            pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                ldb_frag_kn(vb, sV + (kc + kk * 16) * LDS + dp * 16, lane);
                mma16816(oacc[2 * dp], pa[kk], vb[0], vb[1]);
                mma16816(oacc[2 * dp + 1], pa[kk], vb[2], vb[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    bf16* ob = o + (int64_t)f * n * D + h * HD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        int c = dt * 8 + 2 * t;
        if (r0 < n) *reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * D + c) = pack_bf16(oacc[dt][0] * i0, oacc[dt][1] * i0);
        if (r1 < n) *reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * D + c) = pack_bf16(oacc[dt][2] * i1, oacc[dt][3] * i1);
    }
    if (lse && t == 0) {
        float* lr = lse + ((int64_t)f * heads + h) * n;
        if (r0 < n) lr[r0] = m0 * LN2 + __logf(l0);
        if (r1 < n) lr[r1] = m1 * LN2 + __logf(l1);
    }
}

// A fragment (16 rows x 16 k) of a row-major [row][k] smem tile; p -> element (row0, k0)
__device__ __forceinline__ void lda_frag(uint32_t (&a)[4], const bf16* p, int lane) {
    ldsm_x4(a, p + (lane & 15) * LDS + (lane >> 4) * 8);
}

// ------------------------------------------------------------------------------------------ backward
// One CTA per (frame, head) holding Q, K, V, dO in shared memory (129 KB at n = 197): phase 1 (warp = 16 query
// rows) recomputes P and produces dQ, phase 2 (warp = 16 key rows) recomputes P^T and produces dK, dV.  Measured
// faster than splitting the phases into separate 2-matrix CTAs (which doubles the shared-memory fills).
template <int NW>
__global__ void __launch_bounds__(NW * 32) attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                           const bf16* __restrict__ d_o, const float* __restrict__ lse,
                                                           bf16* __restrict__ d_qkv, int n, int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int D = heads * HD, ld = 3 * D;
    const int f = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int npad = (n + 31) & ~31;
    bf16* sQ = reinterpret_cast<bf16*>(smraw);
    bf16* sK = sQ + npad * LDS;
    bf16* sV = sK + npad * LDS;
    bf16* sG = sV + npad * LDS;                                   // dO
    float* sL = reinterpret_cast<float*>(sG + npad * LDS);        // lse * log2(e)
    float* sDl = sL + npad;                                       // delta
    const bf16* base = qkv + (int64_t)f * n * ld + h * HD;
    const bf16* gb = d_o + (int64_t)f * n * D + h * HD;
    const bf16* ob = o + (int64_t)f * n * D + h * HD;
    bf16* db = d_qkv + (int64_t)f * n * ld + h * HD;
    load_tile_async(sQ, base, ld, n, npad, tid, blockDim.x);
    load_tile_async(sK, base + D, ld, n, npad, tid, blockDim.x);
    load_tile_async(sV, base + 2 * D, ld, n, npad, tid, blockDim.x);
    load_tile_async(sG, gb, D, n, npad, tid, blockDim.x);
    const float* lr = lse + ((int64_t)f * heads + h) * n;
    // padded rows get lse = +inf: exp2(s - inf) = 0 masks padded QUERIES for free (phase 2 needs no predicate)
    for (int i = tid; i < npad; i += blockDim.x) sL[i] = i < n ? lr[i] * LOG2E : INFINITY;
    // sDl_i = -0.125 * delta_i, delta_i = sum_d dO[i,d] * O[i,d].  One THREAD per row with all 16 x 16-byte loads in
    // flight at once (ncu: the former one-warp-per-row loop serialised ~17 global round trips per warp = 19 % of the
    // kernel's stall samples).
    for (int i = tid; i < npad; i += blockDim.x) {
        float acc = 0.f;
        if (i < n) {
            const uint4* gp = reinterpret_cast<const uint4*>(gb + (int64_t)i * D);
            const uint4* op = reinterpret_cast<const uint4*>(ob + (int64_t)i * D);
            uint4 a[8], b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = __ldg(gp + j); b[j] = __ldg(op + j); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t* aw = reinterpret_cast<const uint32_t*>(&a[j]);
                const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float2 fa = unpack_bf16(aw[k]), fb = unpack_bf16(bw[k]);
                    acc = fmaf(fa.x, fb.x, acc);
                    acc = fmaf(fa.y, fb.y, acc);
                }
            }
        }
        sDl[i] = -0.125f * acc;
    }
    cp_async_wait_all();
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    // a warp owns 16 rows (queries in phase 1, keys in phase 2); with more row tiles than warps (n = 257 runs 17
    // tiles on 9 warps: 17 warps would be capped at 96 registers and spill) it makes a second pass.
    // No block-wide barriers below.
    for (int r0 = warp * 16; r0 < n; r0 += nwarps * 16) {

    // ---------------- phase 1: dQ for query rows r0..r0+15
    {
        float dq[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
        const float ls0 = sL[r0 + g], ls1 = sL[r0 + g + 8];
        const float dl0 = sDl[r0 + g], dl1 = sDl[r0 + g + 8];
        uint32_t qa[4][4], ga[4][4];     // this warp's Q and dO rows as A fragments (invariant over the key loop)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            lda_frag(qa[ks], sQ + r0 * LDS + ks * 16, lane);
            lda_frag(ga[ks], sG + r0 * LDS + ks * 16, lane);
        }
        for (int kc = 0; kc < npad; kc += 32) {
            float s[4][4], dp[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { s[nt][j] = 0.f; dp[nt][j] = 0.f; }
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    uint32_t kb[4], vb[4];
                    ldb_frag_nk(kb, sK + (kc + nt * 8) * LDS + kp * 32, lane);
                    ldb_frag_nk(vb, sV + (kc + nt * 8) * LDS + kp * 32, lane);
                    mma16816(s[nt], qa[2 * kp], kb[0], kb[1]);
                    mma16816(s[nt], qa[2 * kp + 1], kb[2], kb[3]);
                    mma16816(dp[nt], ga[2 * kp], vb[0], vb[1]);
                    mma16816(dp[nt], ga[2 * kp + 1], vb[2], vb[3]);
                }
            }
            uint32_t dsa[2][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int key = kc + nt * 8 + 2 * t;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float lsj = (j < 2) ? ls0 : ls1, dlj = (j < 2) ? dl0 : dl1;
                    float p = ex2_ftz(fmaf(s[nt][j], SCALE_LOG2, -lsj));
                    v[j] = p * fmaf(dp[nt][j], 0.125f, dlj);
                }
                if (kc + 32 > n) {                 // only the last key chunk holds padded keys
                    if (key >= n) { v[0] = 0.f; v[2] = 0.f; }
                    if (key + 1 >= n) { v[1] = 0.f; v[3] = 0.f; }
                }
                dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(v[0], v[1]);
                dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v[2], v[3]);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int dpair = 0; dpair < 4; ++dpair) {
                    uint32_t kb[4];
                    ldb_frag_kn(kb, sK + (kc + kk * 16) * LDS + dpair * 16, lane);
                    mma16816(dq[2 * dpair], dsa[kk], kb[0], kb[1]);
                    mma16816(dq[2 * dpair + 1], dsa[kk], kb[2], kb[3]);
                }
            }
        }
        const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            int c = dt * 8 + 2 * t;
            if (ra < n) *reinterpret_cast<uint32_t*>(db + (int64_t)ra * ld + c) = pack_bf16(dq[dt][0], dq[dt][1]);
            if (rb < n) *reinterpret_cast<uint32_t*>(db + (int64_t)rb * ld + c) = pack_bf16(dq[dt][2], dq[dt][3]);
        }
    }
    // ---------------- phase 2: dK, dV for key rows r0..r0+15
    {
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
        uint32_t ka[4][4], va[4][4];     // this warp's K and V rows as A fragments (invariant over the query loop)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            lda_frag(ka[ks], sK + r0 * LDS + ks * 16, lane);
            lda_frag(va[ks], sV + r0 * LDS + ks * 16, lane);
        }
        for (int qc = 0; qc < npad; qc += 32) {
            float s[4][4], dp[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { s[nt][j] = 0.f; dp[nt][j] = 0.f; }
#pragma unroll
                for (int kp = 0; kp < 2; ++kp) {
                    uint32_t qb[4], gbf[4];
                    ldb_frag_nk(qb, sQ + (qc + nt * 8) * LDS + kp * 32, lane);
                    ldb_frag_nk(gbf, sG + (qc + nt * 8) * LDS + kp * 32, lane);
                    mma16816(s[nt], ka[2 * kp], qb[0], qb[1]);      // S^T[key, q]
                    mma16816(s[nt], ka[2 * kp + 1], qb[2], qb[3]);
                    mma16816(dp[nt], va[2 * kp], gbf[0], gbf[1]);   // dP^T[key, q]
                    mma16816(dp[nt], va[2 * kp + 1], gbf[2], gbf[3]);
                }
            }
            uint32_t pa[2][4], dsa[2][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int q = qc + nt * 8 + 2 * t;
                float lq0 = sL[q], lq1 = sL[q + 1], dq0 = sDl[q], dq1 = sDl[q + 1];
                float p[4], v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // padded queries: lse = +inf -> p = 0
                    float lsj = (j & 1) ? lq1 : lq0, dlj = (j & 1) ? dq1 : dq0;
                    p[j] = ex2_ftz(fmaf(s[nt][j], SCALE_LOG2, -lsj));
                    v[j] = p[j] * fmaf(dp[nt][j], 0.125f, dlj);
                }
                pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
                pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
                dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(v[0], v[1]);
                dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(v[2], v[3]);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int dpair = 0; dpair < 4; ++dpair) {
                    uint32_t gb4[4], qb4[4];
                    ldb_frag_kn(gb4, sG + (qc + kk * 16) * LDS + dpair * 16, lane);
                    ldb_frag_kn(qb4, sQ + (qc + kk * 16) * LDS + dpair * 16, lane);
                    mma16816(dv[2 * dpair], pa[kk], gb4[0], gb4[1]);
                    mma16816(dv[2 * dpair + 1], pa[kk], gb4[2], gb4[3]);
                    mma16816(dk[2 * dpair], dsa[kk], qb4[0], qb4[1]);
                    mma16816(dk[2 * dpair + 1], dsa[kk], qb4[2], qb4[3]);
                }
            }
        }
        const int ra = r0 + g, rb = r0 + g + 8;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            int c = dt * 8 + 2 * t;
            if (ra < n) {
                *reinterpret_cast<uint32_t*>(db + (int64_t)ra * ld + D + c) = pack_bf16(dk[dt][0], dk[dt][1]);
                *reinterpret_cast<uint32_t*>(db + (int64_t)ra * ld + 2 * D + c) = pack_bf16(dv[dt][0], dv[dt][1]);
            }
            if (rb < n) {
                *reinterpret_cast<uint32_t*>(db + (int64_t)rb * ld + D + c) = pack_bf16(dk[dt][2], dk[dt][3]);
                *reinterpret_cast<uint32_t*>(db + (int64_t)rb * ld + 2 * D + c) = pack_bf16(dv[dt][2], dv[dt][3]);
            }
        }
    }
    }   // row-tile passes
}

// row tiles of 16 are spread over ceil(tiles / 8) CTAs with an equal number of warps each
static void split_rows(int n, int& nsplit, int& wpc) {
    const int tiles = (n + 15) / 16;
    nsplit = (tiles + 7) / 8;
    wpc = (tiles + nsplit - 1) / nsplit;
}

int attn_spatial_fwd_mma(const void* qkv, void* o, float* lse, int frames, int n, int heads, cudaStream_t s) {
    const int npad = (n + 31) & ~31;
    int nsplit, wpc;
    split_rows(n, nsplit, wpc);
    const size_t smem = (size_t)2 * npad * LDS * 2;
    if (smem > 200 * 1024) return AIMB_ERR_UNSUPPORTED;
    AIMB_SET_SMEM_ATTR(200 * 1024, attn_fwd_mma_kernel<2>);     // per device, not a stream operation: capture-safe
    AIMB_SET_SMEM_ATTR(75 * 1024, attn_fwd_mma_kernel<3>);
    dim3 grid(frames * heads, nsplit);
    // three CTAs per SM (80 registers, no spills) when K and V of a head fit three times: 59.4 -> 55.3 us at n = 197;
    // at n = 257 only two fit and the 122-register build is the faster one (114.7 vs 124.9 us)
    if (smem <= 75 * 1024)
        launch_k((attn_fwd_mma_kernel<3>), dim3(grid), dim3(wpc * 32), smem, s, (const bf16*)qkv, (bf16*)o, lse, n, heads, wpc);
    else
        launch_k((attn_fwd_mma_kernel<2>), dim3(grid), dim3(wpc * 32), smem, s, (const bf16*)qkv, (bf16*)o, lse, n, heads, wpc);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

template <int NW>
static int bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames, int n,
                      int heads, size_t smem, cudaStream_t s) {
    AIMB_SET_SMEM_ATTR(227 * 1024, attn_bwd_mma_kernel<NW>);
    const int tiles = (n + 15) / 16;
    launch_k((attn_bwd_mma_kernel<NW>), dim3(frames * heads), dim3((tiles < NW ? tiles : NW) * 32), smem, s, (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o,
                                                                            lse, (bf16*)d_qkv, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

// kernels are compiled for the warp counts of the supported token counts (n <= 128: toy sizes, 197 -> 13 warps:
// ViT-B/16, 257 -> 9 warps x 2 passes: ViT-L/14) so the launch bounds fit the registers.
int attn_spatial_bwd_mma(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames, int n,
                         int heads, cudaStream_t s) {
    int npad = (n + 31) & ~31;
    int nwarps = (n + 15) / 16;
    size_t smem = (size_t)4 * npad * LDS * 2 + (size_t)2 * npad * 4;
    if (smem > 227 * 1024) return AIMB_ERR_UNSUPPORTED;
    if (nwarps <= 8) return bwd_launch<8>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);
    static const int dbg_nw = getenv("AIMB200_ATTN_BWD_NW") ? atoi(getenv("AIMB200_ATTN_BWD_NW")) : 0;   // bench_tools only
    if (dbg_nw == 7) return bwd_launch<7>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);
    if (dbg_nw == 9) return bwd_launch<9>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);
    // 9..13 row tiles (n = 197): 7 warps x 2 passes with the full 255-register budget measured 168 us vs 172 us for 13
    // warps capped at 128 registers (the kernel's time is set by ldmatrix + HMMA + issue work, not by occupancy)
    if (dbg_nw == 13) return bwd_launch<13>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);
    if (nwarps <= 13) return bwd_launch<7>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);
    return bwd_launch<9>(qkv, o, d_o, lse, d_qkv, frames, n, heads, smem, s);   // two passes (three beyond n = 288)
}


// ------------------------------------------------------------------------------------------ temporal attention, T = 8 / 16
// (vitclip_aim.py:196-200: attention over the T frames of one (clip, token), per head.)  The 16 rows of an m16n8k16
// tile are either the 8 frames of a PAIR of heads (T = 8: Q K^T is computed for the full 16 x 16 tile, only the two
// diagonal 8 x 8 blocks are used, P / dS re-enter the tensor core as block-diagonal A operands) or the 16 frames of one
// head (T = 16).  C-fragment == A-fragment layout, so P and dS need no shuffles; the transposes for dK / dV are one
// movmatrix per 8 x 8 block.  Rows are fetched in place (stride n rows) with 16-byte cp.async and results leave through
// a shared-memory tile as 16-byte stores.  ~10x fewer instructions than the SIMT kernel (which ncu showed issue-bound
// at 837 warp instructions per problem).
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
constexpr int TT_TILE = 16 * LDS;   // bf16 elements of one [16 rows][64 + pad] tile

// tile row r  <->  PAIR: frame r & 7 of head h0 + (r >> 3);  else: frame r of head h0
template <bool PAIR>
__device__ __forceinline__ int64_t tt_goff(int r, int64_t frame_stride) {
    return PAIR ? (int64_t)(r & 7) * frame_stride + (r >> 3) * HD : (int64_t)r * frame_stride;
}
template <bool PAIR>
__device__ __forceinline__ void tt_load(bf16* dst, const bf16* src, int64_t frame_stride, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = i * 32 + lane, r = e >> 3, c = (e & 7) * 8;
        cp_async16(dst + r * LDS + c, src + tt_goff<PAIR>(r, frame_stride) + c);
    }
}
// accumulators [16 rows x 64] -> smem tile -> global, 16 bytes per lane
template <bool PAIR>
__device__ __forceinline__ void tt_store(bf16* tile, bf16* dst, int64_t frame_stride, const float (&acc)[8][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        *reinterpret_cast<uint32_t*>(tile + g * LDS + dt * 8 + 2 * t) = pack_bf16(acc[dt][0], acc[dt][1]);
        *reinterpret_cast<uint32_t*>(tile + (g + 8) * LDS + dt * 8 + 2 * t) = pack_bf16(acc[dt][2], acc[dt][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e = i * 32 + lane, r = e >> 3, c = (e & 7) * 8;
        *reinterpret_cast<uint4*>(dst + tt_goff<PAIR>(r, frame_stride) + c) = *reinterpret_cast<const uint4*>(tile + r * LDS + c);
    }
}
// softmax of one row held as 4 values per lane across a quad (-inf entries = masked)
__device__ __forceinline__ void tt_softmax(const float (&x)[4], float (&p)[4]) {
    float m = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { p[j] = ex2_ftz((x[j] - m) * SCALE_LOG2); sum += p[j]; }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = __fdividef(1.f, sum);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] *= inv;
}
// rows g (pa) and g + 8 (pb) of the 16 x 16 score tile -> probabilities; PAIR masks the off-diagonal blocks
template <bool PAIR>
__device__ __forceinline__ void tt_probs(const float (&s)[2][4], float (&pa)[4], float (&pb)[4]) {
    const float xa[4] = {s[0][0], s[0][1], PAIR ? -INFINITY : s[1][0], PAIR ? -INFINITY : s[1][1]};
    const float xb[4] = {PAIR ? -INFINITY : s[0][2], PAIR ? -INFINITY : s[0][3], s[1][2], s[1][3]};
    tt_softmax(xa, pa);
    tt_softmax(xb, pb);
}
// values of rows g / g + 8 -> A fragment of the 16 x 16 matrix, and of its transpose
__device__ __forceinline__ void tt_afrag(const float (&ra)[4], const float (&rb)[4], uint32_t (&a)[4]) {
    a[0] = pack_bf16(ra[0], ra[1]);   // (row g,     k 0-7)
    a[1] = pack_bf16(rb[0], rb[1]);   // (row g + 8, k 0-7)
    a[2] = pack_bf16(ra[2], ra[3]);   // (row g,     k 8-15)
    a[3] = pack_bf16(rb[2], rb[3]);   // (row g + 8, k 8-15)
}
__device__ __forceinline__ void tt_afrag_t(const uint32_t (&a)[4], uint32_t (&at)[4]) {
    at[0] = movmatrix_trans(a[0]);
    at[1] = movmatrix_trans(a[2]);
    at[2] = movmatrix_trans(a[1]);
    at[3] = movmatrix_trans(a[3]);
}
// S tile (2 n-tiles of 8 keys) = A[16 x 64] . B[16 keys x 64]^T
__device__ __forceinline__ void tt_scores(float (&s)[2][4], const uint32_t (&a)[4][4], const bf16* sB, int lane) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t b[4];
            ldb_frag_nk(b, sB + (nt * 8) * LDS + kp * 32, lane);
            mma16816(s[nt], a[2 * kp], b[0], b[1]);
            mma16816(s[nt], a[2 * kp + 1], b[2], b[3]);
        }
    }
}
// acc[16 x 64] = A[16 x 16] . B[16 rows x 64]
__device__ __forceinline__ void tt_apply(float (&acc)[8][4], const uint32_t (&a)[4], const bf16* sB, int lane) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
        uint32_t b[4];
        ldb_frag_kn(b, sB + dp * 16, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[2 * dp][j] = 0.f; acc[2 * dp + 1][j] = 0.f; }
        mma16816(acc[2 * dp], a, b[0], b[1]);
        mma16816(acc[2 * dp + 1], a, b[2], b[3]);
    }
}
// which (clip, token, head) does this warp own
template <bool PAIR>
__device__ __forceinline__ bool tt_problem(int warp, int B, int n, int heads, int& h0, int64_t& row0) {
    constexpr int T_ = PAIR ? 8 : 16;
    const int hp = PAIR ? heads >> 1 : heads;
    const int64_t idx = (int64_t)blockIdx.x * 4 + warp;
    if (idx >= (int64_t)B * n * hp) return false;
    h0 = (PAIR ? 2 : 1) * (int)(idx % hp);
    const int tok = (int)((idx / hp) % n);
    const int b = (int)(idx / ((int64_t)hp * n));
    row0 = (int64_t)b * T_ * n + tok;
    return true;
}

template <bool PAIR>
__global__ void __launch_bounds__(128) attn_temporal_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int B,
                                                                    int n, int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int h0;
    int64_t row0;
    if (!tt_problem<PAIR>(warp, B, n, heads, h0, row0)) return;
    const int D = heads * HD, ld = 3 * D;
    bf16* sQ = reinterpret_cast<bf16*>(smraw) + warp * 3 * TT_TILE;
    bf16* sK = sQ + TT_TILE;
    bf16* sV = sK + TT_TILE;
    const bf16* base = qkv + row0 * ld + h0 * HD;
    const int64_t fs = (int64_t)n * ld;
    tt_load<PAIR>(sQ, base, fs, lane);
    tt_load<PAIR>(sK, base + D, fs, lane);
    tt_load<PAIR>(sV, base + 2 * D, fs, lane);
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lda_frag(qa[ks], sQ + ks * 16, lane);
    float s[2][4];
    tt_scores(s, qa, sK, lane);
    float pa[4], pb[4];
    tt_probs<PAIR>(s, pa, pb);
    uint32_t pf[4];
    tt_afrag(pa, pb, pf);
    float oacc[8][4];
    tt_apply(oacc, pf, sV, lane);
    tt_store<PAIR>(sQ, o + row0 * D + h0 * HD, (int64_t)n * D, oacc, lane);
}

template <bool PAIR>
__global__ void __launch_bounds__(128) attn_temporal_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                                    bf16* __restrict__ d_qkv, int B, int n, int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int h0;
    int64_t row0;
    if (!tt_problem<PAIR>(warp, B, n, heads, h0, row0)) return;
    const int D = heads * HD, ld = 3 * D;
    bf16* sQ = reinterpret_cast<bf16*>(smraw) + warp * 4 * TT_TILE;
    bf16* sK = sQ + TT_TILE;
    bf16* sV = sK + TT_TILE;
    bf16* sG = sV + TT_TILE;
    const bf16* base = qkv + row0 * ld + h0 * HD;
    const int64_t fs = (int64_t)n * ld;
    tt_load<PAIR>(sQ, base, fs, lane);
    tt_load<PAIR>(sK, base + D, fs, lane);
    tt_load<PAIR>(sV, base + 2 * D, fs, lane);
    tt_load<PAIR>(sG, d_o + row0 * D + h0 * HD, (int64_t)n * D, lane);
    cp_async_wait_all();
    __syncwarp();
    uint32_t a[4][4];
    float s[2][4], dp[2][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lda_frag(a[ks], sQ + ks * 16, lane);
    tt_scores(s, a, sK, lane);                   // S  = Q K^T
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lda_frag(a[ks], sG + ks * 16, lane);
    tt_scores(dp, a, sV, lane);                  // dP = dO V^T
    float pa[4], pb[4];
    tt_probs<PAIR>(s, pa, pb);
    const float dpa[4] = {dp[0][0], dp[0][1], dp[1][0], dp[1][1]}, dpb[4] = {dp[0][2], dp[0][3], dp[1][2], dp[1][3]};
    float da = 0.f, db = 0.f;                    // delta = sum_j P_ij dP_ij (masked entries have P = 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) { da = fmaf(pa[j], dpa[j], da); db = fmaf(pb[j], dpb[j], db); }
    da += __shfl_xor_sync(0xffffffffu, da, 1);
    da += __shfl_xor_sync(0xffffffffu, da, 2);
    db += __shfl_xor_sync(0xffffffffu, db, 1);
    db += __shfl_xor_sync(0xffffffffu, db, 2);
    float dsa[4], dsb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        dsa[j] = pa[j] * (dpa[j] - da) * 0.125f;
        dsb[j] = pb[j] * (dpb[j] - db) * 0.125f;
    }
    uint32_t pf[4], dsf[4], tf[4];
    tt_afrag(pa, pb, pf);
    tt_afrag(dsa, dsb, dsf);
    bf16* dbase = d_qkv + row0 * ld + h0 * HD;
    float acc[8][4];
    tt_apply(acc, dsf, sK, lane);                // dQ = dS K        (sV is free after dP: staging tile)
    tt_store<PAIR>(sV, dbase, fs, acc, lane);
    tt_afrag_t(dsf, tf);
    tt_apply(acc, tf, sQ, lane);                 // dK = dS^T Q
    tt_store<PAIR>(sV, dbase + D, fs, acc, lane);
    tt_afrag_t(pf, tf);
    tt_apply(acc, tf, sG, lane);                 // dV = P^T dO
    tt_store<PAIR>(sV, dbase + 2 * D, fs, acc, lane);
}


// ------------------------------------------------------------------------------------------ temporal attention, T = 32
// One warp per (clip, token, head): the 32 frames are two 16-row m-tiles; S is 32 x 32 (4 key n-tiles), P / dS feed
// the second GEMMs as two k-steps of A fragments, and the transposed fragments for dK / dV are 16 movmatrix each.
constexpr int T32_TILE = 32 * LDS;
__device__ __forceinline__ void t32_load(bf16* dst, const bf16* src, int64_t frame_stride, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = i * 32 + lane, r = e >> 3, c = (e & 7) * 8;
        cp_async16(dst + r * LDS + c, src + (int64_t)r * frame_stride + c);
    }
}
__device__ __forceinline__ void t32_store(bf16* tile, bf16* dst, int64_t frame_stride, const float (&acc)[2][8][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            *reinterpret_cast<uint32_t*>(tile + (mt * 16 + g) * LDS + dt * 8 + 2 * t) = pack_bf16(acc[mt][dt][0], acc[mt][dt][1]);
            *reinterpret_cast<uint32_t*>(tile + (mt * 16 + g + 8) * LDS + dt * 8 + 2 * t) = pack_bf16(acc[mt][dt][2], acc[mt][dt][3]);
        }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = i * 32 + lane, r = e >> 3, c = (e & 7) * 8;
        *reinterpret_cast<uint4*>(dst + (int64_t)r * frame_stride + c) = *reinterpret_cast<const uint4*>(tile + r * LDS + c);
    }
}
// s[mt][nt][4] = A[32 x 64] . B[32 rows x 64]^T
__device__ __forceinline__ void t32_scores(float (&s)[2][4][4], const bf16* sA, const bf16* sB, int lane) {
    uint32_t a[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) lda_frag(a[mt][ks], sA + mt * 16 * LDS + ks * 16, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[mt][nt][j] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t b[4];
            ldb_frag_nk(b, sB + (nt * 8) * LDS + kp * 32, lane);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma16816(s[mt][nt], a[mt][2 * kp], b[0], b[1]);
                mma16816(s[mt][nt], a[mt][2 * kp + 1], b[2], b[3]);
            }
        }
    }
}
// acc[mt][dt] = sum over the two k-steps of A[mt][kk] . B[32 rows x 64]
__device__ __forceinline__ void t32_apply(float (&acc)[2][8][4], const uint32_t (&a)[2][2][4], const bf16* sB, int lane) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dt = 0; dt < 8; ++dt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][dt][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
            uint32_t b[4];
            ldb_frag_kn(b, sB + kk * 16 * LDS + dp * 16, lane);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma16816(acc[mt][2 * dp], a[mt][kk], b[0], b[1]);
                mma16816(acc[mt][2 * dp + 1], a[mt][kk], b[2], b[3]);
            }
        }
}
// row softmax over 32 keys: row (mt, g) holds s[mt][nt][0..1], row (mt, g + 8) holds s[mt][nt][2..3]; in place -> P
__device__ __forceinline__ void t32_softmax(float (&s)[2][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float m = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) m = fmaxf(m, fmaxf(s[mt][nt][2 * h], s[mt][nt][2 * h + 1]));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float e = ex2_ftz((s[mt][nt][2 * h + j] - m) * SCALE_LOG2);
                    s[mt][nt][2 * h + j] = e;
                    sum += e;
                }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = __fdividef(1.f, sum);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) { s[mt][nt][2 * h] *= inv; s[mt][nt][2 * h + 1] *= inv; }
        }
}
// A fragments (row tile mt, k-step kk over keys) of a 32 x 32 matrix held in score layout
__device__ __forceinline__ void t32_afrag(const float (&x)[2][4][4], uint32_t (&a)[2][2][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            a[mt][kk][0] = pack_bf16(x[mt][2 * kk][0], x[mt][2 * kk][1]);
            a[mt][kk][1] = pack_bf16(x[mt][2 * kk][2], x[mt][2 * kk][3]);
            a[mt][kk][2] = pack_bf16(x[mt][2 * kk + 1][0], x[mt][2 * kk + 1][1]);
            a[mt][kk][3] = pack_bf16(x[mt][2 * kk + 1][2], x[mt][2 * kk + 1][3]);
        }
}
// fragments of the transposed matrix: 8 x 8 block B[rb][cb] of the source sits in a[rb >> 1][cb >> 1][(rb & 1) + 2 * (cb & 1)]
__device__ __forceinline__ void t32_afrag_t(const uint32_t (&a)[2][2][4], uint32_t (&at)[2][2][4]) {
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)          // row tile of the transpose = column (key) tile of the source
#pragma unroll
        for (int qs = 0; qs < 2; ++qs) {    // k-step of the transpose = row (query) tile of the source
            at[kt][qs][0] = movmatrix_trans(a[qs][kt][0]);   // B[2qs][2kt]
            at[kt][qs][1] = movmatrix_trans(a[qs][kt][2]);   // B[2qs][2kt+1]
            at[kt][qs][2] = movmatrix_trans(a[qs][kt][1]);   // B[2qs+1][2kt]
            at[kt][qs][3] = movmatrix_trans(a[qs][kt][3]);   // B[2qs+1][2kt+1]
        }
}
__device__ __forceinline__ bool t32_problem(int warp, int wpb, int B, int n, int heads, int& h0, int64_t& row0) {
    const int64_t idx = (int64_t)blockIdx.x * wpb + warp;
    if (idx >= (int64_t)B * n * heads) return false;
    h0 = (int)(idx % heads);
    const int tok = (int)((idx / heads) % n);
    const int b = (int)(idx / ((int64_t)heads * n));
    row0 = (int64_t)b * 32 * n + tok;
    return true;
}

__global__ void __launch_bounds__(64) attn_temporal32_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int B, int n,
                                                                     int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int h0;
    int64_t row0;
    if (!t32_problem(warp, 2, B, n, heads, h0, row0)) return;
    const int D = heads * HD, ld = 3 * D;
    bf16* sQ = reinterpret_cast<bf16*>(smraw) + warp * 3 * T32_TILE;
    bf16* sK = sQ + T32_TILE;
    bf16* sV = sK + T32_TILE;
    const bf16* base = qkv + row0 * ld + h0 * HD;
    const int64_t fs = (int64_t)n * ld;
    t32_load(sQ, base, fs, lane);
    t32_load(sK, base + D, fs, lane);
    t32_load(sV, base + 2 * D, fs, lane);
    cp_async_wait_all();
    __syncwarp();
    float s[2][4][4];
    t32_scores(s, sQ, sK, lane);
    t32_softmax(s);
    uint32_t pf[2][2][4];
    t32_afrag(s, pf);
    float acc[2][8][4];
    t32_apply(acc, pf, sV, lane);
    t32_store(sQ, o + row0 * D + h0 * HD, (int64_t)n * D, acc, lane);
}

__global__ void __launch_bounds__(64) attn_temporal32_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                                     bf16* __restrict__ d_qkv, int B, int n, int heads) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int h0;
    int64_t row0;
    if (!t32_problem(warp, 2, B, n, heads, h0, row0)) return;
    const int D = heads * HD, ld = 3 * D;
    bf16* sQ = reinterpret_cast<bf16*>(smraw) + warp * 4 * T32_TILE;
    bf16* sK = sQ + T32_TILE;
    bf16* sV = sK + T32_TILE;
    bf16* sG = sV + T32_TILE;
    const bf16* base = qkv + row0 * ld + h0 * HD;
    const int64_t fs = (int64_t)n * ld;
    t32_load(sQ, base, fs, lane);
    t32_load(sK, base + D, fs, lane);
    t32_load(sV, base + 2 * D, fs, lane);
    t32_load(sG, d_o + row0 * D + h0 * HD, (int64_t)n * D, lane);
    cp_async_wait_all();
    __syncwarp();
    float p[2][4][4], dp[2][4][4];
    t32_scores(p, sQ, sK, lane);                 // S = Q K^T -> P
    t32_softmax(p);
    t32_scores(dp, sG, sV, lane);                // dP = dO V^T
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {            // dS = P (dP - rowsum(P dP)) / 8, in place over dp
            float dl = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dl += p[mt][nt][2 * h] * dp[mt][nt][2 * h] + p[mt][nt][2 * h + 1] * dp[mt][nt][2 * h + 1];
            dl += __shfl_xor_sync(0xffffffffu, dl, 1);
            dl += __shfl_xor_sync(0xffffffffu, dl, 2);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) dp[mt][nt][2 * h + j] = p[mt][nt][2 * h + j] * (dp[mt][nt][2 * h + j] - dl) * 0.125f;
        }
    uint32_t pf[2][2][4], dsf[2][2][4], tf[2][2][4];
    t32_afrag(p, pf);
    t32_afrag(dp, dsf);
    bf16* dbase = d_qkv + row0 * ld + h0 * HD;
    float acc[2][8][4];
    t32_apply(acc, dsf, sK, lane);               // dQ = dS K        (sV is free after dP: staging tile)
    t32_store(sV, dbase, fs, acc, lane);
    t32_afrag_t(dsf, tf);
    t32_apply(acc, tf, sQ, lane);                // dK = dS^T Q
    t32_store(sV, dbase + D, fs, acc, lane);
    t32_afrag_t(pf, tf);
    t32_apply(acc, tf, sG, lane);                // dV = P^T dO
    t32_store(sV, dbase + 2 * D, fs, acc, lane);
}

// T = 8 (heads paired, heads must be even), T = 16 or T = 32
int attn_temporal_fwd_mma(const void* qkv, void* o, int B, int T, int n, int heads, cudaStream_t s) {
    if (T == 32) {
        const int64_t probs32 = (int64_t)B * n * heads;
        launch_k(attn_temporal32_fwd_mma_kernel, dim3((unsigned)((probs32 + 1) / 2)), dim3(64), (size_t)2 * 3 * T32_TILE * 2, s,
                 (const bf16*)qkv, (bf16*)o, B, n, heads);
        AIMB_CHECK_LAUNCH();
        return AIMB_OK;
    }
    const int64_t probs = (int64_t)B * n * (T == 8 ? heads / 2 : heads);
    const size_t smem = (size_t)4 * 3 * TT_TILE * 2;
    const dim3 grid((unsigned)((probs + 3) / 4));
    if (T == 8) launch_k(attn_temporal_fwd_mma_kernel<true>, grid, dim3(128), smem, s, (const bf16*)qkv, (bf16*)o, B, n, heads);
    else launch_k(attn_temporal_fwd_mma_kernel<false>, grid, dim3(128), smem, s, (const bf16*)qkv, (bf16*)o, B, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}
int attn_temporal_bwd_mma(const void* qkv, const void* d_o, void* d_qkv, int B, int T, int n, int heads, cudaStream_t s) {
    if (T == 32) {
        const int64_t probs32 = (int64_t)B * n * heads;
        launch_k(attn_temporal32_bwd_mma_kernel, dim3((unsigned)((probs32 + 1) / 2)), dim3(64), (size_t)2 * 4 * T32_TILE * 2, s,
                 (const bf16*)qkv, (const bf16*)d_o, (bf16*)d_qkv, B, n, heads);
        AIMB_CHECK_LAUNCH();
        return AIMB_OK;
    }
    const int64_t probs = (int64_t)B * n * (T == 8 ? heads / 2 : heads);
    const size_t smem = (size_t)4 * 4 * TT_TILE * 2;
    const dim3 grid((unsigned)((probs + 3) / 4));
    if (T == 8)
        launch_k(attn_temporal_bwd_mma_kernel<true>, grid, dim3(128), smem, s, (const bf16*)qkv, (const bf16*)d_o, (bf16*)d_qkv, B, n, heads);
    else
        launch_k(attn_temporal_bwd_mma_kernel<false>, grid, dim3(128), smem, s, (const bf16*)qkv, (const bf16*)d_o, (bf16*)d_qkv, B, n, heads);
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

}  // namespace aimb

// Fully strided SIMT GEMM with fp32 accumulation.  This is the fp32 parity-mode GEMM (1e-3 logit gate)
// and the cross-check for the tcgen05 kernel; it is not the bf16 hot path.
#include "common.cuh"
#include "epilogue.cuh"

namespace aimb {

constexpr int SBM = 128, SBN = 128, SBK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, int64_t a_sm, int64_t a_sk,
                                                        const T* __restrict__ B, int64_t b_sn, int64_t b_sk,
                                                        EpiParams epi, int64_t M, int N, int K) {
    pdl_grid_sync();
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * SBM;
    const int n0 = blockIdx.x * SBN;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each an 8x8 micro-tile (split 4+4 for bank spread)
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const bool a_kc = (a_sk == 1), b_kc = (b_sk == 1);
    for (int k0 = 0; k0 < K; k0 += SBK) {
        // ---- load A tile (128 x 16) and B tile (128 x 16): 8 elements per thread each
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int e = tid + i * 256;
            int mm, kk;
            if (a_kc) { mm = e >> 4; kk = e & 15; } else { mm = e & 127; kk = e >> 7; }
            int64_t gm = m0 + mm; int gk = k0 + kk;
            float v = 0.f;
            if (gm < M && gk < K) v = ldf<T>(A + gm * a_sm + gk * a_sk);
            As[kk][mm] = v;
            int nn;
            if (b_kc) { nn = e >> 4; kk = e & 15; } else { nn = e & 127; kk = e >> 7; }
            int gn = n0 + nn; gk = k0 + kk;
            v = 0.f;
            if (gn < N && gk < K) v = ldf<T>(B + gn * b_sn + gk * b_sk);
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < N) epilogue_store<T>(epi, m, n, acc[i][j]);
        }
    }
}

int gemm_simt_launch(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sn, int64_t b_sk,
                     const EpiParams& epi, int64_t M, int N, int K, int dtype, cudaStream_t s) {
    if (M == 0) return AIMB_OK;
    dim3 grid((N + SBN - 1) / SBN, (unsigned)((M + SBM - 1) / SBM));
    if (dtype == AIMB_BF16)
        launch_k((gemm_simt_kernel<bf16>), dim3(grid), dim3(256), 0, s, (const bf16*)A, a_sm, a_sk, (const bf16*)B, b_sn, b_sk, epi, M, N, K);
    else if (dtype == AIMB_F32)
        launch_k((gemm_simt_kernel<float>), dim3(grid), dim3(256), 0, s, (const float*)A, a_sm, a_sk, (const float*)B, b_sn, b_sk, epi, M, N,
                                                     K);
    else return AIMB_ERR_ARG;
    AIMB_CHECK_LAUNCH();
    return AIMB_OK;
}

}  // namespace aimb

using namespace aimb;

static bool epi_ok(const aimb_epilogue_t* e) {
    if (!e || !e->out) return false;
    if (e->accumulate && !e->out_f32) return false;
    if (e->bias_rowscaled && !e->row_scale) return false;
    return true;
}

extern "C" int aimb_gemm_strided(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sn, int64_t b_sk,
                                 const aimb_epilogue_t* epi, int64_t M, int32_t N, int32_t K, int32_t dtype,
                                 void* stream) {
    if (!A || !B || !epi_ok(epi) || M < 0 || N <= 0 || K <= 0) return AIMB_ERR_ARG;
    EpiParams p = make_epi(epi, N);
    if (p.colsum_out && !p.colsum_accumulate && cudaMemsetAsync(p.colsum_out, 0, (size_t)N * 4, (cudaStream_t)stream) != cudaSuccess) return AIMB_ERR_CUDA;
    return gemm_simt_launch(A, a_sm, a_sk, B, b_sn, b_sk, p, M, N, K, dtype, (cudaStream_t)stream);
}

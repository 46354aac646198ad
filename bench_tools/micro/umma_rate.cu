// tcgen05.mma issue cost per instruction (K = 16) for the shapes the attention kernels use: SS vs TS (A in TMEM),
// K-major vs MN-major B, N = 16..256.  One CTA per SM, one thread issues `reps` MMAs back to back, then commits and
// waits; clock64 around it.  Operands are whatever is in shared memory / TMEM (values do not matter for timing).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../adapt-image-models_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace aimb;

// warp-uniform issue: every lane of the (converged) warp executes this, one elected lane issues the MMA
__device__ __forceinline__ void umma_ss_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_elect(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int mode>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int reps, int tbase, int nacc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 40960; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tptr);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (tptr != 0) __trap();          // all 512 columns allocated: the base is column 0
    const uint32_t tb = (uint32_t)tbase;      // == 0, but a kernel parameter keeps the address arithmetic in the uniform datapath
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    if (warp_u == 0) {
        const uint32_t sa = ptx::smem_u32(smem), sb = sa + 65536;
        const int b_mn = (mode & 2) ? 1 : 0;
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N, 0, b_mn);
        const uint64_t bd0 = b_mn ? ptx::umma_desc_mnmajor_sw128(sb, 8192) : ptx::umma_desc_kmajor_sw128(sb);
        const uint64_t bstep = b_mn ? 128 : 2;     // per k-step: 2048 B (MN-major) / 32 B (K-major), in 16-byte units
        const uint64_t ad0 = ptx::umma_desc_kmajor_sw128(sa);
        const uint32_t dstep = nacc > 1 ? 64 : 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const uint32_t dcol = tb + 256 + (j & (nacc - 1)) * dstep;
                if (mode & 1) umma_ts_elect(dcol, tb + (j & 7) * 8, bd0 + (j & 3) * bstep, idesc, 1u);
                else umma_ss_elect(dcol, ad0 + 2 * (j & 3), bd0 + (j & 3) * bstep, idesc, 1u);
            }
        }
        if (ptx::elect_one()) ptx::umma_commit(&bar);
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tb);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d; cudaMalloc(&d, sms * 8);
    const char* names[4] = {"SS  B K-major ", "TS  B K-major ", "SS  B MN-major", "TS  B MN-major"};
    for (int mode = 0; mode < 4; ++mode)
        for (int N : {16, 32, 64, 80, 96, 128, 208, 256}) {
            if ((mode & 2) && N > 128) continue;     // MN-major probe keeps to two 64-column groups
            const int reps = 64, ks = 12;
            printf("%s M=128 N=%3d [floor %3.0f]:", names[mode], N, 128.0 * N / 256);
            for (int nacc : {1, 2, 4}) {
                if (nacc > 1 && 256 + (nacc - 1) * 64 + N > 512) continue;
                auto kf = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : k<3>;
                cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                kf<<<sms, 128, 180 * 1024>>>(d, N, 2, 0, nacc);
                kf<<<sms, 128, 180 * 1024>>>(d, N, reps, 0, nacc);
                long long h[256]; cudaMemcpy(h, d, sms * 8, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("  %d acc: %.1f clk/MMA", nacc, (double)h[0] / (reps * ks));
            }
            printf("\n");
        }
    return 0;
}

// Legacy tensor path ceiling on sm_100a: mma.sync.m16n8k16 bf16 issue rate with operands in registers (no smem).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_peak hmma_peak.cu && ./hmma_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(float* out, int iters) {
    float acc[CHAINS][4];
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma16816(acc[c], a, (uint32_t)i, (uint32_t)c);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) for (int j = 0; j < 4; ++j) s += acc[c][j];
    if (s == 123.456f) out[0] = s;
}
template <int CHAINS> void run(int warps_per_sm, int sms) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CHAINS><<<sms, warps_per_sm * 32>>>(d, 100);
    cudaEventRecord(e0);
    k<CHAINS><<<sms, warps_per_sm * 32>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop = (double)sms * warps_per_sm * iters * CHAINS * 4096.0;
    printf("chains %d warps/SM %2d: %.1f TFLOP/s\n", CHAINS, warps_per_sm, flop / ms / 1e9);
    cudaFree(d);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) run<8>(w, sms);
    for (int w : {4, 8, 16}) run<2>(w, sms);
    return 0;
}

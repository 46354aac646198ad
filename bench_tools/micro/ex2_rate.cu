// MUFU.EX2 issue rate on sm_100a: fp32 ex2.approx vs the packed bf16x2 / f16x2 forms, and an FMA-pipe polynomial exp2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2_rate ex2_rate.cu && ./ex2_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    constexpr int CH = 16;
    float f[CH];
    uint32_t h[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { f[c] = -0.001f * (threadIdx.x + c); h[c] = 0xBC00BC00u + threadIdx.x + c; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
            if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[c]));
            if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[c]));
            if (MODE == 3) {   // exp2 on the FMA / ALU pipes: x = i + r, 2^r by a cubic, exponent added as an integer
                float x = f[c];
                float fl = floorf(x);
                float r = x - fl;
                float p = fmaf(fmaf(fmaf(0.0555041f, r, 0.2402265f), r, 0.6931472f), r, 1.0f);
                f[c] = __int_as_float(__float_as_int(p) + ((int)fl << 23)) - 1.5f;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += f[c] + __uint_as_float(h[c]);
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int warps, int sms, int per_inst) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps * 32>>>(d, 10);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)sms * warps * 32 * iters * 16 * per_inst;
    printf("%-28s warps/SM %2d: %7.2f G results/s/SM  (%.2f results/clk/SM at 1.9 GHz)\n", name, warps, n / ms / 1e6 / sms, n / ms / 1e6 / sms / 1.9);
    cudaFree(d);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16}) {
        run<0>("ex2.approx.ftz.f32", w, sms, 1);
        run<1>("ex2.approx.ftz.bf16x2", w, sms, 2);
        run<2>("ex2.approx.f16x2", w, sms, 2);
        run<3>("cubic exp2 (fma/alu pipes)", w, sms, 1);
    }
    return 0;
}

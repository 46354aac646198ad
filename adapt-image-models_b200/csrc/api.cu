// Library-level entry points and the dispatch between SIMT and tensor-core attention kernels.
#include "common.cuh"

namespace aimb {
int attn_spatial_fwd_simt_dispatch(const void* qkv, void* o, float* lse, int frames, int n, int heads, int dtype,
                                   cudaStream_t s);
int attn_spatial_bwd_simt_dispatch(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv,
                                   int frames, int n, int heads, int dtype, cudaStream_t s);
int attn_spatial_fwd_mma(const void* qkv, void* o, float* lse, int frames, int n, int heads, cudaStream_t s);
int attn_spatial_bwd_mma(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames, int n,
                         int heads, cudaStream_t s);
// attn_tc.cu: tcgen05 / TMEM / TMA kernels (n <= 256)
bool attn_spatial_tc_supported(int n, int heads);
int attn_spatial_fwd_tc(const void* qkv, void* o, float* lse, int frames, int n, int heads, cudaStream_t s);
void attn_tc_set_timeline(long long* p);
int attn_spatial_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv, int frames, int n,
                        int heads, cudaStream_t s);
}  // namespace aimb

namespace aimb { int g_pdl_enabled = 1; }

using namespace aimb;

extern "C" void aimb_debug_set_pdl(int on) { g_pdl_enabled = on; }
static int g_attn_mode = 0;   // 0: tcgen05 kernels where they apply, 1: mma.sync kernels only (cross-check / A-B timing)
extern "C" void aimb_debug_attn_mode(int mode) { g_attn_mode = mode; }
// bench_tools only: device buffer (>= 16 int64 per unit of CTA 0) that receives clock64 stamps of the pipeline events
extern "C" void aimb_debug_attn_timeline(void* dev_buf) { attn_tc_set_timeline((long long*)dev_buf); }

extern "C" int aimb_version(void) { return 100; }

extern "C" const char* aimb_last_error(void) { return cudaGetErrorString(cudaPeekAtLastError()); }

extern "C" int aimb_device_ok(int device) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return AIMB_ERR_CUDA;
    return (p.major == 10) ? AIMB_OK : AIMB_ERR_UNSUPPORTED;
}

extern "C" int aimb_attn_spatial_fwd(const void* qkv, void* o, float* lse, int32_t frames, int32_t n, int32_t heads,
                                     int32_t dtype, int32_t impl, void* stream) {
    if (!qkv || !o || frames < 0 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (frames == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && g_attn_mode == 0 && attn_spatial_tc_supported(n, heads))
        return attn_spatial_fwd_tc(qkv, o, lse, frames, n, heads, s);
    if (dtype == AIMB_BF16 && (impl == AIMB_IMPL_AUTO || impl == AIMB_IMPL_MMA)) return attn_spatial_fwd_mma(qkv, o, lse, frames, n, heads, s);
    return attn_spatial_fwd_simt_dispatch(qkv, o, lse, frames, n, heads, dtype, s);
}

extern "C" int aimb_attn_spatial_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* d_qkv,
                                     int32_t frames, int32_t n, int32_t heads, int32_t dtype, int32_t impl, void* stream) {
    if (!qkv || !o || !d_o || !lse || !d_qkv || frames < 0 || n <= 0 || heads <= 0) return AIMB_ERR_ARG;
    if (frames == 0) return AIMB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == AIMB_BF16 && impl == AIMB_IMPL_AUTO && g_attn_mode == 0 && attn_spatial_tc_supported(n, heads))
        return attn_spatial_bwd_tc(qkv, o, d_o, lse, d_qkv, frames, n, heads, s);
    if (dtype == AIMB_BF16 && (impl == AIMB_IMPL_AUTO || impl == AIMB_IMPL_MMA))
        return attn_spatial_bwd_mma(qkv, o, d_o, lse, d_qkv, frames, n, heads, s);
    return attn_spatial_bwd_simt_dispatch(qkv, o, d_o, lse, d_qkv, frames, n, heads, dtype, s);
}

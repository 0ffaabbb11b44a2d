// peaks.cu — measured on-chip bandwidth peaks of the device this library runs on: the denominators of the online step's
// roofline when the planes never leave the chip ("L2 or distributed-SMEM bandwidth when the planes fit
// on chip").  Nothing here is on a product path; vsom_debug_measure_peaks is called by bench.py, which prints the numbers
// next to the kernel's achieved figure.
//   * shared memory: every SM reads a ~200 KB shared buffer over and over with conflict-free 128-bit loads (LDS.128), one
//     CTA of 1024 threads per SM — the access pattern of the resident rows in K1 / K1F;
//   * L2: all SMs read one buffer that fits L2 (32 MiB by default) over and over with 128-bit L2-only loads (ld.global.cg),
//     every SM walking the whole buffer from a different offset, so each byte is served by L2, not by L1 or HBM.
#include "common.cuh"

#include <algorithm>

namespace vsom
{

__global__ void __launch_bounds__(1024, 1) smem_read_kernel(int words4, int iters, float *sink)
{
    extern __shared__ __align__(16) float4 buf[];
    for (int i = threadIdx.x; i < words4; i += blockDim.x)
        buf[i] = make_float4(1.0f, 2.0f, 3.0f, 4.0f);
    __syncthreads();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll 8
        for (int i = threadIdx.x; i < words4; i += 1024)
        {
            const float4 v = buf[i];
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) // never true: keeps the loads alive
        sink[blockIdx.x] = acc.x;
}

__global__ void __launch_bounds__(1024, 1) l2_read_kernel(const float4 *__restrict__ src, size_t words4, int iters, float *sink)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // every CTA starts at its own offset and wraps: at any moment the SMs pull different lines
    const size_t start = (words4 / gridDim.x) * blockIdx.x;
    for (int it = 0; it < iters; ++it)
    {
        size_t i = start + threadIdx.x;
#pragma unroll 4
        for (size_t k = 0; k < words4; k += 1024)
        {
            if (i >= words4)
                i -= words4;
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
            i += 1024;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f)
        sink[blockIdx.x] = acc.x;
}

} // namespace vsom

using namespace vsom;

extern "C" int vsom_debug_measure_peaks(vsom_ctx *ctx, double out[4])
{
    if (!ctx || !out)
        return VSOM_ERR_INVALID;
    VSOM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int sms = ctx->numSMs;
    float *sink = nullptr;
    float4 *buf = nullptr;
    const size_t l2Bytes = size_t{32} << 20, words4 = l2Bytes / 16;
    VSOM_CUDA(ctx, cudaMalloc(&sink, sizeof(float) * sms));
    VSOM_CUDA(ctx, cudaMalloc(&buf, l2Bytes));
    VSOM_CUDA(ctx, cudaMemsetAsync(buf, 0, l2Bytes, ctx->stream));
    cudaEvent_t e0, e1;
    VSOM_CUDA(ctx, cudaEventCreate(&e0));
    VSOM_CUDA(ctx, cudaEventCreate(&e1));
    auto timed = [&](auto launch, float &ms) -> int {
        launch(); // warm-up
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep)
        {
            VSOM_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
            launch();
            VSOM_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
            VSOM_CUDA(ctx, cudaEventSynchronize(e1));
            float t = 0.f;
            VSOM_CUDA(ctx, cudaEventElapsedTime(&t, e0, e1));
            best = std::min(best, t);
        }
        ms = best;
        return VSOM_OK;
    };
    // ---- shared memory
    const int smemBytes = std::min(ctx->smemOptin - 1024, 200 * 1024) & ~16383, sWords4 = smemBytes / 16, sIters = 2000;
    VSOM_CUDA(ctx, cudaFuncSetAttribute(smem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    float ms = 0.f;
    int rc = timed([&] { smem_read_kernel<<<sms, 1024, smemBytes, ctx->stream>>>(sWords4, sIters, sink); }, ms);
    if (rc)
        return rc;
    out[0] = static_cast<double>(smemBytes) * sIters * sms / (ms * 1e-3) / 1e9; // GB/s, all SMs
    out[1] = static_cast<double>(smemBytes) * sIters / (ms * 1e-3) / 1e9;       // GB/s per SM
    // ---- L2
    const int lIters = 8;
    rc = timed([&] { l2_read_kernel<<<sms, 1024, 0, ctx->stream>>>(buf, words4, lIters, sink); }, ms);
    if (rc)
        return rc;
    out[2] = static_cast<double>(l2Bytes) * lIters * sms / (ms * 1e-3) / 1e9; // GB/s, all SMs reading a 32 MiB L2-resident set
    out[3] = static_cast<double>(l2Bytes) / (1 << 20);
    VSOM_CUDA(ctx, cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(buf);
    ctx->launches += 8;
    return VSOM_OK;
}

// Micro-benchmark (measurement tool, not product code): cycle cost of the building blocks of K1F's per-sample chain.
#include "../../variational-self-organizing-maps_b200/csrc/online_step_fast.cu"
#include <cstdio>
namespace vsom { int cuda_fail(vsom_ctx *, cudaError_t, const char *, const char *, int) { return -1; } int set_error(vsom_ctx *, int c, const std::string &) { return c; } }
using namespace vsom;

__global__ void __launch_bounds__(1024, 1) k_phase(long long *out, int stride, int n16, int L)
{
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 64 * stride; i += blockDim.x)
        sm[i] = 1.0f + (i % 7) * 0.25f;
    __syncthreads();
    long long t0, t1;
    float acc = 0.f;
    // (1) register-only dependent FADD chain, 128 long
    if (warp == 0)
    {
        float s = sm[lane];
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < 128; ++i)
            s = __fadd_rn(s, 1.25f);
        t1 = clock64();
        acc += s;
        if (tid == 0)
            out[0] = t1 - t0;
    }
    __syncthreads();
    // (2) chain_terms as the kernel calls it: lanes = nodes
    if (warp == 0)
    {
        t0 = clock64();
        float s = 0.f;
        if (lane < L)
            s = chain_terms(sm + lane * stride, n16);
        t1 = clock64();
        acc += s;
        if (tid == 0)
            out[1] = t1 - t0;
    }
    __syncthreads();
    // (3) chain + key + two REDUX
    if (warp == 0)
    {
        t0 = clock64();
        u64 key = ~0ull;
        if (lane < L)
            key = make_key_xy(chain_terms(sm + lane * stride, n16), lane, 3, 5);
        key = warp_min_key(key);
        t1 = clock64();
        acc += (float)(key & 0xffff);
        if (tid == 0)
            out[2] = t1 - t0;
    }
    __syncthreads();
    // (4) __syncthreads with 32 warps, back to back x16
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i)
        __syncthreads();
    t1 = clock64();
    if (tid == 0)
        out[3] = (t1 - t0) / 16;
    // (5) LDS dependent latency (pointer chase in smem)
    if (warp == 0)
    {
        int *ism = reinterpret_cast<int *>(sm);
        __syncwarp();
        if (lane == 0)
            for (int i = 0; i < 64; ++i)
                ism[i] = (i + 1) & 63;
        __syncwarp();
        int j = 0;
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            j = ism[j];
        t1 = clock64();
        acc += j;
        if (tid == 0)
            out[4] = (t1 - t0) / 32;
    }
    __syncthreads();
    // (6) chain while the other 31 warps sit at a barrier (as in the kernel)
    if (warp == 0)
    {
        t0 = clock64();
        float s = 0.f;
        if (lane < L)
            s = chain_terms(sm + lane * stride, n16);
        t1 = clock64();
        acc += s;
        if (tid == 0)
            out[5] = t1 - t0;
    }
    __syncthreads();
    // (7) one update item per warp for 28 warps (median), like the update phase: cycles from barrier to barrier
    {
        __syncthreads();
        t0 = clock64();
        if (warp < L)
        {
            float *mp = sm + warp * stride + lane * 4;
            float *sp = sm + (32 + warp) * stride + lane * 4;
            float4 mv = *reinterpret_cast<float4 *>(mp), sv = *reinterpret_cast<float4 *>(sp);
            const float4 xv = make_float4(1.5f, 0.5f, 2.5f, 1.0f), nv = make_float4(0.5f, 0.25f, 2.0f, 1.5f);
            const float c = 0.05f, nwf = 0.9f;
            fast_update_element<VSOM_MEDIAN>(xv.x, c, nwf, mv.x, sv.x);
            fast_update_element<VSOM_MEDIAN>(xv.y, c, nwf, mv.y, sv.y);
            fast_update_element<VSOM_MEDIAN>(xv.z, c, nwf, mv.z, sv.z);
            fast_update_element<VSOM_MEDIAN>(xv.w, c, nwf, mv.w, sv.w);
            *reinterpret_cast<float4 *>(mp) = mv;
            *reinterpret_cast<float4 *>(sp) = sv;
            *reinterpret_cast<float4 *>(sp + 4 * stride) = make_float4(sq_res(mv.x, nv.x), sq_res(mv.y, nv.y), sq_res(mv.z, nv.z), sq_res(mv.w, nv.w));
        }
        __syncthreads();
        t1 = clock64();
        if (tid == 0)
            out[6] = t1 - t0;
    }
    // (8) REDUX pair alone
    if (warp == 0)
    {
        u64 key = (u64)lane * 77 + 5;
        t0 = clock64();
        key = warp_min_key(key);
        t1 = clock64();
        acc += (float)(key & 0xff);
        if (tid == 0)
            out[7] = t1 - t0;
    }
    // (9) clock64 back to back
    t0 = clock64();
    t1 = clock64();
    if (tid == 0)
        out[8] = t1 - t0;
    if (acc == 12345.678f)
        out[15] = 1;
}

int main()
{
    long long *d;
    cudaMalloc(&d, 16 * 8);
    cudaMemset(d, 0, 16 * 8);
    const int stride = 132;
    cudaFuncSetAttribute(k_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep)
    {
        k_phase<<<1, 1024, 70 * stride * 4>>>(d, stride, 8, 28);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess)
        {
            printf("error %s\n", cudaGetErrorString(e));
            return 1;
        }
    }
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[] = {"128 dependent FADD (registers)", "chain_terms 28 lanes n16=8", "chain + key + 2 REDUX", "__syncthreads (1024 thr) each", "LDS dependent latency", "chain_terms again", "update item x28 warps barrier-to-barrier", "warp_min_key (2 REDUX)", "clock64 pair"};
    for (int i = 0; i < 9; ++i)
        printf("%-44s %lld cycles\n", names[i], h[i]);
    return 0;
}

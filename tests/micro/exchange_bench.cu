// Micro-benchmark (measurement tool, not product code): latency of one grid-wide min-loc exchange between the
// persistent CTAs of K1, in isolation, for several protocols.  Build: nvcc -arch=sm_100a -O3 -o exchange_bench exchange_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned long long u64;

__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_volatile(const u64 *p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 warp_min(u64 k)
{
    for (int o = 16; o; o >>= 1)
    {
        u64 other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
    }
    return k;
}

// variant 0: push all-to-all, poll own row (the protocol of online_step.cu)
// variant 1: same, polling with 10 lanes x 128-bit... (not used)
// variant 2: pull: one slot per CTA in one shared row, everyone polls the whole row
// variant 3: red.min on one word + counter
// variant 4: two-level push: groups of `gs` CTAs
__global__ void __launch_bounds__(1024, 1) bench(u64 *slots, u64 *out, long long *cyc, int iters, int variant, int gs, int work)
{
    const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Gp = (G + 15) & ~15;
    __shared__ u64 sRes;
    u64 acc = 0;
    long long t0 = clock64();
    for (int t = 0; t < iters; ++t)
    {
        // pseudo work so that keys differ per step
        const u64 tag = (u64)((t >> 1) & 0xff);
        u64 key = (((u64)((b * 2654435761u + t * 40503u) & 0xffffffu)) << 8) | tag;
        if (work > 0)
        {
            long long w0 = clock64();
            while (clock64() - w0 < work)
                ;
        }
        if (warp == 0)
        {
            u64 m = ~0ull;
            if (variant == 0)
            {
                u64 *buf = slots + (size_t)(t & 1) * G * Gp;
                for (int d = lane; d < G; d += 32)
                    st_relaxed(buf + (size_t)d * Gp + b, key);
                const u64 *row = buf + (size_t)b * Gp;
                const u64 filler = (~0ull << 8) | tag;
                for (;;)
                {
                    u64 v[5];
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                    {
                        int i = lane + 32 * j;
                        v[j] = i < G ? ld_relaxed(row + i) : filler;
                    }
                    int ok = 1;
                    m = ~0ull;
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                    {
                        ok &= ((v[j] & 0xff) == tag);
                        m = v[j] < m ? v[j] : m;
                    }
                    if (__all_sync(0xffffffffu, ok))
                        break;
                }
                m = warp_min(m);
            }
            else if (variant == 2)
            {
                u64 *buf = slots + (size_t)(t & 1) * Gp;
                if (lane == 0)
                    st_relaxed(buf + b, key);
                const u64 filler = (~0ull << 8) | tag;
                for (;;)
                {
                    u64 v[5];
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                    {
                        int i = lane + 32 * j;
                        v[j] = i < G ? ld_relaxed(buf + i) : filler;
                    }
                    int ok = 1;
                    m = ~0ull;
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                    {
                        ok &= ((v[j] & 0xff) == tag);
                        m = v[j] < m ? v[j] : m;
                    }
                    if (__all_sync(0xffffffffu, ok))
                        break;
                }
                m = warp_min(m);
            }
            else if (variant == 3)
            {
                // slots[0..3]: min words (ring of 4, reset two steps later), slots[16]: counter
                u64 *mn = slots + (t & 3) * 16;
                u64 *cnt = slots + 64;
                if (lane == 0)
                {
                    slots[((t + 2) & 3) * 16] = ~0ull; // benign: everyone writes the same reset value two steps ahead
                    atomicMin(mn, key);
                    __threadfence();
                    atomicAdd(cnt, 1ull);
                    const u64 want = (u64)(t + 1) * G;
                    while (ld_relaxed(cnt) < want)
                        ;
                    __threadfence();
                    m = ld_relaxed(mn);
                }
                m = __shfl_sync(0xffffffffu, m, 0);
            }
            else if (variant == 4)
            {
                // level 1: inside the group of gs consecutive CTAs; level 2: between CTAs of equal rank in every group
                const int ng = (G + gs - 1) / gs, g = b / gs, r = b % gs;
                const int gsize = (g == ng - 1) ? G - g * gs : gs;
                u64 *buf1 = slots + (size_t)(t & 1) * G * 32;              // [G][32] level-1 rows
                u64 *buf2 = slots + (size_t)2 * G * 32 + (size_t)(t & 1) * G * 32; // [G][32] level-2 rows
                if (lane < gsize)
                    st_relaxed(buf1 + (size_t)(g * gs + lane) * 32 + r, key);
                const u64 filler = (~0ull << 8) | tag;
                u64 v;
                for (;;)
                {
                    v = lane < gsize ? ld_relaxed(buf1 + (size_t)b * 32 + lane) : filler;
                    if (__all_sync(0xffffffffu, (v & 0xff) == tag))
                        break;
                }
                m = warp_min(v);
                // level 2: rank r of group g pushes to rank r' = min(r, size-1) of every group; groups may be ragged, so
                // every CTA pushes to ALL ranks r2 with r2 % gs == r ... keep it simple: push to CTA (g2*gs + r) if it exists,
                // and the last (ragged) group's missing ranks are covered by rank r pushing to everyone in a short group
                for (int g2 = lane; g2 < ng; g2 += 32)
                {
                    const int size2 = (g2 == ng - 1) ? G - g2 * gs : gs;
                    if (r < size2)
                        st_relaxed(buf2 + (size_t)(g2 * gs + r) * 32 + g, m);
                }
                // a CTA of rank r needs every group's min; groups whose size <= r (only the ragged last one can be) have no
                // rank r: then rank 0 of that group pushes for them
                const int lastSize = G - (ng - 1) * gs;
                if (g == ng - 1 && r == 0 && lastSize < gs)
                    for (int r2 = lastSize; r2 < gs; ++r2)
                        for (int g2 = lane; g2 < ng - 1; g2 += 32)
                            st_relaxed(buf2 + (size_t)(g2 * gs + r2) * 32 + g, m);
                for (;;)
                {
                    v = lane < ng ? ld_relaxed(buf2 + (size_t)b * 32 + lane) : filler;
                    if (__all_sync(0xffffffffu, (v & 0xff) == tag))
                        break;
                }
                m = warp_min(v);
            }
            if (lane == 0)
                sRes = m;
        }
        __syncthreads();
        acc ^= sRes;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0)
    {
        out[b] = acc;
        cyc[b] = t1 - t0;
    }
}

// ping-pong between CTA 0 and CTA `peer`: round-trip latency of store -> visible to a polling peer -> store back
__global__ void pingpong(u64 *slots, long long *cyc, int iters, int peer)
{
    if (threadIdx.x != 0)
        return;
    const int b = blockIdx.x;
    if (b != 0 && b != peer)
        return;
    u64 *a = slots, *c = slots + 64;
    long long t0 = clock64();
    for (int t = 1; t <= iters; ++t)
    {
        if (b == 0)
        {
            st_relaxed(a, (u64)t);
            while (ld_relaxed(c) != (u64)t)
                ;
        }
        else
        {
            while (ld_relaxed(a) != (u64)t)
                ;
            st_relaxed(c, (u64)t);
        }
    }
    if (b == 0)
        cyc[0] = clock64() - t0;
}

#define CK(x)                                                                                  \
    do                                                                                         \
    {                                                                                          \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess)                                                                  \
        {                                                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

int main()
{
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    u64 *slots, *out;
    long long *cyc;
    const size_t slotBytes = sizeof(u64) * 4 * 160 * 160;
    CK(cudaMalloc(&slots, slotBytes));
    CK(cudaMalloc(&out, sizeof(u64) * 256));
    CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    const int iters = 20000;
    printf("variant,G,threads,gs,work,cycles_per_round\n");
    struct Cfg
    {
        int variant, G, threads, gs, work;
    };
    std::vector<Cfg> cfgs;
    for (int G : {sms, 128, 64, 32, 16, 8, 2})
        for (int th : {32, 1024})
            cfgs.push_back({0, G, th, 0, 0});
    cfgs.push_back({0, sms, 1024, 0, 1000});
    cfgs.push_back({0, sms, 1024, 0, 3000});
    cfgs.push_back({2, sms, 32, 0, 0});
    cfgs.push_back({2, sms, 1024, 0, 0});
    cfgs.push_back({3, sms, 32, 0, 0});
    for (int gs : {4, 8, 12, 13, 16, 32})
        cfgs.push_back({4, sms, 32, gs, 0});
    cfgs.push_back({4, sms, 1024, 12, 0});
    for (auto c : cfgs)
    {
        CK(cudaMemset(slots, 0xff, slotBytes));
        if (c.variant == 3)
        {
            CK(cudaMemset(slots + 64, 0, 8));
        }
        int it = iters, variant = c.variant, gs = c.gs, work = c.work;
        void *args[] = {&slots, &out, &cyc, &it, &variant, &gs, &work};
        CK(cudaLaunchCooperativeKernel((void *)bench, dim3(c.G), dim3(c.threads), args, 0, 0));
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(c.G);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * c.G, cudaMemcpyDeviceToHost));
        double s = 0;
        for (auto v : h)
            s += (double)v;
        printf("%d,%d,%d,%d,%d,%.1f\n", c.variant, c.G, c.threads, c.gs, c.work, s / c.G / iters - c.work);
    }
    for (int peer : {1, 2, 37, 74, 100, 147})
    {
        CK(cudaMemset(slots, 0, 1024));
        int it = 20000;
        void *args[] = {&slots, &cyc, &it, &peer};
        CK(cudaLaunchCooperativeKernel((void *)pingpong, dim3(sms), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize());
        long long h;
        CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        printf("pingpong,peer=%d,,,,%.1f\n", peer, (double)h / it);
    }
    return 0;
}

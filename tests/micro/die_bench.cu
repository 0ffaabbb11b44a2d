// Micro-benchmark (measurement tool, not product code): which die is an SM on, which die is a 2 KB block of global
// memory homed on, and what does the push/poll exchange cost when every CTA's poll row is homed on its own die.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
typedef unsigned long long u64;

__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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

// latency of a strong load from every CTA to every 2 KB block; one CTA at a time would be cleaner, but 148 single loads
// do not load the L2 at all
__global__ void latmap(const u64 *pool, int blocks, int *lat, int *smid)
{
    const int b = blockIdx.x;
    if (threadIdx.x != 0)
        return;
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid[b] = (int)s;
    for (int k = 0; k < blocks; ++k)
    {
        const u64 *p = pool + (size_t)k * 256; // 2 KB apart
        u64 x = 0;
        x += ld_relaxed(p + (x & 1)); // warm
        long long t0 = clock64();
        for (int i = 0; i < 32; ++i)
            x += ld_relaxed(p + (x & 1)); // dependent chain (pool is zero)
        long long t1 = clock64();
        lat[(size_t)b * blocks + k] = (int)((t1 - t0) / 32) + (int)(x & 1);
    }
}

__global__ void pingpong(u64 *f1, u64 *f2, long long *cyc, int iters, int ctaA, int ctaB)
{
    if (threadIdx.x != 0)
        return;
    const int b = blockIdx.x;
    if (b != ctaA && b != ctaB)
        return;
    long long t0 = clock64();
    for (int t = 1; t <= iters; ++t)
    {
        if (b == ctaA)
        {
            st_relaxed(f1, (u64)t);
            while (ld_relaxed(f2) != (u64)t)
                ;
        }
        else
        {
            while (ld_relaxed(f1) != (u64)t)
                ;
            st_relaxed(f2, (u64)t);
        }
    }
    if (b == ctaA)
        cyc[0] = clock64() - t0;
}

// push all-to-all with a per-CTA row pointer table: rows[(t&1)*G + b] is where CTA b polls
__global__ void __launch_bounds__(1024, 1) exch(u64 *const *rows, u64 *out, long long *cyc, int iters)
{
    const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ u64 sRes;
    __shared__ u64 *sRow[2][160];
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
        sRow[i / G][i % G] = rows[i];
    __syncthreads();
    u64 acc = 0;
    long long t0 = clock64();
    for (int t = 0; t < iters; ++t)
    {
        const u64 tag = (u64)((t >> 1) & 0xff);
        u64 key = (((u64)((b * 2654435761u + t * 40503u) & 0xffffffu)) << 8) | tag;
        if (warp == 0)
        {
            u64 m;
            for (int d = lane; d < G; d += 32)
                st_relaxed(sRow[t & 1][d] + b, key);
            const u64 *row = sRow[t & 1][b];
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

#define CK(x)                                                                              \
    do                                                                                     \
    {                                                                                      \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess)                                                              \
        {                                                                                  \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

int main()
{
    int sms = 0;
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = 1024; // 2 MB pool
    u64 *pool, *out;
    long long *cyc;
    int *lat, *smid;
    CK(cudaMalloc(&pool, (size_t)blocks * 2048));
    CK(cudaMemset(pool, 0, (size_t)blocks * 2048));
    CK(cudaMalloc(&out, 8 * 256));
    CK(cudaMalloc(&cyc, 8 * 256));
    CK(cudaMalloc(&lat, sizeof(int) * sms * blocks));
    CK(cudaMalloc(&smid, sizeof(int) * sms));
    {
        int nb = blocks;
        void *args[] = {&pool, &nb, &lat, &smid};
        CK(cudaLaunchCooperativeKernel((void *)latmap, dim3(sms), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize());
    }
    std::vector<int> hl((size_t)sms * blocks), hs(sms);
    CK(cudaMemcpy(hl.data(), lat, sizeof(int) * hl.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hs.data(), smid, sizeof(int) * sms, cudaMemcpyDeviceToHost));
    // die of a CTA: relative to block 0
    std::vector<int> col(sms);
    for (int b = 0; b < sms; ++b)
        col[b] = hl[(size_t)b * blocks];
    std::vector<int> sorted = col;
    std::sort(sorted.begin(), sorted.end());
    printf("block0 latency by CTA: min %d p25 %d median %d p75 %d max %d\n", sorted[0], sorted[sms / 4], sorted[sms / 2], sorted[3 * sms / 4], sorted[sms - 1]);
    // threshold = biggest gap
    int thr = 0, gap = 0;
    for (int i = 1; i < sms; ++i)
        if (sorted[i] - sorted[i - 1] > gap)
        {
            gap = sorted[i] - sorted[i - 1];
            thr = (sorted[i] + sorted[i - 1]) / 2;
        }
    std::vector<int> die(sms);
    int n0 = 0;
    for (int b = 0; b < sms; ++b)
    {
        die[b] = col[b] > thr ? 1 : 0;
        n0 += die[b] == 0;
    }
    printf("threshold %d (gap %d): %d CTAs near block 0, %d far\n", thr, gap, n0, sms - n0);
    printf("cta:smid:die ");
    for (int b = 0; b < sms; ++b)
        printf("%d:%d:%d ", b, hs[b], die[b]);
    printf("\n");
    // home die of every block: majority of (die-0 CTAs see it fast)
    std::vector<int> home(blocks);
    int h0 = 0;
    double nearSum = 0, farSum = 0;
    long nearN = 0, farN = 0;
    for (int k = 0; k < blocks; ++k)
    {
        double s0 = 0, s1 = 0;
        int c0 = 0, c1 = 0;
        for (int b = 0; b < sms; ++b)
            if (die[b] == 0)
                s0 += hl[(size_t)b * blocks + k], ++c0;
            else
                s1 += hl[(size_t)b * blocks + k], ++c1;
        s0 /= c0;
        s1 /= c1;
        home[k] = s0 < s1 ? 0 : 1;
        h0 += home[k] == 0;
        nearSum += std::min(s0, s1);
        farSum += std::max(s0, s1);
        ++nearN;
        ++farN;
    }
    printf("blocks homed on die 0: %d of %d; mean strong-load latency near %.1f far %.1f\n", h0, blocks, nearSum / nearN, farSum / farN);
    printf("home of first 64 blocks: ");
    for (int k = 0; k < 64; ++k)
        printf("%d", home[k]);
    printf("\n");

    // ping-pong matrix: A on die 0, B on die 0 or 1; flags homed on 0 or 1
    int a0 = -1, a0b = -1, b1 = -1;
    for (int b = 0; b < sms; ++b)
    {
        if (die[b] == 0 && a0 < 0)
            a0 = b;
        else if (die[b] == 0 && a0b < 0 && hs[b] / 2 != hs[a0] / 2)
            a0b = b;
        if (die[b] == 1 && b1 < 0)
            b1 = b;
    }
    int blk0 = -1, blk1 = -1, blk0b = -1, blk1b = -1;
    for (int k = 1; k < blocks; ++k)
    {
        if (home[k] == 0 && blk0 < 0)
            blk0 = k;
        else if (home[k] == 0 && blk0b < 0)
            blk0b = k;
        if (home[k] == 1 && blk1 < 0)
            blk1 = k;
        else if (home[k] == 1 && blk1b < 0)
            blk1b = k;
    }
    printf("pingpong: A=cta %d (die0), same-die peer cta %d, cross-die peer cta %d; blocks homed die0: %d,%d die1: %d,%d\n", a0, a0b, b1, blk0, blk0b, blk1, blk1b);
    struct PP
    {
        const char *name;
        int A, B, f1, f2;
    };
    // f1 is written by A and polled by B; f2 is written by B and polled by A
    std::vector<PP> pps = {
        {"same die, flags near both", a0, a0b, blk0, blk0b},
        {"same die, flags far from both", a0, a0b, blk1, blk1b},
        {"cross die, flags homed at POLLER", a0, b1, blk1, blk0},
        {"cross die, flags homed at WRITER", a0, b1, blk0, blk1},
        {"cross die, both flags on die 0", a0, b1, blk0, blk0b},
        {"cross die, both flags on die 1", a0, b1, blk1, blk1b},
    };
    for (auto &pp : pps)
    {
        CK(cudaMemset(pool, 0, (size_t)blocks * 2048));
        u64 *f1 = pool + (size_t)pp.f1 * 256, *f2 = pool + (size_t)pp.f2 * 256;
        int it = 20000;
        void *args[] = {&f1, &f2, &cyc, &it, &pp.A, &pp.B};
        CK(cudaLaunchCooperativeKernel((void *)pingpong, dim3(sms), dim3(32), args, 0, 0));
        CK(cudaDeviceSynchronize());
        long long h;
        CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        printf("pingpong round trip, %-36s: %.1f cycles\n", pp.name, (double)h / it);
    }

    // exchange with rows homed (a) at the poller's die, (b) at the other die, (c) arbitrary (consecutive blocks)
    for (int mode = 0; mode < 3; ++mode)
    {
        std::vector<u64 *> rows(2 * sms);
        int next[2] = {1, 1};
        int seq = 1;
        for (int i = 0; i < 2 * sms; ++i)
        {
            const int b = i % sms;
            int k;
            if (mode == 2)
                k = seq++;
            else
            {
                const int want = mode == 0 ? die[b] : 1 - die[b];
                int &n = next[want];
                while (home[n] != want)
                    ++n;
                k = n++;
            }
            rows[i] = pool + (size_t)k * 256;
        }
        u64 **drows;
        CK(cudaMalloc(&drows, sizeof(u64 *) * rows.size()));
        CK(cudaMemcpy(drows, rows.data(), sizeof(u64 *) * rows.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(pool, 0xff, (size_t)blocks * 2048));
        for (int th : {32, 1024})
        {
            int it = 20000;
            void *args[] = {&drows, &out, &cyc, &it};
            CK(cudaMemset(pool, 0xff, (size_t)blocks * 2048));
            CK(cudaLaunchCooperativeKernel((void *)exch, dim3(sms), dim3(th), args, 0, 0));
            CK(cudaDeviceSynchronize());
            std::vector<long long> h(sms);
            CK(cudaMemcpy(h.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost));
            double s = 0;
            for (auto v : h)
                s += (double)v;
            printf("exchange G=%d threads=%d rows %s: %.1f cycles/round\n", sms, th, mode == 0 ? "homed at poller" : mode == 1 ? "homed at other die" : "consecutive", s / sms / it);
        }
        CK(cudaFree(drows));
    }
    return 0;
}

// Micro-benchmark (measurement tool, not product code): classify SMs and 2 KB memory blocks by L2 die using the round
// trip of strong store -> poll ping-pongs, then time the push/poll min-loc exchange with die-aware row placement.
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

__global__ void pingpong(u64 *f1, u64 *f2, long long *cyc, int iters, int ctaA, int ctaB, u64 base)
{
    if (threadIdx.x != 0)
        return;
    const int b = blockIdx.x;
    if (b != ctaA && b != ctaB)
        return;
    long long t0 = clock64();
    for (int t = 1; t <= iters; ++t)
    {
        const u64 v = base + t;
        if (b == ctaA)
        {
            st_relaxed(f1, v);
            while (ld_relaxed(f2) != v)
                ;
        }
        else
        {
            while (ld_relaxed(f1) != v)
                ;
            st_relaxed(f2, v);
        }
    }
    if (b == ctaA)
        cyc[0] = clock64() - t0;
}

// push all-to-all among the CTAs with part[b] >= 0; part[b] = index of the CTA among the participants, P of them;
// rows[(t&1)*G + b] = row CTA b polls (P words used); dest list = all participants
__global__ void __launch_bounds__(1024, 1) exch(u64 *const *rows, const int *part, const int *members, int P, u64 *out, long long *cyc, int iters)
{
    const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ u64 sRes;
    __shared__ u64 *sRow[2][160];
    __shared__ int sMem[160];
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
        sRow[i / G][i % G] = rows[i];
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        sMem[i] = members[i];
    __syncthreads();
    const int me = part[b];
    if (me < 0)
        return;
    u64 acc = 0;
    long long t0 = clock64();
    for (int t = 0; t < iters; ++t)
    {
        const u64 tag = (u64)((t >> 1) & 0xff);
        u64 key = (((u64)((b * 2654435761u + t * 40503u) & 0xffffffu)) << 8) | tag;
        if (warp == 0)
        {
            u64 m;
            for (int d = lane; d < P; d += 32)
                st_relaxed(sRow[t & 1][sMem[d]] + me, key);
            const u64 *row = sRow[t & 1][b];
            const u64 filler = (~0ull << 8) | tag;
            for (;;)
            {
                u64 v[5];
#pragma unroll
                for (int j = 0; j < 5; ++j)
                {
                    int i = lane + 32 * j;
                    v[j] = i < P ? ld_relaxed(row + i) : filler;
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

static u64 *pool;
static long long *cyc;
static int sms;
static u64 ppBase = 0;
static double pp(int A, int B, int blk1, int blk2, int iters = 1500)
{
    u64 *f1 = pool + (size_t)blk1 * 256, *f2 = pool + (size_t)blk2 * 256 + 16;
    void *args[] = {&f1, &f2, &cyc, &iters, &A, &B, &ppBase};
    CK(cudaLaunchCooperativeKernel((void *)pingpong, dim3(sms), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    ppBase += iters + 8;
    long long h;
    CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    return (double)h / iters;
}

int main()
{
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int blocks = 1024;
    u64 *out;
    CK(cudaMalloc(&pool, (size_t)blocks * 2048));
    CK(cudaMemset(pool, 0, (size_t)blocks * 2048));
    CK(cudaMalloc(&out, 8 * 256));
    CK(cudaMalloc(&cyc, 8 * 256));

    // 1. peers of CTA 0 through block 0
    std::vector<double> rt(sms, 0);
    for (int p = 1; p < sms; ++p)
        rt[p] = pp(0, p, 0, 0);
    std::vector<double> s(rt.begin() + 1, rt.end());
    std::sort(s.begin(), s.end());
    printf("ping-pong CTA0<->peer via block 0: min %.0f p10 %.0f p25 %.0f median %.0f p75 %.0f p90 %.0f max %.0f\n", s[0], s[s.size() / 10], s[s.size() / 4], s[s.size() / 2], s[3 * s.size() / 4], s[9 * s.size() / 10], s.back());
    double gap = 0, thr = 0;
    for (size_t i = 1; i < s.size(); ++i)
        if (s[i] - s[i - 1] > gap)
        {
            gap = s[i] - s[i - 1];
            thr = 0.5 * (s[i] + s[i - 1]);
        }
    std::vector<int> die(sms, 0);
    int n0 = 1;
    for (int p = 1; p < sms; ++p)
    {
        die[p] = rt[p] > thr ? 1 : 0;
        n0 += die[p] == 0;
    }
    printf("threshold %.0f (gap %.0f): %d CTAs in CTA 0's class, %d in the other\n", thr, gap, n0, sms - n0);
    printf("round trips: ");
    for (int p = 1; p < sms; ++p)
        printf("%.0f ", rt[p]);
    printf("\n");
    int peer0 = -1, peer1 = -1, peer1b = -1;
    for (int p = 1; p < sms; ++p)
    {
        if (die[p] == 0 && peer0 < 0)
            peer0 = p;
        if (die[p] == 1 && peer1 < 0)
            peer1 = p;
        else if (die[p] == 1 && peer1b < 0)
            peer1b = p;
    }
    // 2. home class of blocks: ping-pong between two class-0 CTAs and between two class-1 CTAs
    const int NB = 800;
    std::vector<double> b0(NB), b1(NB);
    for (int k = 0; k < NB; ++k)
    {
        b0[k] = pp(0, peer0, k, k, 400);
        b1[k] = k < 24 ? pp(peer1, peer1b, k, k, 400) : 1500.0;
    }
    printf("block k: class-0 pair rt / class-1 pair rt\n");
    std::vector<int> home(NB);
    int h0 = 0;
    for (int k = 0; k < NB; ++k)
    {
        home[k] = b0[k] < b1[k] ? 0 : 1;
        h0 += home[k] == 0;
        if (k < 24)
            printf("  %d: %.0f / %.0f\n", k, b0[k], b1[k]);
    }
    printf("%d of %d blocks homed with class 0\n", h0, NB);
    // 3. cross pairs
    int blkH0 = -1, blkH1 = -1;
    for (int k = 1; k < NB; ++k)
    {
        if (home[k] == 0 && blkH0 < 0)
            blkH0 = k;
        if (home[k] == 1 && blkH1 < 0)
            blkH1 = k;
    }
    printf("cross pair (CTA0 class0 <-> CTA%d class1): flag polled by each homed at poller %.0f, homed at writer %.0f, both class0 %.0f, both class1 %.0f\n", peer1,
           pp(0, peer1, blkH1, blkH0), pp(0, peer1, blkH0, blkH1), pp(0, peer1, blkH0, blkH0), pp(0, peer1, blkH1, blkH1));

    // 4. exchanges
    std::vector<int> blk0, blk1;
    for (int k = 1; k < NB; ++k)
        (home[k] == 0 ? blk0 : blk1).push_back(k);
    auto run = [&](const char *name, int rowMode, int who) {
        // who: -1 all CTAs, 0 / 1 only that class;  rowMode 0: homed at poller, 1: homed at other class, 2: consecutive
        std::vector<int> part(sms, -1), members;
        for (int b = 0; b < sms; ++b)
            if (who < 0 || die[b] == who)
            {
                part[b] = (int)members.size();
                members.push_back(b);
            }
        std::vector<u64 *> rows(2 * sms);
        size_t i0 = 0, i1 = 0;
        int seq = 1;
        for (int i = 0; i < 2 * sms; ++i)
        {
            const int b = i % sms;
            int k;
            if (rowMode == 2)
                k = seq++;
            else
            {
                const int want = rowMode == 0 ? die[b] : 1 - die[b];
                if (want == 0)
                    k = blk0[i0++ % blk0.size()];
                else
                    k = blk1[i1++ % blk1.size()];
            }
            // rows sharing a block (when the pool of classified blocks is small) use different halves / quarters
            rows[i] = pool + (size_t)k * 256;
        }
        // avoid aliasing: give every row its own block by extending with unclassified blocks only in mode 2; for
        // modes 0/1 make sure no two rows share a block
        if (rowMode != 2 && (i0 > blk0.size() || i1 > blk1.size()))
        {
            printf("%s: not enough classified blocks (%zu/%zu needed %zu/%zu)\n", name, blk0.size(), blk1.size(), i0, i1);
            return;
        }
        u64 **drows;
        int *dpart, *dmem;
        CK(cudaMalloc(&drows, sizeof(u64 *) * rows.size()));
        CK(cudaMemcpy(drows, rows.data(), sizeof(u64 *) * rows.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dpart, 4 * sms));
        CK(cudaMemcpy(dpart, part.data(), 4 * sms, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dmem, 4 * sms));
        CK(cudaMemcpy(dmem, members.data(), 4 * members.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(pool, 0xff, (size_t)blocks * 2048));
        int it = 20000, P = (int)members.size();
        void *args[] = {&drows, &dpart, &dmem, &P, &out, &cyc, &it};
        CK(cudaLaunchCooperativeKernel((void *)exch, dim3(sms), dim3(1024), args, 0, 0));
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost));
        double sum = 0;
        for (int b : members)
            sum += (double)h[b];
        printf("exchange %-46s P=%3d: %.1f cycles/round\n", name, P, sum / P / it);
        cudaFree(drows);
        cudaFree(dpart);
        cudaFree(dmem);
    };
    // need 2*sms blocks per mode: classify more blocks if necessary
    run("all CTAs, rows consecutive", 2, -1);
    run("all CTAs, rows homed at poller", 0, -1);
    run("all CTAs, rows homed at other class", 1, -1);
    run("class 0 only, rows homed at poller", 0, 0);
    run("class 1 only, rows homed at poller", 0, 1);
    run("class 0 only, rows homed at other class", 1, 0);
    run("class 0 only, rows consecutive", 2, 0);
    return 0;
}

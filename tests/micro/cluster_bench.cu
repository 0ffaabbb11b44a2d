// Micro-benchmark (measurement tool, not product code): two-level min-loc exchange — DSMEM inside a thread-block cluster,
// L2 push/poll between the equal-rank CTAs of different clusters.
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
__device__ __forceinline__ void st_cluster(unsigned remoteAddr, u64 v) { asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(remoteAddr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_local_relaxed(unsigned addr)
{
    u64 v;
    asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// mode 0: DSMEM all-to-all inside the cluster only (no global phase)
// mode 1: DSMEM all-to-all, then L2 push/poll among equal ranks of the NC clusters
__global__ void __launch_bounds__(1024, 1) exch(u64 *slots, u64 *out, long long *cyc, int iters, int mode)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = cluster.num_blocks(), rank = cluster.block_rank();
    const int b = blockIdx.x, cid = b / C, NC = gridDim.x / C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ __align__(16) u64 cl[2][16];
    __shared__ u64 sRes;
    if (threadIdx.x < 32)
        cl[threadIdx.x / 16][threadIdx.x % 16] = ~0ull;
    cluster.sync();
    const unsigned clBase = static_cast<unsigned>(__cvta_generic_to_shared(&cl[0][0]));
    u64 acc = 0;
    long long t0 = clock64();
    for (int t = 0; t < iters; ++t)
    {
        const u64 tag = (u64)((t >> 1) & 0xff);
        u64 key = (((u64)((b * 2654435761u + t * 40503u) & 0xffffffu)) << 8) | tag;
        if (warp == 0)
        {
            const u64 filler = (~0ull << 8) | tag;
            const unsigned mine = clBase + ((t & 1) * 16 + rank) * 8;
            if (lane < C)
                st_cluster(mapa(mine, lane), key);
            u64 v;
            const unsigned pollAddr = clBase + ((t & 1) * 16 + lane) * 8;
            for (;;)
            {
                v = lane < C ? ld_local_relaxed(pollAddr) : filler;
                if (__all_sync(0xffffffffu, (v & 0xff) == tag))
                    break;
            }
            u64 m = warp_min(v);
            if (mode == 1)
            {
                u64 *buf = slots + (size_t)(t & 1) * gridDim.x * 32;
                if (lane < NC)
                    st_relaxed(buf + (size_t)(lane * C + rank) * 32 + cid, m);
                const u64 *row = buf + (size_t)b * 32;
                for (;;)
                {
                    v = lane < NC ? ld_relaxed(row + lane) : filler;
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
    cluster.sync();
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
    setvbuf(stdout, NULL, _IONBF, 0);
    CK(cudaSetDevice(0));
    u64 *slots, *out;
    long long *cyc;
    CK(cudaMalloc(&slots, 8 * 2 * 2048 * 32));
    CK(cudaMalloc(&out, 8 * 256));
    CK(cudaMalloc(&cyc, 8 * 256));
    CK(cudaFuncSetAttribute(exch, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int threads : {1024})
        for (int C : {2, 4, 8, 16})
        {
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = 0;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeCooperative;
            at[1].val.cooperative = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(C); // occupancy query wants a grid that is a multiple of the cluster
            int maxClusters = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&maxClusters, exch, &cfg);
            if (e != cudaSuccess)
            {
                printf("C=%d: occupancy query failed: %s\n", C, cudaGetErrorString(e));
                cudaGetLastError();
                continue;
            }
            printf("C=%d threads=%d: max active clusters %d (%d CTAs)\n", C, threads, maxClusters, maxClusters * C);
            if (maxClusters > 32)
                maxClusters = 32;
            for (int mode = 0; mode < 2; ++mode)
                for (int coop = 0; coop < 1; ++coop)
                {
                    cfg.gridDim = dim3(maxClusters * C);
                    cfg.numAttrs = coop ? 2 : 1;
                    CK(cudaMemset(slots, 0xff, 8 * 2 * 2048 * 32));
                    int iters = 20000;
                    e = cudaLaunchKernelEx(&cfg, exch, slots, out, cyc, iters, mode);
                    if (e != cudaSuccess)
                    {
                        printf("  mode %d coop %d: launch failed: %s\n", mode, coop, cudaGetErrorString(e));
                        cudaGetLastError();
                        continue;
                    }
                    CK(cudaDeviceSynchronize());
                    std::vector<long long> h(maxClusters * C);
                    CK(cudaMemcpy(h.data(), cyc, 8 * h.size(), cudaMemcpyDeviceToHost));
                    double s = 0;
                    for (auto v : h)
                        s += (double)v;
                    printf("  mode %d (%s) coop %d: %.1f cycles/round\n", mode, mode ? "DSMEM + L2 between clusters" : "DSMEM only", coop, s / h.size() / iters);
                }
        }
    return 0;
}

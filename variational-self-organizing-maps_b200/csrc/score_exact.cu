// score_exact.cu — K3: exact batch BMU search in the reference's own arithmetic.
//
// Replaces the row loops of Som::evaluate (src/Som.cpp:503-520) / Som::measureSimilarity (:641-711) and any
// per-row Som::findBmu (:291-309) / Som::findRestrictedBmu (:313-332) over a loaded DataSet chunk:
// for every row, argmin over nodes of Som::euclidianWeightedDist (:124-141) with the strict '<',
// lowest-index-wins rule, and the f32 distance itself, summed k = 0,1,2,... like `comparer.dot(comparer)`.
//
// Each (row, node) pair is one sequential f32 chain owned by one thread, so every distance is bit-identical
// to the reference; the parallelism is over pairs (a 64 x 64 pair tile per CTA, 4 x 4 chains per thread,
// operands staged through shared memory in 32-wide slices of the model vector).  This is the exact path:
// it is what the tensor-core candidate search (score_tc.cu) rescoring falls back to, and it is the scorer
// for small / medium batches.
#include "common.cuh"

namespace vsom
{

constexpr int RT = 64, NT = 64, KC = 32, kScoreThreads = 256;

template <int TR>
__global__ void __launch_bounds__(kScoreThreads) find_bmu_exact_kernel(const float *__restrict__ x, u64 n, const float *__restrict__ mean,
                                                                       const u64 *__restrict__ hits, u64 minHits, int N, int Din, int Dr, int P,
                                                                       int rowStride, const unsigned short *__restrict__ pairI,
                                                                       const unsigned short *__restrict__ pairJ, unsigned *__restrict__ outBmu,
                                                                       float *__restrict__ outDist)
{
    __shared__ float xa[RT][KC + 1];
    __shared__ float ma[NT][KC + 1];
    __shared__ float xb[TR == VSOM_CLR ? RT : 1][KC + 1];
    __shared__ float mb[TR == VSOM_CLR ? NT : 1][KC + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4; // nodes tx + 16 j, rows ty + 16 i
    const u64 row0 = static_cast<u64>(blockIdx.x) * RT;

    u64 best[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        best[i] = ~0ull;

    for (int node0 = 0; node0 < N; node0 += NT)
    {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                acc[i][j] = 0.0f;

        for (int k0 = 0; k0 < Dr; k0 += KC)
        {
            __syncthreads(); // previous slice fully consumed
            const int k = k0 + lane;
            // stage the slice: one warp per row / node, lanes along k (128-byte coalesced segments)
            for (int r = warp; r < RT; r += kScoreThreads / 32)
            {
                const u64 row = row0 + r;
                float a = 0.0f, bq = 0.0f;
                if (row < n && k < Dr)
                {
                    if (TR != VSOM_CLR)
                        a = x[row * Din + k];
                    else
                    {
                        a = x[row * Din + pairI[k]];
                        bq = x[row * Din + pairJ[k]];
                    }
                }
                xa[r][lane] = a;
                if (TR == VSOM_CLR)
                    xb[r][lane] = bq;
            }
            for (int r = warp; r < NT; r += kScoreThreads / 32)
            {
                const int node = node0 + r;
                float a = 0.0f, bq = 0.0f;
                if (node < N && k < Dr)
                {
                    a = mean[static_cast<size_t>(node) * rowStride + k];
                    if (TR == VSOM_CLR)
                        bq = mean[static_cast<size_t>(node) * rowStride + P + k];
                }
                ma[r][lane] = a;
                if (TR == VSOM_CLR)
                    mb[r][lane] = bq;
            }
            __syncthreads();
            // padded entries are all-zero on both sides: their residual is +0 and s + 0 == s exactly.
#pragma unroll 4
            for (int kk = 0; kk < KC; ++kk)
            {
                float xv[4], mv[4], yv[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    xv[i] = xa[ty + 16 * i][kk];
                    mv[i] = ma[tx + 16 * i][kk];
                    if (TR == VSOM_CLR)
                    {
                        yv[i] = xb[ty + 16 * i][kk];
                        bv[i] = mb[tx + 16 * i][kk];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        float r;
                        if (TR != VSOM_CLR)
                            r = __fsub_rn(mv[j], xv[i]);
                        else
                            r = __fsub_rn(__fadd_rn(__fmul_rn(mv[j], xv[i]), bv[j]), yv[i]);
                        acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(r, r));
                    }
            }
        }
        // fold this node tile into the running min of each row (node 0 seeds regardless of its hit count)
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const int node = node0 + tx + 16 * j;
            if (node < N && (node == 0 || minHits == 0 || hits[node] >= minHits))
            {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const float d = acc[i][j];
                    best[i] = u64_min(best[i], make_key(d, static_cast<unsigned>(node), (d != d) ? 1u : 0u));
                }
            }
        }
    }
    // the 16 threads sharing a row sit in one half-warp: xor offsets 8,4,2,1 stay inside it
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        u64 k = best[i];
#pragma unroll
        for (int o = 8; o; o >>= 1)
            k = u64_min(k, __shfl_xor_sync(0xffffffffu, k, o));
        const u64 row = row0 + ty + 16 * i;
        if (tx == 0 && row < n)
        {
            if (outBmu)
                outBmu[row] = key_node(k);
            if (outDist)
            {
                float d = __uint_as_float(static_cast<unsigned>(k >> 32));
                if (k & 1ull)
                    d = __uint_as_float(0x7fc00000u);
                outDist[row] = d;
            }
        }
    }
}

template <int TR>
__global__ void all_dists_kernel(const float *__restrict__ v, const float *__restrict__ mean, int N, int Dr, int P, int rowStride,
                                 const unsigned short *__restrict__ pairI, const unsigned short *__restrict__ pairJ, double *__restrict__ out)
{
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= N)
        return;
    out[node] = static_cast<double>(dist_sequential<TR>(mean + static_cast<size_t>(node) * rowStride, v, Dr, P, pairI, pairJ));
}

int launch_find_bmu(vsom_ctx *ctx, const float *xDev, size_t n, uint64_t minHits, unsigned *outBmuDev, float *outDistDev)
{
    if (n == 0)
        return VSOM_OK;
    const unsigned grid = static_cast<unsigned>((n + RT - 1) / RT);
#define VSOM_LAUNCH_SCORE(TR)                                                                                                              \
    find_bmu_exact_kernel<TR><<<grid, kScoreThreads, 0, ctx->stream>>>(xDev, n, ctx->mean, ctx->hits, minHits, ctx->N, ctx->Din, ctx->Dr,   \
                                                                       ctx->P, ctx->rowStride, ctx->pairI, ctx->pairJ, outBmuDev, outDistDev)
    switch (ctx->transform)
    {
    case VSOM_STANDARD:
    case VSOM_MEDIAN: // same Comparer (src/Transformation.cpp:7-8, :45-46)
        VSOM_LAUNCH_SCORE(VSOM_STANDARD);
        break;
    default:
        VSOM_LAUNCH_SCORE(VSOM_CLR);
        break;
    }
#undef VSOM_LAUNCH_SCORE
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

int launch_all_dists(vsom_ctx *ctx, const float *vDev, double *outDev)
{
    const int threads = 128, grid = (ctx->N + threads - 1) / threads;
    if (ctx->transform == VSOM_CLR)
        all_dists_kernel<VSOM_CLR><<<grid, threads, 0, ctx->stream>>>(vDev, ctx->mean, ctx->N, ctx->Dr, ctx->P, ctx->rowStride, ctx->pairI, ctx->pairJ, outDev);
    else
        all_dists_kernel<VSOM_STANDARD><<<grid, threads, 0, ctx->stream>>>(vDev, ctx->mean, ctx->N, ctx->Dr, ctx->P, ctx->rowStride, ctx->pairI, ctx->pairJ, outDev);
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

} // namespace vsom

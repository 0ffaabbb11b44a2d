// batch_map.cu — K6: one chunk-epoch of the batch-map trainer (a "next" row of the scope table).
//
// Replaces Som::trainBatchSomEpoch (src/Som.cpp:756-879), in the order the reference has when its parallel algorithms
// run on the sequential backend (rows in order, neurons independent):
//   phase A  per row: BMU — global on the first epoch (findBmu, :771), the greedy local walk from the row's lastBMU
//            afterwards (findLocalBmu, :793) — hit count and squared residual;
//   phase B  per neuron: West/Finch incremental weighted mean and variance over ALL rows of the chunk in row order,
//            weights = neighbourhood of the row's BMU evaluated with the reference's SomIndex(map, index) coordinates
//            (which divide by the map HEIGHT, src/SomIndex.cpp:15-18):
//                sumW += w;  delta = Stepper(x, cur);  cur += (w / sumW) * delta;  S += (w * delta) * delta
//            then map = cur, sigmaMap = sqrt(S / sumW), weightMap = sumW  (:871-875).
// Every (neuron, component) pair is a strictly sequential f32 chain over the rows, so the result is bit-identical to
// the reference; the parallelism is over chains: a CTA owns a few neurons, streams the rows through shared memory in
// batches, a handful of threads turn (row, neuron) into the two coefficients w and w / sumW (one IEEE division per
// pair instead of one per chain), and all threads then advance their chains (up to 16 per thread, in registers).
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace vsom
{

constexpr int BM_THREADS = 256;
constexpr int BM_CPT = 16;    // chains per thread (Standard / Median); CLR keeps two models per chain and uses 8
constexpr int BM_ROWS = 16;   // rows per shared-memory batch
constexpr int BM_MAXNB = 32;  // neurons per CTA at most

// reference coordinates of a linear index: SomIndex(const Som&, index), src/SomIndex.cpp:15-18
__host__ __device__ inline void quirky_xy(unsigned index, int W, int H, int &x, int &y)
{
    x = static_cast<int>(index % static_cast<unsigned>(W));
    y = static_cast<int>((index - index % static_cast<unsigned>(W)) / static_cast<unsigned>(H));
}

__global__ void bmu_xy_kernel(const unsigned *__restrict__ bmu, u64 n, int W, int H, int2 *__restrict__ out)
{
    const u64 j = static_cast<u64>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    int x, y;
    quirky_xy(bmu[j], W, H, x, y);
    out[j] = make_int2(x, y);
}

__global__ void hits_add_kernel(const unsigned *__restrict__ bmu, u64 n, u64 *__restrict__ hits)
{
    const u64 j = static_cast<u64>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j < n)
        atomicAdd(hits + bmu[j], 1ull);
}

// ---- phase A for the later epochs: findLocalBmu (src/Som.cpp:335-454) per row, one warp per row.  Lanes 0..7 evaluate
// the candidate cells' distances (each a sequential f32 chain like the reference), every lane then replays the
// reference's comparisons on the shuffled values, so the walk state stays warp-uniform.
template <int TR>
__global__ void __launch_bounds__(256) local_bmu_rows_kernel(const float *__restrict__ x, u64 n, const float *__restrict__ mean, int W, int H, int Din,
                                                             int Dr, int P, int rowStride, const unsigned short *__restrict__ pairI,
                                                             const unsigned short *__restrict__ pairJ, int order, const u64 *__restrict__ start,
                                                             unsigned *__restrict__ outBmu, float *__restrict__ outDist)
{
    const u64 row = static_cast<u64>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n)
        return;
    const float *xr = x + row * Din;
    auto dist = [&](u64 node) { return dist_ordered<TR>(mean + node * rowStride, xr, Dr, P, pairI, pairJ, order); };
    const u64 uW = static_cast<u64>(W), uH = static_cast<u64>(H), M1 = ~0ull;
    const u64 fx[8] = {M1, 0, 1, 1, 1, 0, M1, M1}, fy[8] = {1, 1, 1, 0, M1, M1, M1, 0};
    u64 lastBMU = start ? start[row] : 0ull, minIndex = lastBMU, lastMeasured = lastBMU;
    float minDist = __shfl_sync(0xffffffffu, lane == 0 ? dist(lastBMU) : 0.0f, 0);
    for (;;)
    {
        const u64 lmX = lastMeasured % uW, lmY = lastMeasured / uW, lbX = lastBMU % uW;
        if (lastMeasured == lastBMU)
        {
            u64 idx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const u64 cx = lmX + fx[i] < uW - 1 ? lmX + fx[i] : uW - 1, cy = lmY + fy[i] < uH - 1 ? lmY + fy[i] : uH - 1;
                idx[i] = cy * uW + cx;
            }
            float mine = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (lane == i)
                    mine = dist(idx[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const float v = __shfl_sync(0xffffffffu, mine, i);
                if (v < minDist)
                {
                    minDist = v;
                    minIndex = idx[i];
                }
            }
            if (minIndex == lastBMU)
                break;
            lastMeasured = minIndex;
        }
        else
        {
            if (lmX - lbX)
            {
                u64 idx[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    const u64 ox = lmX + lmX - lbX, oy = lmY + static_cast<u64>(static_cast<long long>(i - 1));
                    idx[i] = (oy < uH - 1 ? oy : uH - 1) * uW + (ox < uW - 1 ? ox : uW - 1);
                }
                float mine = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (lane == i)
                        mine = dist(idx[i]);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    const float v = __shfl_sync(0xffffffffu, mine, i);
                    if (v < minDist)
                    {
                        minDist = v;
                        minIndex = idx[i];
                    }
                }
            }
            if (minIndex == lastMeasured)
                break;
            lastBMU = lastMeasured;
            lastMeasured = minIndex;
        }
    }
    if (lane == 0)
    {
        outBmu[row] = static_cast<unsigned>(minIndex);
        // the squared residual at the chosen node (src/Som.cpp:803): the same f32 sum as its distance
        outDist[row] = dist(minIndex);
    }
}

// ---- phase B
template <int TR>
__global__ void __launch_bounds__(BM_THREADS) batch_update_kernel(const float *__restrict__ x, u64 n, int Din, int Dq, int P, int W, int H, int N,
                                                                  int nbPerCta, const int2 *__restrict__ bmuXY, const float *__restrict__ lut,
                                                                  int lutW, const unsigned short *__restrict__ pairI,
                                                                  const unsigned short *__restrict__ pairJ, float *__restrict__ mean,
                                                                  float *__restrict__ sigma, float *__restrict__ weight, int rowStride)
{
    constexpr int CPT = TR == VSOM_CLR ? BM_CPT / 2 : BM_CPT;
    extern __shared__ __align__(16) float bmSmem[];
    float *xs = bmSmem;                                  // [BM_ROWS][Din]
    float *wS = xs + BM_ROWS * Din;                      // [BM_MAXNB][BM_ROWS]  w
    float *cS = wS + BM_MAXNB * BM_ROWS;                 // [BM_MAXNB][BM_ROWS]  w / sumW
    int2 *xyS = reinterpret_cast<int2 *>(cS + BM_MAXNB * BM_ROWS); // [BM_ROWS] BMU coordinates of the batch's rows
    __shared__ float sumWs[BM_MAXNB];

    const int tid = threadIdx.x;
    const int p0 = blockIdx.x * nbPerCta;
    const int nb = min(nbPerCta, N - p0);
    const int chains = nb * Dq;

    // this thread's chains: c = tid + s * BM_THREADS  ->  (neuron pl = c / Dq, component k = c % Dq)
    float cur[CPT], S[CPT], curB[TR == VSOM_CLR ? CPT : 1], SB[TR == VSOM_CLR ? CPT : 1];
    short pl[CPT];
    short kk[CPT];
#pragma unroll
    for (int s = 0; s < CPT; ++s)
    {
        const int c = tid + s * BM_THREADS;
        cur[s] = 0.0f;
        S[s] = 0.0f;
        if (TR == VSOM_CLR)
        {
            curB[s] = 0.0f;
            SB[s] = 0.0f;
        }
        pl[s] = static_cast<short>(c < chains ? c / Dq : -1);
        kk[s] = static_cast<short>(c < chains ? c % Dq : 0);
    }
    // coefficient threads: thread p < nb carries sumW of neuron p0 + p over all rows
    float sumW = 0.0f;
    int cx = 0, cy = 0;
    if (tid < nb)
        quirky_xy(static_cast<unsigned>(p0 + tid), W, H, cx, cy);

    for (u64 j0 = 0; j0 < n; j0 += BM_ROWS)
    {
        const int rows = static_cast<int>(min(static_cast<u64>(BM_ROWS), n - j0));
        __syncthreads(); // the previous batch is consumed
        for (int i = tid; i < rows * Din; i += BM_THREADS)
            xs[i] = x[j0 * Din + i];
        if (tid < rows)
            xyS[tid] = bmuXY[j0 + tid];
        __syncthreads();
        if (tid < nb)
        {
            for (int r = 0; r < rows; ++r)
            {
                const int2 b = xyS[r];
                const int dx = cx > b.x ? cx - b.x : b.x - cx, dy = cy > b.y ? cy - b.y : b.y - cy;
                const float w = __ldg(lut + dy * lutW + dx);
                sumW = __fadd_rn(sumW, w);                       // Eq. 47
                wS[tid * BM_ROWS + r] = w;
                cS[tid * BM_ROWS + r] = __fdiv_rn(w, sumW);      // currentWeight / sumOfWeights
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < CPT; ++s)
        {
            if (pl[s] < 0)
                continue;
            const float *wp = wS + pl[s] * BM_ROWS, *cp = cS + pl[s] * BM_ROWS;
            if (TR != VSOM_CLR)
            {
                float m = cur[s], acc = S[s];
                for (int r = 0; r < rows; ++r)
                {
                    float d = __fsub_rn(xs[r * Din + kk[s]], m);
                    if (TR == VSOM_MEDIAN)
                    {
                        const float a = fabsf(d);
                        d = a > 0.0f ? __uint_as_float((__float_as_uint(d) & 0x80000000u) | 0x3f800000u) : a;
                    }
                    const float w = wp[r];
                    m = __fadd_rn(m, __fmul_rn(cp[r], d));                 // Eq. 53
                    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(w, d), d));   // Eq. 68
                }
                cur[s] = m;
                S[s] = acc;
            }
            else
            {
                const int xiAt = pairI[kk[s]], xjAt = pairJ[kk[s]];
                float a = cur[s], b = curB[s], sa = S[s], sb = SB[s];
                for (int r = 0; r < rows; ++r)
                {
                    const float xi = xs[r * Din + xiAt], xj = xs[r * Din + xjAt];
                    const float inner = __fsub_rn(__fadd_rn(__fmul_rn(a, xi), b), xj);
                    const float db = __fmul_rn(-2.0f, inner), da = __fmul_rn(db, xi);
                    const float w = wp[r], c = cp[r];
                    a = __fadd_rn(a, __fmul_rn(c, da));
                    b = __fadd_rn(b, __fmul_rn(c, db));
                    sa = __fadd_rn(sa, __fmul_rn(__fmul_rn(w, da), da));
                    sb = __fadd_rn(sb, __fmul_rn(__fmul_rn(w, db), db));
                }
                cur[s] = a;
                curB[s] = b;
                S[s] = sa;
                SB[s] = sb;
            }
        }
    }
    if (tid < nb)
    {
        sumWs[tid] = sumW;
        weight[p0 + tid] = sumW; // :875
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < CPT; ++s)
    {
        if (pl[s] < 0)
            continue;
        const size_t at = static_cast<size_t>(p0 + pl[s]) * rowStride;
        const float sw = sumWs[pl[s]];
        mean[at + kk[s]] = cur[s];                                             // :871
        sigma[at + kk[s]] = __fsqrt_rn(__fdiv_rn(S[s], sw));                   // :873 (no abs)
        if (TR == VSOM_CLR)
        {
            mean[at + P + kk[s]] = curB[s];
            sigma[at + P + kk[s]] = __fsqrt_rn(__fdiv_rn(SB[s], sw));
        }
    }
}

// one chunk-epoch; x / bmu / dist on the device.  isFirst selects the global search; otherwise lastDev holds the start
// nodes of the local walks.  On return bmuDev holds every row's BMU and distDev its squared residual.
int launch_batch_epoch(vsom_ctx *ctx, const float *xDev, size_t n, double sigma, int isFirst, const u64 *lastDev, unsigned *bmuDev, float *distDev)
{
    if (ctx->world > 1)
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "batch-map trainer needs an unsharded context");
    if (n == 0)
        return VSOM_OK;
    int rc;
    // ---- phase A
    if (isFirst)
    {
        if (score_tc_supported(ctx) && n >= 1024)
            rc = launch_find_bmu_tc(ctx, xDev, n, 0, bmuDev, distDev, nullptr);
        else
            rc = launch_find_bmu(ctx, xDev, n, 0, bmuDev, distDev);
        if (rc)
            return rc;
    }
    else
    {
        const unsigned grid = static_cast<unsigned>((n + 7) / 8);
        if (ctx->transform == VSOM_CLR)
            local_bmu_rows_kernel<VSOM_CLR><<<grid, 256, 0, ctx->stream>>>(xDev, n, ctx->mean, ctx->W, ctx->H, ctx->Din, ctx->Dr, ctx->P, ctx->rowStride,
                                                                           ctx->pairI, ctx->pairJ, ctx->order, lastDev, bmuDev, distDev);
        else
            local_bmu_rows_kernel<VSOM_STANDARD><<<grid, 256, 0, ctx->stream>>>(xDev, n, ctx->mean, ctx->W, ctx->H, ctx->Din, ctx->Dr, ctx->P,
                                                                                ctx->rowStride, ctx->pairI, ctx->pairJ, ctx->order, lastDev, bmuDev, distDev);
        ctx->launches += 1;
    }
    hits_add_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(bmuDev, n, ctx->hits);
    // ---- phase B: neighbourhood table over the reference's (quirky) coordinates, (float)nw
    int mx, my;
    quirky_xy(static_cast<unsigned>(ctx->N - 1), ctx->W, ctx->H, mx, my);
    const int lw = ctx->W, lh = my + 1;
    std::vector<float> lutHost(static_cast<size_t>(lw) * lh);
    for (int dy = 0; dy < lh; ++dy)
        for (int dx = 0; dx < lw; ++dx)
        {
            double nw;
            if (sigma > 1.0)
            {
                const double ddx = dx, ddy = dy;
                nw = std::exp(-(ddx * ddx / 2.0 / sigma / sigma + ddy * ddy / 2.0 / sigma / sigma));
            }
            else
                nw = (dx == 0 && dy == 0) ? 1.0 : 0.0;
            lutHost[static_cast<size_t>(dy) * lw + dx] = static_cast<float>(nw);
        }
    rc = stage_reserve(ctx, 6, sizeof(float) * lutHost.size() + 256);
    if (rc)
        return rc;
    rc = stage_reserve(ctx, 7, sizeof(int2) * n);
    if (rc)
        return rc;
    float *lutDev = static_cast<float *>(ctx->stage[6]);
    int2 *xyDev = static_cast<int2 *>(ctx->stage[7]);
    VSOM_CUDA(ctx, cudaMemcpyAsync(lutDev, lutHost.data(), sizeof(float) * lutHost.size(), cudaMemcpyHostToDevice, ctx->stream));
    VSOM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // lutHost is a local
    bmu_xy_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(bmuDev, n, ctx->W, ctx->H, xyDev);
    const int Dq = ctx->transform == VSOM_CLR ? ctx->P : ctx->Dm;
    const int cpt = ctx->transform == VSOM_CLR ? BM_CPT / 2 : BM_CPT;
    if (Dq > BM_THREADS * cpt)
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "batch-map trainer: model vector longer than 4096 components");
    const int nbPerCta = std::max(1, std::min(BM_MAXNB, BM_THREADS * cpt / Dq));
    const unsigned grid = static_cast<unsigned>((ctx->N + nbPerCta - 1) / nbPerCta);
    const size_t smem = sizeof(float) * (static_cast<size_t>(BM_ROWS) * ctx->Din + 2 * BM_MAXNB * BM_ROWS) + sizeof(int2) * BM_ROWS;
#define VSOM_BM_LAUNCH(TR)                                                                                                                       \
    {                                                                                                                                            \
        VSOM_CUDA(ctx, cudaFuncSetAttribute(batch_update_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));       \
        batch_update_kernel<TR><<<grid, BM_THREADS, smem, ctx->stream>>>(xDev, n, ctx->Din, Dq, ctx->P, ctx->W, ctx->H, ctx->N, nbPerCta, xyDev, \
                                                                         lutDev, lw, ctx->pairI, ctx->pairJ, ctx->mean, ctx->sigma, ctx->weight, \
                                                                         ctx->rowStride);                                                        \
    }
    if (ctx->transform == VSOM_STANDARD)
        VSOM_BM_LAUNCH(VSOM_STANDARD)
    else if (ctx->transform == VSOM_MEDIAN)
        VSOM_BM_LAUNCH(VSOM_MEDIAN)
    else
        VSOM_BM_LAUNCH(VSOM_CLR)
#undef VSOM_BM_LAUNCH
    ctx->launches += 3;
    VSOM_CUDA(ctx, cudaGetLastError());
    return VSOM_OK;
}

} // namespace vsom

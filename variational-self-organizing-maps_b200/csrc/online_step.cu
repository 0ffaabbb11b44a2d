// online_step.cu — K1: the online VSOM training step as ONE persistent sm_100a kernel per sample chunk.
//
// Replaces, per sample and in sample order: Som::trainSingle (src/Som.cpp:885-947) = Som::findBmu (:291-309)
// over Som::euclidianWeightedDist (:124-141) + the Gaussian-neighbourhood update of map / weightMap / SMap /
// sigmaMap (:899-944, Som::calculateNeighbourhoodWeight :949-975) + Som::addBmu (:1189-1192).
//
// Design
//   * grid = one CTA per SM, launched cooperatively so that all CTAs are co-resident.  Grid nodes are dealt
//     round-robin: CTA b owns shard-local nodes b, b+G, b+2G, ...  Ownership never changes, so a node's rows
//     are read and written by exactly one CTA: when the owned rows of the three planes fit in shared memory
//     they are loaded once, stay there for the whole chunk and are written back at the end (configs 1-3);
//     otherwise they are addressed in global memory (large maps).
//   * per sample there is exactly ONE grid-wide exchange: every CTA publishes the min (distance,node) key of
//     its nodes in a tagged 64-bit slot, then reads all G slots and takes the min itself.  No second barrier
//     is needed because the window update of a node is done by its owner, which is also the only CTA that
//     scans it for the next sample.  Round-robin ownership spreads any update window evenly over the SMs.
//   * reduction order is a template parameter (vsom_reduction_order): REFERENCE sums the squared residuals
//     sequentially like the reference's dot product, so BMUs and every plane stay bit-identical to it.
//   * the distance of the sample to its UPDATED BMU (trainSingle's return value, :946) is computed by an
//     otherwise idle thread / warp of the owner CTA during the next sample's scan, off the critical path.
#include "common.cuh"

#include <cmath>

namespace vsom
{

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <int TR>
__device__ __forceinline__ float stepper_scalar(float xv, float m)
{
    // Transformation::Stepper for Standard (src/Transformation.cpp:11-12) and Median (:49-50).
    const float d = __fsub_rn(xv, m);
    if (TR == VSOM_MEDIAN)
        return (d != d) ? d : static_cast<float>((0.0f < d) - (d < 0.0f));
    return d;
}

template <int TR, int ORDER>
__global__ void __launch_bounds__(kThreads, 1) online_step_kernel(const StepParams p)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ u64 sWarpKey[kWarps];
    __shared__ u64 sBmuKey;
    __shared__ int sCnt;
    __shared__ int sAbort;
    __shared__ int sPendL;  // local node whose post-update distance is still owed (-1: none)
    __shared__ u64 sPendT;  // ... for this sample

    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Lmax = (p.nodeCount + G - 1) / G;
    const int L = (p.nodeCount > b) ? (p.nodeCount - b + G - 1) / G : 0;
    const int DinPad = (p.Din + 3) & ~3;

    // ---- carve shared memory
    float *xs = reinterpret_cast<float *>(smemRaw);       // [3][DinPad] sample ring
    float *wloc = xs + 3 * DinPad;                        // [Lmax] weightMap of the owned nodes
    float *coefC = wloc + Lmax;                           // [Lmax] per window node: step coefficient
    float *coefN = coefC + Lmax;                          // [Lmax] (float)nw
    float *coefT = coefN + Lmax;                          // [Lmax] (float)tempWeight
    int *list = reinterpret_cast<int *>(coefT + Lmax);    // [Lmax] owned nodes inside the window
    unsigned short *pi = reinterpret_cast<unsigned short *>(list + Lmax); // [P] CLR pair tables
    unsigned short *pj = pi + ((p.P + 1) & ~1);
    float *planes = reinterpret_cast<float *>(pj + ((p.P + 1) & ~1));     // resident rows: 3 x Lmax x smStride
    // (P rounded to even keeps `planes` 4-byte aligned)

    float *mBase, *sBase, *gBase;
    size_t stride;
    if (p.resident)
    {
        stride = static_cast<size_t>(p.smStride);
        mBase = planes;
        sBase = planes + static_cast<size_t>(Lmax) * stride;
        gBase = sBase + static_cast<size_t>(Lmax) * stride;
    }
    else
    {
        stride = static_cast<size_t>(G) * p.rowStride;
        mBase = p.mean + static_cast<size_t>(b) * p.rowStride;
        sBase = p.S + static_cast<size_t>(b) * p.rowStride;
        gBase = p.sigma + static_cast<size_t>(b) * p.rowStride;
    }

    // ---- prologue: pair tables, owned weights, resident rows, first sample
    if (TR == VSOM_CLR)
        for (int q = tid; q < p.P; q += kThreads)
        {
            pi[q] = p.pairI[q];
            pj[q] = p.pairJ[q];
        }
    for (int l = tid; l < L; l += kThreads)
        wloc[l] = p.weight[static_cast<size_t>(l) * G + b];
    if (p.resident)
        for (int l = warp; l < L; l += kWarps)
        {
            const size_t g = (static_cast<size_t>(l) * G + b) * p.rowStride;
            for (int k = lane; k < p.Dm; k += 32)
            {
                mBase[l * stride + k] = p.mean[g + k];
                sBase[l * stride + k] = p.S[g + k];
                gBase[l * stride + k] = p.sigma[g + k];
            }
        }
    if (tid == 0)
    {
        sAbort = 0;
        sPendL = -1;
        sPendT = 0;
    }
    if (p.n > 0)
        for (int k = tid; k < p.Din; k += kThreads)
            cp_async4(xs + k, p.x + k);

    const double dW = static_cast<double>(p.W), dH = static_cast<double>(p.H);

    for (u64 t = 0; t < p.n; ++t)
    {
        const float *xt = xs + (t % 3) * DinPad;
        cp_async_wait_all();
        __syncthreads(); // sample t landed; update of sample t-1 is complete; sPend* of t-1 visible
        if (sAbort)
            break;
        if (t + 1 < p.n)
        {
            float *xn = xs + ((t + 1) % 3) * DinPad;
            const float *src = p.x + (t + 1) * static_cast<u64>(p.Din);
            for (int k = tid; k < p.Din; k += kThreads)
                cp_async4(xn + k, src + k);
        }
        const unsigned tag = static_cast<unsigned>((t >> 1) & 0xff);

        // ---- owed output of sample t-1: distance to its updated BMU (src/Som.cpp:946) + addBmu (:1189-1192)
        const int pendL = sPendL;
        const u64 pendT = sPendT;
        const float *xprev = xs + ((t + 2) % 3) * DinPad; // == (t-1) % 3

        // ---- scan: distance of sample t to every owned node, min key per thread
        u64 best = ~0ull;
        if (ORDER == VSOM_ORDER_REFERENCE)
        {
            if (tid == kThreads - 1 && pendL >= 0)
            {
                const float d = dist_sequential<TR>(mBase + pendL * stride, xprev, p.Dr, p.P, pi, pj);
                const size_t q = static_cast<size_t>(pendL) * G + b;
                if (p.outBmu)
                    p.outBmu[pendT] = static_cast<unsigned>(p.node0 + q);
                if (p.outDist)
                    p.outDist[pendT] = d;
                p.hits[q] += 1;
            }
            for (int l = tid; l < L; l += kThreads)
            {
                const float d = dist_sequential<TR>(mBase + l * stride, xt, p.Dr, p.P, pi, pj);
                best = u64_min(best, make_key(d, static_cast<unsigned>(p.node0 + l * G + b), tag));
            }
        }
        else
        {
            if (warp == kWarps - 1 && pendL >= 0)
            {
                const float d = dist_lanes<TR>(mBase + pendL * stride, xprev, p.Dr, p.P, pi, pj, lane);
                if (lane == 0)
                {
                    const size_t q = static_cast<size_t>(pendL) * G + b;
                    if (p.outBmu)
                        p.outBmu[pendT] = static_cast<unsigned>(p.node0 + q);
                    if (p.outDist)
                        p.outDist[pendT] = d;
                    p.hits[q] += 1;
                }
            }
            for (int l = warp; l < L; l += kWarps)
            {
                const float d = dist_lanes<TR>(mBase + l * stride, xt, p.Dr, p.P, pi, pj, lane);
                best = u64_min(best, make_key(d, static_cast<unsigned>(p.node0 + l * G + b), tag));
            }
        }
        best = warp_min_u64(best);
        if (lane == 0)
            sWarpKey[warp] = best;
        __syncthreads();

        // ---- grid-wide min-loc: publish, then read every CTA's slot of this step
        if (warp == 0)
        {
            u64 k = lane < kWarps ? sWarpKey[lane] : ~0ull;
            k = warp_min_u64(k);
            u64 *slots = p.slots + static_cast<size_t>(t & 1) * G;
            if (k == ~0ull) // CTA without nodes (cannot happen: G <= nodeCount) — still publish a tagged key
                k = (~0ull << 8) | tag;
            if (lane == 0)
                st_relaxed_gpu(slots + b, k);
            const long long t0 = clock64();
            u64 m;
            bool abort = false;
            for (;;)
            {
                m = ~0ull;
                int ok = 1;
                for (int i = lane; i < G; i += 32)
                {
                    const u64 v = ld_relaxed_gpu(slots + i);
                    ok &= (static_cast<unsigned>(v & 0xff) == tag);
                    m = u64_min(m, v);
                }
                if (__all_sync(0xffffffffu, ok))
                    break;
                if (__any_sync(0xffffffffu, clock64() - t0 > p.timeoutCycles))
                {
                    abort = true;
                    break;
                }
            }
            m = warp_min_u64(m);
            if (lane == 0)
            {
                sBmuKey = m;
                sCnt = 0;
                sPendL = -1;
                if (abort)
                {
                    sAbort = 1;
                    *p.err = 1;
                }
            }
        }
        __syncthreads();
        if (sAbort)
            break;

        // ---- window of the update (src/Som.cpp:899-903): [startX,endX) x [startY,endY), asymmetric
        const unsigned bmu = key_node(sBmuKey);
        const int bx = static_cast<int>(bmu % static_cast<unsigned>(p.W));
        const int by = static_cast<int>(bmu / static_cast<unsigned>(p.W));
        double lo = __dsub_rn(static_cast<double>(bx), p.radius);
        const int startX = static_cast<int>(static_cast<u64>(lo > 0. ? lo : 0.));
        lo = __dsub_rn(static_cast<double>(by), p.radius);
        const int startY = static_cast<int>(static_cast<u64>(lo > 0. ? lo : 0.));
        double hi = __dadd_rn(static_cast<double>(bx), p.radius);
        const int endX = static_cast<int>(static_cast<u64>(hi < dW ? hi : dW));
        hi = __dadd_rn(static_cast<double>(by), p.radius);
        const int endY = static_cast<int>(static_cast<u64>(hi < dH ? hi : dH));

        // ---- phase A: one thread per owned node inside the window: weightMap + the node's three coefficients
        for (int l = tid; l < L; l += kThreads)
        {
            const unsigned node = static_cast<unsigned>(p.node0 + l * G + b);
            const int x = static_cast<int>(node % static_cast<unsigned>(p.W));
            const int y = static_cast<int>(node / static_cast<unsigned>(p.W));
            if (x >= startX && x < endX && y >= startY && y < endY)
            {
                const int dx = x > bx ? x - bx : bx - x, dy = y > by ? y - by : by - y;
                const LutEntry e = p.lut[dy * p.lutW + dx];
                float w = wloc[l];
                float c;
                if (p.decay == VSOM_EXPONENTIAL)
                {
                    w = __fadd_rn(w, e.cexp); // :924
                    c = e.cexp;               // :925
                }
                else
                {
                    w = __fadd_rn(w, e.nwf);                                                // :930
                    const double tw = (w == 0.0f) ? 1.0 : __ddiv_rn(e.nw, static_cast<double>(w)); // :933
                    c = static_cast<float>(tw);                                              // :935
                }
                wloc[l] = w;
                const double tempWeight = (w == 0.0f) ? 0.000001 : static_cast<double>(w); // :939
                const int slot = atomicAdd(&sCnt, 1);
                list[slot] = l;
                coefC[slot] = c;
                coefN[slot] = e.nwf;
                coefT[slot] = static_cast<float>(tempWeight);
                if (node == bmu)
                {
                    sPendL = l;
                    sPendT = t;
                }
            }
        }
        __syncthreads();

        // ---- phase B: one warp per window node, lanes over the model vector (src/Som.cpp:912-942)
        const int cnt = sCnt;
        for (int e = warp; e < cnt; e += kWarps)
        {
            const int l = list[e];
            const float c = coefC[e], nwf = coefN[e], twf = coefT[e];
            float *m = mBase + l * stride, *S = sBase + l * stride, *sg = gBase + l * stride;
            if (TR != VSOM_CLR)
            {
                for (int k = lane; k < p.Dm; k += 32)
                {
                    const float xv = xt[k];
                    const float m0 = m[k];
                    const float d0 = stepper_scalar<TR>(xv, m0);               // :912 (and :935, same value)
                    const float m1 = __fadd_rn(m0, __fmul_rn(c, d0));          // :925 / :935
                    const float d1 = stepper_scalar<TR>(xv, m1);               // :941 Stepper on the new mean
                    const float s1 = __fadd_rn(S[k], __fmul_rn(nwf, __fmul_rn(d0, d1)));
                    m[k] = m1;
                    S[k] = s1;
                    sg[k] = __fsqrt_rn(fabsf(__fdiv_rn(s1, twf)));             // :942
                }
            }
            else
            {
                // Stepper of CLR (src/Transformation.cpp:107-142): inner = (A x' + B) - y';
                // delta = [ (-2 inner) x' || -2 inner ]
                const int P = p.P;
                for (int q = lane; q < P; q += 32)
                {
                    const float xi = xt[pi[q]], xj = xt[pj[q]];
                    const float a0 = m[q], b0 = m[P + q];
                    const float in0 = __fsub_rn(__fadd_rn(__fmul_rn(a0, xi), b0), xj);
                    const float db0 = __fmul_rn(-2.0f, in0);
                    const float da0 = __fmul_rn(db0, xi);
                    const float a1 = __fadd_rn(a0, __fmul_rn(c, da0));
                    const float b1 = __fadd_rn(b0, __fmul_rn(c, db0));
                    const float in1 = __fsub_rn(__fadd_rn(__fmul_rn(a1, xi), b1), xj);
                    const float db1 = __fmul_rn(-2.0f, in1);
                    const float da1 = __fmul_rn(db1, xi);
                    const float sa = __fadd_rn(S[q], __fmul_rn(nwf, __fmul_rn(da0, da1)));
                    const float sb = __fadd_rn(S[P + q], __fmul_rn(nwf, __fmul_rn(db0, db1)));
                    m[q] = a1;
                    m[P + q] = b1;
                    S[q] = sa;
                    S[P + q] = sb;
                    sg[q] = __fsqrt_rn(fabsf(__fdiv_rn(sa, twf)));
                    sg[P + q] = __fsqrt_rn(fabsf(__fdiv_rn(sb, twf)));
                }
            }
        }
        // the __syncthreads at the top of the next iteration orders these writes before the next scan
    }

    // ---- epilogue: last owed output, then write the owned rows back
    __syncthreads();
    if (sPendL >= 0)
    {
        const int pendL = sPendL;
        const u64 pendT = sPendT;
        const float *xprev = xs + (pendT % 3) * DinPad;
        float d = 0.0f;
        bool writer = false;
        if (ORDER == VSOM_ORDER_REFERENCE)
        {
            if (tid == 0)
            {
                d = dist_sequential<TR>(mBase + pendL * stride, xprev, p.Dr, p.P, pi, pj);
                writer = true;
            }
        }
        else if (warp == 0)
        {
            d = dist_lanes<TR>(mBase + pendL * stride, xprev, p.Dr, p.P, pi, pj, lane);
            writer = lane == 0;
        }
        if (writer)
        {
            const size_t q = static_cast<size_t>(pendL) * G + b;
            if (p.outBmu)
                p.outBmu[pendT] = static_cast<unsigned>(p.node0 + q);
            if (p.outDist)
                p.outDist[pendT] = d;
            p.hits[q] += 1;
        }
    }
    for (int l = tid; l < L; l += kThreads)
        p.weight[static_cast<size_t>(l) * G + b] = wloc[l];
    if (p.resident)
        for (int l = warp; l < L; l += kWarps)
        {
            const size_t g = (static_cast<size_t>(l) * G + b) * p.rowStride;
            for (int k = lane; k < p.Dm; k += 32)
            {
                p.mean[g + k] = mBase[l * stride + k];
                p.S[g + k] = sBase[l * stride + k];
                p.sigma[g + k] = gBase[l * stride + k];
            }
        }
}

// --------------------------------------------------------------------------------------------- host side

static size_t online_step_smem(const vsom_ctx *ctx, int G, bool resident, int smStride)
{
    const int Lmax = (ctx->N + G - 1) / G;
    const int DinPad = (ctx->Din + 3) & ~3;
    size_t bytes = sizeof(float) * (3 * static_cast<size_t>(DinPad) + 4 * static_cast<size_t>(Lmax)) + sizeof(int) * static_cast<size_t>(Lmax);
    bytes += 2 * sizeof(unsigned short) * static_cast<size_t>((ctx->P + 1) & ~1);
    if (resident)
        bytes += sizeof(float) * 3 * static_cast<size_t>(Lmax) * smStride;
    return bytes;
}

typedef void (*StepKernel)(const StepParams);
static StepKernel pick_kernel(int transform, int order)
{
    static const StepKernel table[3][2] = {
        {online_step_kernel<VSOM_STANDARD, VSOM_ORDER_REFERENCE>, online_step_kernel<VSOM_STANDARD, VSOM_ORDER_LANES>},
        {online_step_kernel<VSOM_MEDIAN, VSOM_ORDER_REFERENCE>, online_step_kernel<VSOM_MEDIAN, VSOM_ORDER_LANES>},
        {online_step_kernel<VSOM_CLR, VSOM_ORDER_REFERENCE>, online_step_kernel<VSOM_CLR, VSOM_ORDER_LANES>}};
    return table[transform][order];
}

int configure_online_step(vsom_ctx *ctx)
{
    StepKernel k = pick_kernel(ctx->transform, ctx->order);
    const int G = ctx->N < ctx->numSMs ? ctx->N : ctx->numSMs;
    // an odd row stride makes the thread-per-node sequential scan bank-conflict free; the warp-per-node
    // paths read consecutive words and do not care.
    const int smStride = ctx->Dm | 1;
    const size_t statics = 512; // static __shared__ of the kernel, rounded up
    size_t bytes = online_step_smem(ctx, G, true, smStride);
    int resident = 1;
    if (bytes + statics > static_cast<size_t>(ctx->smemOptin))
    {
        resident = 0;
        bytes = online_step_smem(ctx, G, false, smStride);
        if (bytes + statics > static_cast<size_t>(ctx->smemOptin))
            return set_error(ctx, VSOM_ERR_UNSUPPORTED, "online step: per-CTA bookkeeping does not fit in shared memory for this map");
    }
    VSOM_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    int perSm = 0;
    VSOM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, kThreads, bytes));
    if (perSm < 1)
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "online step: kernel does not fit on an SM");
    ctx->gridTrain = G;
    ctx->residentTrain = resident;
    ctx->smStrideTrain = smStride;
    ctx->smemTrain = bytes;
    return VSOM_OK;
}

// Neighbourhood table for one (eta, sigma): Som::calculateNeighbourhoodWeight (src/Som.cpp:949-975) with the
// division order of :962-963, evaluated with the host libm like the reference does.
static int build_lut(vsom_ctx *ctx, double eta, double sigma)
{
    if (ctx->lutEta == eta && ctx->lutSigma == sigma && ctx->lut)
        return VSOM_OK;
    const double radius = 2.5 * sigma;
    const int reach = static_cast<int>(std::ceil(radius));
    const int lw = (ctx->W - 1 < reach ? ctx->W - 1 : reach) + 1;
    const int lh = (ctx->H - 1 < reach ? ctx->H - 1 : reach) + 1;
    ctx->lutHost.resize(static_cast<size_t>(lw) * lh);
    for (int dy = 0; dy < lh; ++dy)
        for (int dx = 0; dx < lw; ++dx)
        {
            double nw;
            if (sigma > 1.0)
            {
                const double ddx = static_cast<double>(dx), ddy = static_cast<double>(dy);
                nw = std::exp(-(ddx * ddx / 2.0 / sigma / sigma + ddy * ddy / 2.0 / sigma / sigma));
            }
            else
                nw = (dx == 0 && dy == 0) ? 1.0 : 0.0;
            LutEntry e;
            e.nw = nw;
            e.cexp = static_cast<float>(nw * eta);
            e.nwf = static_cast<float>(nw);
            ctx->lutHost[static_cast<size_t>(dy) * lw + dx] = e;
        }
    const size_t bytes = ctx->lutHost.size() * sizeof(LutEntry);
    if (bytes > ctx->lutCap)
    {
        if (ctx->lut)
            VSOM_CUDA(ctx, cudaFree(ctx->lut));
        ctx->lut = nullptr;
        VSOM_CUDA(ctx, cudaMalloc(&ctx->lut, bytes));
        ctx->lutCap = bytes;
    }
    // the previous chunk's kernel may still be reading the old table: stream-ordered copy from a pageable
    // buffer is staged by the runtime before the call returns, so lutHost can be reused afterwards.
    VSOM_CUDA(ctx, cudaMemcpyAsync(ctx->lut, ctx->lutHost.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->lutEta = eta;
    ctx->lutSigma = sigma;
    ctx->lutW = lw;
    ctx->lutH = lh;
    return VSOM_OK;
}

int launch_online_step(vsom_ctx *ctx, const float *xDev, size_t n, double eta, double sigma, int decay, unsigned *outBmuDev, float *outDistDev)
{
    if (!(sigma > 1.0))
        return set_error(ctx, VSOM_ERR_UNSUPPORTED,
                         "online step: sigma <= 1 selects the reference's findLocalBmu regime (src/Som.cpp:335-454), not built on the device yet");
    if (decay != VSOM_EXPONENTIAL && decay != VSOM_INVERSE_PROPORTIONAL)
        return set_error(ctx, VSOM_ERR_INVALID, "online step: decay must be Exponential or InverseProportional");
    if (n == 0)
        return VSOM_OK;
    if (ctx->gridTrain == 0)
    {
        int rc = configure_online_step(ctx);
        if (rc)
            return rc;
    }
    int rc = build_lut(ctx, eta, sigma);
    if (rc)
        return rc;
    const int G = ctx->gridTrain;
    VSOM_CUDA(ctx, cudaMemsetAsync(ctx->slots, 0xff, sizeof(u64) * 2 * static_cast<size_t>(G), ctx->stream));
    VSOM_CUDA(ctx, cudaMemsetAsync(ctx->errFlag, 0, sizeof(int), ctx->stream));

    StepParams p;
    p.W = ctx->W;
    p.H = ctx->H;
    p.node0 = 0;
    p.nodeCount = ctx->N;
    p.Din = ctx->Din;
    p.Dm = ctx->Dm;
    p.Dr = ctx->Dr;
    p.P = ctx->P;
    p.rowStride = ctx->rowStride;
    p.mean = ctx->mean;
    p.S = ctx->S;
    p.sigma = ctx->sigma;
    p.weight = ctx->weight;
    p.hits = ctx->hits;
    p.x = xDev;
    p.n = n;
    p.decay = decay;
    p.radius = 2.5 * sigma;
    p.lut = ctx->lut;
    p.lutW = ctx->lutW;
    p.pairI = ctx->pairI;
    p.pairJ = ctx->pairJ;
    p.slots = ctx->slots;
    p.err = ctx->errFlag;
    p.outBmu = outBmuDev;
    p.outDist = outDistDev;
    p.resident = ctx->residentTrain;
    p.smStride = ctx->smStrideTrain;
    p.timeoutCycles = 4000000000ll; // ~2 s at 1.9 GHz: a peer CTA that never publishes is a bug, not a wait

    StepKernel k = pick_kernel(ctx->transform, ctx->order);
    void *args[] = {&p};
    VSOM_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k), dim3(G), dim3(kThreads), args, ctx->smemTrain, ctx->stream));
    ctx->launches += 1;
    return VSOM_OK;
}

} // namespace vsom

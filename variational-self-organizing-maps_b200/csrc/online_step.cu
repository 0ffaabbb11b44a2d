// online_step.cu — K1: the online VSOM training step as ONE persistent sm_100a kernel per sample chunk.
//
// Replaces, per sample and in sample order: Som::trainSingle (src/Som.cpp:885-947) = Som::findBmu (:291-309)
// over Som::euclidianWeightedDist (:124-141) + the Gaussian-neighbourhood update of map / weightMap / SMap /
// sigmaMap (:899-944, Som::calculateNeighbourhoodWeight :949-975) + Som::addBmu (:1189-1192).
//
// Design
//   * grid = one CTA per SM, launched cooperatively so that all CTAs are co-resident.  Grid nodes are dealt
//     round-robin: CTA b owns shard-local nodes b, b+G, b+2G, ...  Ownership never changes, so a node's rows
//     are read and written by exactly one CTA: when the owned rows of the mean and S planes fit in shared
//     memory they are loaded once, stay there for the whole chunk and are written back at the end
//     (configs 1-3); otherwise they are addressed in global memory (large maps).
//   * per sample there is exactly ONE grid-wide exchange: every CTA pushes the min (distance,node) key of its
//     nodes, tagged with the step number, into a private row of every other CTA (slots[dest][src]); each CTA
//     then polls only its own row (G words, lines nobody else reads) and takes the min itself.  Pushing keeps
//     the polling traffic off shared cache lines: with a single shared row, 148 pollers hammering the same
//     ten L2 lines cost ~4000 cycles per sample (profiles/r01_bench_k1_v1.json).  No second barrier is needed
//     because the window update of a node is done by its owner, which is also the only CTA that scans it for
//     the next sample.  Round-robin ownership spreads any update window evenly over the SMs.
//   * sigmaMap is a pure function of the node's SMap and weightMap at its LAST visit
//     (sigma = sqrt(|S / (W==0 ? 1e-6 : W)|), src/Som.cpp:939-942, overwritten at every visit), and nothing on
//     this path reads it (all shipped Comparers ignore `dispersion`).  The kernel therefore only marks the
//     visited nodes and evaluates that expression once per visited node at the end of the chunk — the stored
//     plane is bit-identical, and the IEEE divide + square root leave the per-sample loop.
//   * reduction order is a template parameter (vsom_reduction_order): REFERENCE sums the squared residuals
//     sequentially like the reference's dot product (one thread per node), so BMUs and every plane stay
//     bit-identical to it; LANES uses a warp per node.
//   * the distance of the sample to its UPDATED BMU (trainSingle's return value, :946) is computed by an
//     otherwise idle thread / warp of the owner CTA during the next sample's scan, off the critical path.
#include "common.cuh"

#include <cmath>

namespace vsom
{

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxSlotsPerLane = 5;   // the exchange row of a CTA is read by one warp: at most 160 CTAs
constexpr size_t kStaticSmem = 1024; // static __shared__ of the kernel, rounded up

template <int TR>
__device__ __forceinline__ float stepper_scalar(float xv, float m)
{
    // Transformation::Stepper for Standard (src/Transformation.cpp:11-12) and Median (:49-50).
    const float d = __fsub_rn(xv, m);
    if (TR == VSOM_MEDIAN)
    {
        // sign(d) in {-1, +0, +1}; NaN stays NaN: |d| > 0 ? copysign(1, d) : |d|   (a zero gives +0 like the reference's
        // (0<d)-(d<0); NaN fails the comparison and |NaN| is NaN)
        const float a = fabsf(d);
        const float one = __uint_as_float((__float_as_uint(d) & 0x80000000u) | 0x3f800000u); // copysign(1, d)
        return a > 0.0f ? one : a;
    }
    return d;
}

// One element of the window update (src/Som.cpp:912-941) for Standard / Median; sigma is evaluated lazily.
template <int TR>
__device__ __forceinline__ void update_element(float xv, float c, float nwf, float &m, float &S)
{
    const float d0 = stepper_scalar<TR>(xv, m);                  // :912 (and :935, same value)
    const float m1 = __fadd_rn(m, __fmul_rn(c, d0));             // :925 / :935
    const float d1 = stepper_scalar<TR>(xv, m1);                 // :941 Stepper on the new mean
    S = __fadd_rn(S, __fmul_rn(nwf, __fmul_rn(d0, d1)));         // :941
    m = m1;
}

// Squared distance, reference order, over zero-padded rows with 128-bit loads (Standard / Median).
// The chain s += r*r is 4 cycles per element and cannot be shortened without changing the order; the loads of
// the next group are issued before the current group's adds so that their latency stays off the chain.
__device__ __forceinline__ void acc4(float &s, const float4 a, const float4 b)
{
    float r = __fsub_rn(a.x, b.x);
    s = __fadd_rn(s, __fmul_rn(r, r));
    r = __fsub_rn(a.y, b.y);
    s = __fadd_rn(s, __fmul_rn(r, r));
    r = __fsub_rn(a.z, b.z);
    s = __fadd_rn(s, __fmul_rn(r, r));
    r = __fsub_rn(a.w, b.w);
    s = __fadd_rn(s, __fmul_rn(r, r));
}
__device__ __forceinline__ float dist_sequential_v4(const float *m, const float *xs, int n4)
{
    const float4 *m4 = reinterpret_cast<const float4 *>(m);
    const float4 *x4 = reinterpret_cast<const float4 *>(xs);
    float s = 0.0f;
    const int pairs = n4 >> 1;
    if (pairs > 0)
    {
        float4 a0 = m4[0], b0 = x4[0], a1 = m4[1], b1 = x4[1];
        for (int g = 1; g < pairs; ++g)
        {
            const float4 c0 = m4[2 * g], d0 = x4[2 * g], c1 = m4[2 * g + 1], d1 = x4[2 * g + 1];
            acc4(s, a0, b0);
            acc4(s, a1, b1);
            a0 = c0;
            b0 = d0;
            a1 = c1;
            b1 = d1;
        }
        acc4(s, a0, b0);
        acc4(s, a1, b1);
    }
    if (n4 & 1)
        acc4(s, m4[n4 - 1], x4[n4 - 1]);
    return s;
}

template <int TR>
__device__ __forceinline__ float node_dist_reference(const float *m, const float *xs, const StepParams &p, int n4, const unsigned short *pi,
                                                     const unsigned short *pj)
{
    if (TR != VSOM_CLR)
        return dist_sequential_v4(m, xs, n4); // the zero padding adds +0 terms: s + 0 == s exactly
    return dist_sequential<TR>(m, xs, p.Dr, p.P, pi, pj);
}

// Som::findLocalBmu (src/Som.cpp:335-454) on the per-step distance array: greedy walk from `start` over the
// 8-neighbourhood, then three cells one step further in the X direction of travel, until the best node stops moving.
// The reference's size_t arithmetic is kept: the "-1" offsets are 2^64-1, min(x + off, W-1) therefore wraps the left / up
// neighbour of column / row 0 to the last column / row, and its Y-direction continuation loop never runs (it starts at
// size_t(-1)).  Distances are compared as the reference does (strict '<', first candidate in its order wins).
__device__ __forceinline__ unsigned local_bmu_walk(const float *dist, u64 W, u64 H, u64 start)
{
    const u64 M1 = ~0ull;
    const u64 fx[8] = {M1, 0, 1, 1, 1, 0, M1, M1};
    const u64 fy[8] = {1, 1, 1, 0, M1, M1, M1, 0};
    u64 lastBMU = start, minIndex = start, lastMeasured = start;
    float minDist = ld_relaxed_gpu_f32(dist + start);
    for (;;)
    {
        const u64 lmX = lastMeasured % W, lmY = lastMeasured / W, lbX = lastBMU % W;
        if (lastMeasured == lastBMU)
        {
            u64 idx[8];
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const u64 cx = lmX + fx[i] < W - 1 ? lmX + fx[i] : W - 1; // min(x + off, W-1) in size_t
                const u64 cy = lmY + fy[i] < H - 1 ? lmY + fy[i] : H - 1;
                idx[i] = cy * W + cx;
                v[i] = ld_relaxed_gpu_f32(dist + idx[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (v[i] < minDist)
                {
                    minDist = v[i];
                    minIndex = idx[i];
                }
            if (minIndex == lastBMU)
                return static_cast<unsigned>(minIndex);
            lastMeasured = minIndex;
        }
        else
        {
            if (lmX - lbX) // moving in X
            {
                u64 idx[3];
                float v[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                {
                    const u64 ox = lmX + lmX - lbX, oy = lmY + static_cast<u64>(static_cast<long long>(i - 1));
                    const u64 cx = ox < W - 1 ? ox : W - 1, cy = oy < H - 1 ? oy : H - 1;
                    idx[i] = cy * W + cx;
                    v[i] = ld_relaxed_gpu_f32(dist + idx[i]);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (v[i] < minDist)
                    {
                        minDist = v[i];
                        minIndex = idx[i];
                    }
            }
            if (minIndex == lastMeasured)
                return static_cast<unsigned>(minIndex);
            lastBMU = lastMeasured;
            lastMeasured = minIndex;
        }
    }
}

template <int TR, int ORDER, bool RES>
__global__ void __launch_bounds__(kThreads, 1) online_step_kernel(const StepParams p)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ u64 sWarpKey[kWarps];
    __shared__ int sWin[8]; // bmu, bx, by, startX, endX, startY, endY of the current sample
    __shared__ int sAbort;
    __shared__ int sPendL; // local node whose post-update distance is still owed (-1: none)
    __shared__ u64 sPendT; // ... for this sample
    __shared__ long long sClk[6], sProf[5]; // diagnostics (thread 0 only)

    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Lmax = (p.nodeCount + G - 1) / G;
    const int L = (p.nodeCount > b) ? (p.nodeCount - b + G - 1) / G : 0;
    const int DinPad = (p.Din + 3) & ~3;
    const int DmPad = (p.Dm + 3) & ~3;
    const int n4 = DmPad >> 2;
    const int Gp = (G + 15) & ~15; // pitch of one destination row of the exchange (whole 128-byte lines)

    // ---- carve shared memory (every region 16-byte aligned)
    float *xs = reinterpret_cast<float *>(smemRaw);                          // [3][DinPad] sample ring
    float *wbuf = xs + 3 * DinPad;                                           // [Lpad] weightMap of the owned nodes
    const int Lpad = (Lmax + 3) & ~3;
    float2 *coef = reinterpret_cast<float2 *>(wbuf + Lpad);                  // [Lpad] per owned node in the window: {step coefficient, (float)nw}
    int2 *nodeXY = reinterpret_cast<int2 *>(coef + Lpad);                    // [Lpad] grid position of the owned nodes
    unsigned *touched = reinterpret_cast<unsigned *>(nodeXY + Lpad);         // [Lpad] visited in this chunk
    unsigned *inWin = touched + Lpad;                                        // [Lpad] inside the window of the current sample
    unsigned short *pi = reinterpret_cast<unsigned short *>(inWin + Lpad);   // [Ppad] CLR pair tables
    const int Ppad = (p.P + 7) & ~7;
    unsigned short *pj = pi + Ppad;
    float *planes = reinterpret_cast<float *>(pj + Ppad);                    // resident rows: 2 x Lmax x smStride
    LutEntry *lutS = reinterpret_cast<LutEntry *>(planes + (RES ? 2 * static_cast<size_t>(Lmax) * p.smStride : 0)); // optional copy of the table

    float *mBase, *sBase;
    size_t stride;
    if (RES)
    {
        stride = static_cast<size_t>(p.smStride);
        mBase = planes;
        sBase = planes + static_cast<size_t>(Lmax) * stride;
    }
    else
    {
        stride = static_cast<size_t>(G) * p.rowStride;
        mBase = p.mean + static_cast<size_t>(b) * p.rowStride;
        sBase = p.S + static_cast<size_t>(b) * p.rowStride;
    }

    // slices of 128 elements of the model vector (CLR: of the pair index): one warp per (window node, slice)
    const int nq = TR == VSOM_CLR ? p.P : DmPad;
    const int nCh = (nq + 127) >> 7;

    // ---- prologue: pair tables, neighbourhood table, owned weights / positions, resident rows, first sample
    if (TR == VSOM_CLR)
        for (int q = tid; q < p.P; q += kThreads)
        {
            pi[q] = p.pairI[q];
            pj[q] = p.pairJ[q];
        }
    if (p.lutSmem)
        for (int i = tid; i < p.lutCount; i += kThreads)
            lutS[i] = p.lut[i];
    for (int l = tid; l < L; l += kThreads)
    {
        const unsigned node = static_cast<unsigned>(p.node0 + l * G + b);
        wbuf[l] = p.weight[static_cast<size_t>(l) * G + b];
        nodeXY[l] = make_int2(static_cast<int>(node % static_cast<unsigned>(p.W)), static_cast<int>(node / static_cast<unsigned>(p.W)));
        touched[l] = 0;
    }
    for (int k = tid; k < 3 * DinPad; k += kThreads)
        xs[k] = 0.0f; // keeps the pad lanes of the 128-bit paths at zero
    if (RES)
        for (int l = warp; l < L; l += kWarps)
        {
            const size_t g = (static_cast<size_t>(l) * G + b) * p.rowStride;
            for (int k = lane; k < static_cast<int>(stride); k += 32)
            {
                const bool in = k < p.Dm;
                mBase[l * stride + k] = in ? p.mean[g + k] : 0.0f;
                sBase[l * stride + k] = in ? p.S[g + k] : 0.0f;
            }
        }
    if (tid == 0)
    {
        sAbort = 0;
        sPendL = -1;
        sPendT = 0;
    }
    __syncthreads();

    // sample prefetch: 16-byte cp.async.cg (L2 only — a streamed sample must not evict the L1-resident table)
    // when rows are 16-byte aligned, else 4-byte cp.async.ca
    auto prefetch = [&](u64 t, int slot) {
        float *dst = xs + slot * DinPad;
        const float *src = p.x + t * static_cast<u64>(p.Din);
        if (p.xVec)
        {
            for (int k = tid * 4; k < p.Din; k += kThreads * 4)
                cp_async16(dst + k, src + k);
        }
        else
        {
            for (int k = tid; k < p.Din; k += kThreads)
                cp_async4(dst + k, src + k);
        }
    };
    if (p.n > 0)
        prefetch(0, 0);

    const double dW = static_cast<double>(p.W), dH = static_cast<double>(p.H);
    if (tid == 0)
        for (int i = 0; i < 5; ++i)
            sProf[i] = 0;
    u64 done = 0;

    int ring = 0; // t % 3
    for (u64 t = 0; t < p.n; ++t, ring = ring == 2 ? 0 : ring + 1)
    {
        if (p.prof && tid == 0)
            sClk[0] = clock64();
        const float *xt = xs + ring * DinPad;
        cp_async_wait_all();
        __syncthreads(); // sample t landed; update of sample t-1 is complete; sPend* of t-1 visible
        if (t + 1 < p.n)
            prefetch(t + 1, ring == 2 ? 0 : ring + 1);
        if (p.prof && tid == 0)
            sClk[1] = clock64();
        const unsigned tag = static_cast<unsigned>((t >> 1) & 0xff);

        // ---- owed output of sample t-1: distance to its updated BMU (src/Som.cpp:946) + addBmu (:1189-1192)
        const int pendL = sPendL;
        const u64 pendT = sPendT;
        const float *xprev = xs + (ring == 0 ? 2 : ring - 1) * DinPad; // sample t-1

        // ---- scan: distance of sample t to every owned node, min key per thread
        u64 best = ~0ull;
        if (ORDER == VSOM_ORDER_REFERENCE)
        {
            if (tid == kThreads - 1 && pendL >= 0)
            {
                const float d = node_dist_reference<TR>(mBase + pendL * stride, xprev, p, n4, pi, pj);
                const size_t q = static_cast<size_t>(pendL) * G + b;
                if (p.outBmu)
                    p.outBmu[pendT] = static_cast<unsigned>(p.node0 + q);
                if (p.outDist)
                    p.outDist[pendT] = d;
                p.hits[q] += 1;
            }
            for (int l = tid; l < L; l += kThreads)
            {
                const float d = node_dist_reference<TR>(mBase + l * stride, xt, p, n4, pi, pj);
                if (p.localSearch)
                    p.distBuf[(t & 1) * static_cast<u64>(p.nodeCount) + static_cast<u64>(l) * G + b] = d;
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
                if (p.localSearch && lane == 0)
                    p.distBuf[(t & 1) * static_cast<u64>(p.nodeCount) + static_cast<u64>(l) * G + b] = d;
                best = u64_min(best, make_key(d, static_cast<unsigned>(p.node0 + l * G + b), tag));
            }
        }
        if (p.localSearch)
            __threadfence(); // the distances written above must be visible before this CTA's key is
        // REFERENCE order with at most 32 owned nodes: every key already sits in warp 0 — no CTA barrier
        const bool crossWarp = ORDER != VSOM_ORDER_REFERENCE || L > 32;
        if (crossWarp)
        {
            best = warp_min_key(best);
            if (lane == 0)
                sWarpKey[warp] = best;
            __syncthreads();
        }

        // ---- grid-wide min-loc: push the key into every CTA's row, then poll the own row
        if (warp == 0)
        {
            u64 k = crossWarp ? (lane < kWarps ? sWarpKey[lane] : ~0ull) : best;
            k = warp_min_key(k);
            if (k == ~0ull) // CTA without nodes (cannot happen: G <= nodeCount) — still publish a tagged key
                k = (~0ull << 8) | tag;
            u64 *buf = p.slots + static_cast<size_t>(t & 1) * G * Gp;
            if (p.prof && tid == 0)
                sClk[2] = clock64();
            for (int d = lane; d < G; d += 32)
                st_relaxed_gpu(buf + static_cast<size_t>(d) * Gp + b, k);
            const u64 *row = buf + static_cast<size_t>(b) * Gp;
            const long long t0 = clock64();
            u64 m;
            bool abort = false;
            const u64 filler = (~0ull << 8) | tag;
            for (;;)
            {
                // all loads of a round are issued before any is consumed: one L2 round trip per round, not G/32
                u64 v[kMaxSlotsPerLane];
#pragma unroll
                for (int j = 0; j < kMaxSlotsPerLane; ++j)
                {
                    const int i = lane + 32 * j;
                    v[j] = i < G ? ld_relaxed_gpu(row + i) : filler;
                }
                m = ~0ull;
                int ok = 1;
#pragma unroll
                for (int j = 0; j < kMaxSlotsPerLane; ++j)
                {
                    ok &= (static_cast<unsigned>(v[j] & 0xff) == tag);
                    m = u64_min(m, v[j]);
                }
                if (__all_sync(0xffffffffu, ok))
                    break;
                if (__any_sync(0xffffffffu, clock64() - t0 > p.timeoutCycles))
                {
                    abort = true;
                    break;
                }
            }
            m = warp_min_key(m);
            if (p.world > 1 && !abort)
            {
                // ---- node-sharded map: the GPU-local winner is pushed into every rank's memory over NVLink by CTA 0
                // (system-scope stores to peer-mapped pointers); every CTA then polls ITS OWN GPU's row of `world`
                // tagged keys.  Indexed by the global step, so it needs neither a reset nor a host barrier between
                // launches; ranks can never be more than one step apart.
                const u64 gs = p.stepBase + t;
                const unsigned gtag = static_cast<unsigned>((gs >> 1) & 0xff);
                const int gbuf = static_cast<int>(gs & 1);
                const u64 mine = (m & ~0xffull) | gtag;
                if (b == 0 && lane < p.world)
                    st_relaxed_sys(p.peerSlots[lane] + gbuf * p.world + p.rank, mine);
                const u64 *grow = p.rankSlots + gbuf * p.world;
                const long long g0 = clock64();
                u64 gv;
                for (;;)
                {
                    gv = lane < p.world ? ld_relaxed_sys(grow + lane) : ((~0ull << 8) | gtag);
                    if (__all_sync(0xffffffffu, static_cast<unsigned>(gv & 0xff) == gtag))
                        break;
                    if (__any_sync(0xffffffffu, clock64() - g0 > p.timeoutCycles))
                    {
                        abort = true;
                        break;
                    }
                }
                m = warp_min_key(gv);
                if (b == 0 && lane == 0 && p.outBmu && !abort)
                    p.outBmu[t] = key_node(m); // every rank records the global BMU of every sample
            }
            if (p.localSearch && !abort)
            {
                // every CTA has seen every CTA's key of this step, hence (fence above) every distance of this step
                __threadfence();
                __syncwarp();
            }
            // window of the update (src/Som.cpp:899-903): [startX,endX) x [startY,endY), asymmetric.  Lane 0 computes it
            // (f64, exactly the reference's expressions) and broadcasts it to the warp.
            int wbmu = 0, wbx = 0, wby = 0, wsx = 0, wex = 0, wsy = 0, wey = 0;
            if (lane == 0)
            {
                unsigned bmu = key_node(m);
                if (p.localSearch && !abort)
                {
                    bmu = local_bmu_walk(p.distBuf + (t & 1) * static_cast<u64>(p.nodeCount), static_cast<u64>(p.W), static_cast<u64>(p.H),
                                         p.lastIn ? p.lastIn[t] : 0ull);
                    if (b == 0 && p.outBmu)
                        p.outBmu[t] = bmu;
                }
                wbmu = static_cast<int>(bmu);
                wbx = static_cast<int>(bmu % static_cast<unsigned>(p.W));
                wby = static_cast<int>(bmu / static_cast<unsigned>(p.W));
                double lo = __dsub_rn(static_cast<double>(wbx), p.radius);
                wsx = static_cast<int>(static_cast<u64>(lo > 0. ? lo : 0.));
                lo = __dsub_rn(static_cast<double>(wby), p.radius);
                wsy = static_cast<int>(static_cast<u64>(lo > 0. ? lo : 0.));
                double hi = __dadd_rn(static_cast<double>(wbx), p.radius);
                wex = static_cast<int>(static_cast<u64>(hi < dW ? hi : dW));
                hi = __dadd_rn(static_cast<double>(wby), p.radius);
                wey = static_cast<int>(static_cast<u64>(hi < dH ? hi : dH));
                sPendL = -1;
                if (abort)
                {
                    sAbort = 1;
                    *p.err = 1;
                }
            }
            wbmu = __shfl_sync(0xffffffffu, wbmu, 0);
            wbx = __shfl_sync(0xffffffffu, wbx, 0);
            wby = __shfl_sync(0xffffffffu, wby, 0);
            wsx = __shfl_sync(0xffffffffu, wsx, 0);
            wex = __shfl_sync(0xffffffffu, wex, 0);
            wsy = __shfl_sync(0xffffffffu, wsy, 0);
            wey = __shfl_sync(0xffffffffu, wey, 0);
            if (lane == 0)
            {
                sWin[0] = wbmu;
                sWin[1] = wbx;
                sWin[2] = wby;
                sWin[3] = wsx;
                sWin[4] = wex;
                sWin[5] = wsy;
                sWin[6] = wey;
            }
            // ---- per owned node (lanes = nodes): neighbourhood entry, weightMap update and the step coefficient
            //      (src/Som.cpp:915-939).  When a node's vector spans several warps (nCh > 1) this is done once here;
            //      with one warp per node (nCh == 1) each update warp does it for its own node after the barrier,
            //      which keeps it off this warp's critical path.
            for (int l = lane; nCh > 1 && l < L; l += 32)
            {
                const int2 xy = nodeXY[l];
                float2 cf = make_float2(-1.0f, 0.0f);
                if (xy.x >= wsx && xy.x < wex && xy.y >= wsy && xy.y < wey)
                {
                    const int dx = xy.x > wbx ? xy.x - wbx : wbx - xy.x, dy = xy.y > wby ? xy.y - wby : wby - xy.y;
                    const int li = dy * p.lutW + dx;
                    float4 raw;
                    if (p.lutSmem)
                        raw = *reinterpret_cast<const float4 *>(lutS + li);
                    else
                        raw = __ldg(reinterpret_cast<const float4 *>(p.lut + li));
                    const float cexp = raw.z, nwf = raw.w;
                    float w = wbuf[l], c;
                    if (p.decay == VSOM_EXPONENTIAL)
                    {
                        w = __fadd_rn(w, cexp); // :924
                        c = cexp;               // :925
                    }
                    else
                    {
                        w = __fadd_rn(w, nwf); // :930
                        const double nw = __hiloint2double(__float_as_int(raw.y), __float_as_int(raw.x));
                        const double tw = (w == 0.0f) ? 1.0 : __ddiv_rn(nw, static_cast<double>(w)); // :933
                        c = static_cast<float>(tw);                                                    // :935
                    }
                    wbuf[l] = w;
                    touched[l] = 1;
                    cf = make_float2(c, nwf);
                    if (static_cast<unsigned>(p.node0 + l * G + b) == static_cast<unsigned>(wbmu))
                    {
                        sPendL = l;
                        sPendT = t;
                    }
                    inWin[l] = 1;
                }
                else
                    inWin[l] = 0;
                coef[l] = cf;
            }
            if (p.prof && tid == 0)
                sClk[3] = clock64();
        }
        __syncthreads();
        if (sAbort)
            break;
        if (p.prof && tid == 0)
            sClk[4] = clock64();
        // ---- update: one warp per (window node, 128-element slice) of the model vector (src/Som.cpp:912-941)
        {
            const int items = L * nCh;
            for (int it = warp; it < items; it += kWarps)
            {
                const int l = nCh > 1 ? it / nCh : it, ch = nCh > 1 ? it - l * nCh : 0;
                float c, nwf;
                if (nCh > 1)
                {
                    if (!inWin[l])
                        continue;
                    const float2 cf = coef[l];
                    c = cf.x;
                    nwf = cf.y;
                }
                else
                {
                    // one warp owns the whole node: every lane derives the coefficient (redundantly, no traffic),
                    // lane 0 records the new weight
                    const int2 xy = nodeXY[l];
                    const int bx = sWin[1], by = sWin[2];
                    if (!(xy.x >= sWin[3] && xy.x < sWin[4] && xy.y >= sWin[5] && xy.y < sWin[6]))
                        continue;
                    const int dx = xy.x > bx ? xy.x - bx : bx - xy.x, dy = xy.y > by ? xy.y - by : by - xy.y;
                    const int li = dy * p.lutW + dx;
                    float4 raw;
                    if (p.lutSmem)
                        raw = *reinterpret_cast<const float4 *>(lutS + li);
                    else
                        raw = __ldg(reinterpret_cast<const float4 *>(p.lut + li));
                    nwf = raw.w;
                    float w = wbuf[l];
                    if (p.decay == VSOM_EXPONENTIAL)
                    {
                        w = __fadd_rn(w, raw.z); // :924
                        c = raw.z;               // :925
                    }
                    else
                    {
                        w = __fadd_rn(w, nwf); // :930
                        const double nw = __hiloint2double(__float_as_int(raw.y), __float_as_int(raw.x));
                        const double tw = (w == 0.0f) ? 1.0 : __ddiv_rn(nw, static_cast<double>(w)); // :933
                        c = static_cast<float>(tw);                                                    // :935
                    }
                    __syncwarp(); // every lane has read the old weight
                    if (lane == 0)
                    {
                        wbuf[l] = w;
                        touched[l] = 1;
                        if (static_cast<unsigned>(p.node0 + l * G + b) == static_cast<unsigned>(sWin[0]))
                        {
                            sPendL = l;
                            sPendT = t;
                        }
                    }
                }
                float *m = mBase + l * stride, *S = sBase + l * stride;
                if (TR != VSOM_CLR)
                {
                    const int k = (ch << 7) + (lane << 2);
                    if (k < DmPad)
                    {
                        float4 mv = *reinterpret_cast<float4 *>(m + k);
                        float4 sv = *reinterpret_cast<float4 *>(S + k);
                        const float4 xv = *reinterpret_cast<const float4 *>(xt + k);
                        update_element<TR>(xv.x, c, nwf, mv.x, sv.x);
                        update_element<TR>(xv.y, c, nwf, mv.y, sv.y);
                        update_element<TR>(xv.z, c, nwf, mv.z, sv.z);
                        update_element<TR>(xv.w, c, nwf, mv.w, sv.w);
                        *reinterpret_cast<float4 *>(m + k) = mv;
                        *reinterpret_cast<float4 *>(S + k) = sv;
                    }
                }
                else
                {
                    // Stepper of CLR (src/Transformation.cpp:107-142): inner = (A x' + B) - y';
                    // delta = [ (-2 inner) x' || -2 inner ]
                    const int P = p.P;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const int q = (ch << 7) + (j << 5) + lane;
                        if (q < P)
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
                            m[q] = a1;
                            m[P + q] = b1;
                            S[q] = __fadd_rn(S[q], __fmul_rn(nwf, __fmul_rn(da0, da1)));
                            S[P + q] = __fadd_rn(S[P + q], __fmul_rn(nwf, __fmul_rn(db0, db1)));
                        }
                    }
                }
            }
        }
        done = t + 1;
        if (p.prof && tid == 0)
        {
            sClk[5] = clock64();
            for (int i = 0; i < 5; ++i)
                sProf[i] += sClk[i + 1] - sClk[i];
        }
        // the __syncthreads at the top of the next iteration orders these writes before the next scan
    }

    // ---- epilogue: last owed output, lazy sigma, then write the owned rows back
    __syncthreads();
    if (sPendL >= 0)
    {
        const int pendL = sPendL;
        const u64 pendT = sPendT;
        const float *xprev = xs + static_cast<int>(pendT % 3) * DinPad;
        float d = 0.0f;
        bool writer = false;
        if (ORDER == VSOM_ORDER_REFERENCE)
        {
            if (tid == 0)
            {
                d = node_dist_reference<TR>(mBase + pendL * stride, xprev, p, n4, pi, pj);
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
    const float *wfin = wbuf;
    (void)done;
    for (int l = tid; l < L; l += kThreads)
        p.weight[static_cast<size_t>(l) * G + b] = wfin[l];
    for (int l = warp; l < L; l += kWarps)
    {
        const size_t g = (static_cast<size_t>(l) * G + b) * p.rowStride;
        const bool vis = touched[l] != 0;
        // sigmaMap of a visited node: sqrt(|S / (float)(W == 0 ? 1e-6 : W)|) with its final S and W (src/Som.cpp:939-942)
        const float w = wfin[l];
        const float twf = static_cast<float>((w == 0.0f) ? 0.000001 : static_cast<double>(w));
        for (int k = lane; k < p.Dm; k += 32)
        {
            const float sv = sBase[l * stride + k];
            if (RES)
            {
                p.mean[g + k] = mBase[l * stride + k];
                p.S[g + k] = sv;
            }
            if (vis)
                p.sigma[g + k] = __fsqrt_rn(fabsf(__fdiv_rn(sv, twf)));
        }
    }
    if (p.prof && tid == 0)
        for (int i = 0; i < 5; ++i)
            p.prof[static_cast<size_t>(b) * 5 + i] = sProf[i];
}

// --------------------------------------------------------------------------------------------- host side

static int resident_stride(const vsom_ctx *ctx)
{
    // rows start 16-byte aligned and (stride / 4) is odd: the thread-per-node scan (lanes = nodes, 128-bit
    // loads) and the warp-per-node update (lanes = consecutive words) are both bank-conflict free.
    const int n4 = (ctx->Dm + 3) / 4;
    return 4 * (n4 | 1);
}

static size_t online_step_smem(const vsom_ctx *ctx, int G, bool resident, int smStride)
{
    const int Lmax = (ctx->localN + G - 1) / G;
    const int Lpad = (Lmax + 3) & ~3;
    const int DinPad = (ctx->Din + 3) & ~3;
    const int Ppad = (ctx->P + 7) & ~7;
    size_t bytes = sizeof(float) * (3 * static_cast<size_t>(DinPad) + 3 * static_cast<size_t>(Lpad));
    bytes += (sizeof(int2) + 2 * sizeof(unsigned)) * static_cast<size_t>(Lpad);
    bytes += 2 * sizeof(unsigned short) * static_cast<size_t>(Ppad);
    if (resident)
        bytes += sizeof(float) * 2 * static_cast<size_t>(Lmax) * smStride;
    return bytes;
}

typedef void (*StepKernel)(const StepParams);
static StepKernel pick_kernel(int transform, int order, int resident)
{
#define VSOM_K(TR, ORD) {online_step_kernel<TR, ORD, false>, online_step_kernel<TR, ORD, true>}
    static const StepKernel table[3][2][2] = {{VSOM_K(VSOM_STANDARD, VSOM_ORDER_REFERENCE), VSOM_K(VSOM_STANDARD, VSOM_ORDER_LANES)},
                                              {VSOM_K(VSOM_MEDIAN, VSOM_ORDER_REFERENCE), VSOM_K(VSOM_MEDIAN, VSOM_ORDER_LANES)},
                                              {VSOM_K(VSOM_CLR, VSOM_ORDER_REFERENCE), VSOM_K(VSOM_CLR, VSOM_ORDER_LANES)}};
#undef VSOM_K
    return table[transform][order][resident ? 1 : 0];
}

int configure_online_step(vsom_ctx *ctx)
{
    int G = ctx->localN < ctx->numSMs ? ctx->localN : ctx->numSMs;
    if (G > 32 * kMaxSlotsPerLane)
        G = 32 * kMaxSlotsPerLane;
    const int smStride = resident_stride(ctx);
    const size_t statics = kStaticSmem;
    size_t bytes = online_step_smem(ctx, G, true, smStride);
    int resident = 1;
    if (bytes + statics > static_cast<size_t>(ctx->smemOptin))
    {
        resident = 0;
        bytes = online_step_smem(ctx, G, false, smStride);
        if (bytes + statics > static_cast<size_t>(ctx->smemOptin))
            return set_error(ctx, VSOM_ERR_UNSUPPORTED, "online step: per-CTA bookkeeping does not fit in shared memory for this map");
    }
    StepKernel k = pick_kernel(ctx->transform, ctx->order, resident);
    VSOM_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smemOptin - static_cast<int>(kStaticSmem)));
    int perSm = 0;
    VSOM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, kThreads, bytes));
    if (perSm < 1)
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "online step: kernel does not fit on an SM");
    ctx->gridTrain = G;
    ctx->residentTrain = resident;
    ctx->smStrideTrain = smStride;
    ctx->smemTrain = bytes;
    if (!ctx->fastDisabled)
        configure_online_step_fast(ctx); // K1F for the common regime, when it fits
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
    // the previous chunk's kernel may still be reading the old table: the copy is stream-ordered behind it, and
    // a copy from pageable memory is staged by the runtime before the call returns, so lutHost can be reused.
    VSOM_CUDA(ctx, cudaMemcpyAsync(ctx->lut, ctx->lutHost.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->lutEta = eta;
    ctx->lutSigma = sigma;
    ctx->lutW = lw;
    ctx->lutH = lh;
    return VSOM_OK;
}

int launch_online_step(vsom_ctx *ctx, const float *xDev, size_t n, double eta, double sigma, int decay, unsigned *outBmuDev, float *outDistDev,
                       const u64 *lastInDev)
{
    const bool localSearch = !(sigma > 1.0); // SIGMA_SWITCH_TO_LOCAL (include/SOM.hpp:37, src/Som.cpp:889-892)
    if (localSearch && ctx->world > 1)
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "online step: the findLocalBmu regime (sigma <= 1) is not available on node-sharded contexts");
    if (localSearch && !ctx->distBuf)
        VSOM_CUDA(ctx, cudaMalloc(&ctx->distBuf, sizeof(float) * 2 * static_cast<size_t>(ctx->N)));
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
    const int Gp = (G + 15) & ~15;
    VSOM_CUDA(ctx, cudaMemsetAsync(ctx->slots, 0xff, sizeof(u64) * 2 * static_cast<size_t>(G) * Gp, ctx->stream));
    VSOM_CUDA(ctx, cudaMemsetAsync(ctx->errFlag, 0, sizeof(int), ctx->stream));

    StepParams p;
    p.W = ctx->W;
    p.H = ctx->H;
    p.node0 = ctx->node0;
    p.nodeCount = ctx->localN;
    p.world = ctx->world;
    p.rank = ctx->rank;
    p.rankSlots = ctx->rankSlots;
    for (int r = 0; r < 8; ++r)
        p.peerSlots[r] = ctx->peerSlots[r];
    p.stepBase = ctx->stepBase;
    if (ctx->world > 1)
        for (int r = 0; r < ctx->world; ++r)
            if (!ctx->peerSlots[r])
                return set_error(ctx, VSOM_ERR_INVALID, "online step: sharded context without vsom_peer_import for every rank");
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
    p.localSearch = localSearch ? 1 : 0;
    p.distBuf = ctx->distBuf;
    p.lastIn = lastInDev;
    p.prof = ctx->profDev;
    ctx->profSamples = n;
    // the neighbourhood table rides in shared memory behind the resident rows when it fits
    p.lutCount = ctx->lutW * ctx->lutH;
    const size_t lutBytes = sizeof(LutEntry) * static_cast<size_t>(p.lutCount);
    p.lutSmem = (ctx->smemTrain + lutBytes + kStaticSmem <= static_cast<size_t>(ctx->smemOptin)) ? 1 : 0;
    const size_t smemBytes = ctx->smemTrain + (p.lutSmem ? lutBytes : 0);
    p.xVec = (ctx->Din % 4 == 0 && (reinterpret_cast<uintptr_t>(xDev) & 15) == 0) ? 1 : 0;

    p.winTab = nullptr;
    p.pollDelay = 0;
    ctx->lastTrainFast = 0;
    if (ctx->fastTrain && !ctx->fastDisabled)
    {
        // common regime (Standard / Median, reference order, resident planes, sigma > 1, one GPU): K1F
        rc = launch_online_step_fast(ctx, p, sigma);
        if (rc < 0)
            return rc;
        if (rc == 1)
        {
            ctx->lastTrainFast = 1;
            ctx->launches += 1;
            ctx->stepBase += n;
            return VSOM_OK;
        }
    }

    StepKernel k = pick_kernel(ctx->transform, ctx->order, ctx->residentTrain);
    void *args[] = {&p};
    VSOM_CUDA(ctx, cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k), dim3(G), dim3(kThreads), args, smemBytes, ctx->stream));
    ctx->launches += 1;
    ctx->stepBase += n;
    return VSOM_OK;
}

} // namespace vsom

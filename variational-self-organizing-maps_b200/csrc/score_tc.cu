// score_tc.cu — K2: batch BMU search as a dense contraction on the 5th-generation tensor cores, + exact rescore.
//
// Replaces the same row loops as score_exact.cu (Som::evaluate src/Som.cpp:503-520, findBmu :291-309,
// findRestrictedBmu :313-332 over a loaded DataSet) when the batch is large:
//     d(r,p) = |x_r|^2 - 2 x_r . m_p + |m_p|^2            (Standard / Median Comparer, src/Transformation.cpp:7-8)
// The cross term X[R x K] . M^T[K x N] runs as tcgen05.mma (bf16 operands, fp32 accumulators in TMEM), operands
// staged by TMA (128-byte swizzle).  The tensor cores only SELECT candidates; the exact pass re-evaluates them in
// the reference's own f32 arithmetic and order (bit-identical distances) and applies findBmu's lowest-index rule.
//
// Candidate rule (margin list).  With approximate score a_p = c_p - 2 acc_p, E = 2^-7 |x| max_p|m_p| bounds
// |a_p + |x|^2 - d_p| (bf16 rounding of both operands, Cauchy-Schwarz; f32 effects are covered by a relative slack).
// While streaming over the node tiles each row keeps its running best score and appends every node with
// a_p < best + Delta, Delta = 2.2 E + slack, to a small per-row list.  Any node NOT appended had a_p >= best_final +
// Delta, hence d_p >= best_final + |x|^2 + Delta - E > d(best node): it cannot be the BMU, nor tie with it.  So the
// BMU is always in the list — by construction, no statistical recall argument.  At the end of the row tile the list
// is filtered against the final threshold; up to 8 survivors go to the exact rescore, rows with more (or whose list
// overflowed) are re-scored by the exact full scan (score_exact.cu).
//
// Kernel structure (one CTA per SM, persistent over 128-row tiles; 256 threads):
//   warp 0   TMA producer : A tile (128 rows x K, resident for the row tile) + B tiles (256 nodes x 64) through a
//                           3-stage mbarrier ring
//   warp 1   MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16) into one
//                           of two 256-column TMEM accumulators; tcgen05.commit frees the smem stage / publishes
//                           the accumulator
//   warp 2   TMEM allocator (512 columns)
//   warps 4-7 epilogue    : thread = one row (TMEM lane); tcgen05.ld 32 columns at a time (next chunk in flight),
//                           per column: FFMA (score), compare, predicated append; running best via FMNMX
// Limits of this version: Standard / Median transformation, Dm <= 256 (A tile resident); other shapes use K3.
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace vsom
{

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 3, TC_TOPK = 16, TC_THREADS = 384;
constexpr int TC_LIST = 24;    // per-thread candidate list (shared memory); a full list sends the row to the exact scan
constexpr int TC_LIST_HI = 12; // lists are compacted against the current threshold when one passes this length
constexpr unsigned TC_OVERFLOW = 255;
constexpr int TC_MAXK = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB per k-block
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;  // 32 KB per stage

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// bounded wait: a pipeline bug must end the kernel (error flag), not hang the GPU
__device__ __forceinline__ bool mbar_wait(void *bar, unsigned parity, int *err)
{
    for (long long spin = 0; spin < (1ll << 26); ++spin)
        if (mbar_try_wait(bar, parity))
            return true;
    *err = 2;
    return false;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, void *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(unsigned tmemC, u64 descA, u64 descB, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemC), "l"(descA),
                 "l"(descB), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned addr, unsigned (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void sts64(unsigned addr, unsigned lo, unsigned hi)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO (ignored for swizzled K-major, set 1) in [16,30), SBO = 1024 B (8 rows x 128 B)
// in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ u64 umma_desc_sw128(unsigned smemAddr)
{
    return static_cast<u64>((smemAddr & 0x3ffff) >> 4) | (1ull << 16) | (static_cast<u64>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (1<<4), a=b=bf16 (1<<7, 1<<10), both K-major, N>>3 at 17, M>>4 at 24
constexpr unsigned kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<unsigned>(TC_BN >> 3) << 17) | (static_cast<unsigned>(TC_BM >> 4) << 24);

struct TcShared
{
    // offsets inside the dynamic shared buffer (1024-byte aligned base)
    static constexpr int A_OFF = 0;                                   // up to 4 k-blocks x 16 KB
    static constexpr int B_OFF = A_OFF + (TC_MAXK / TC_BK) * TC_A_BYTES; // 4 stages x 32 KB
    static constexpr int CN_OFF = B_OFF + TC_STAGES * TC_B_BYTES;      // 2 x 256 floats
    static constexpr int BAR_OFF = CN_OFF + 2 * TC_BN * 4;             // mbarriers
    static constexpr int LIST_OFF = BAR_OFF + 256;                     // TC_LIST x 256 x {score, node}: per-thread candidate lists
    static constexpr int MERGE_OFF = LIST_OFF + TC_LIST * 256 * 8;     // 256 floats + 256 ints: merge of the two half rows
    static constexpr int TOTAL = MERGE_OFF + 256 * 8;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapM, const float *__restrict__ cnorm, int rowsTotal,
                int numRowTiles, int numNodeTiles, int kBlocks, int stagger, const float *__restrict__ xnorm2, const float *__restrict__ maxNorm2, unsigned *__restrict__ candOut,
                unsigned *__restrict__ countOut, int *err)
{
    // 128-byte-swizzled TMA / UMMA tiles need a 1024-byte aligned base; declaring the alignment (instead of rounding
    // a pointer up by hand) also lets the compiler keep every derived pointer in the shared address space (LDS/STS)
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem + TcShared::A_OFF;
    unsigned char *sB = smem + TcShared::B_OFF;
    float *sCn = reinterpret_cast<float *>(smem + TcShared::CN_OFF);
    u64 *bars = reinterpret_cast<u64 *>(smem + TcShared::BAR_OFF);
    u64 *bFull = bars, *bEmpty = bars + TC_STAGES, *aFull = bars + 2 * TC_STAGES, *aEmpty = aFull + 1, *tFull = aEmpty + 1, *tEmpty = tFull + 2;
    unsigned *tmemBaseSlot = reinterpret_cast<unsigned *>(tEmpty + 2);
    uint2 *lists = reinterpret_cast<uint2 *>(smem + TcShared::LIST_OFF);
    float *sBest = reinterpret_cast<float *>(smem + TcShared::MERGE_OFF);
    int *sCnt = reinterpret_cast<int *>(sBest + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTAs walk the node tiles from different starting points so that at any moment they pull DIFFERENT B tiles out
    // of L2 (all starting at tile 0 makes 148 SMs ask the same lines at once)
    const int ntStart = (stagger & 1) ? static_cast<int>((static_cast<unsigned>(blockIdx.x) * 7u) % static_cast<unsigned>(numNodeTiles)) : 0;

    if (warp == 1 && lane == 0)
    {
        for (int s = 0; s < TC_STAGES; ++s)
        {
            mbar_init(&bFull[s], 1);
            mbar_init(&bEmpty[s], 1);
        }
        mbar_init(aFull, 1);
        mbar_init(aEmpty, 1);
        for (int a = 0; a < 2; ++a)
        {
            mbar_init(&tFull[a], 1);
            mbar_init(&tEmpty[a], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmemBaseSlot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmemBase = *tmemBaseSlot;

    if (warp == 0)
    {
        // ===================================================== TMA producer
        if (lane == 0)
        {
            unsigned stage = 0, phase = 0, aPhase = 0;
            bool ok = true;
            for (int rt = blockIdx.x; rt < numRowTiles && ok; rt += gridDim.x)
            {
                ok = mbar_wait(aEmpty, aPhase ^ 1, err); // previous row tile's MMAs are done with A
                aPhase ^= 1;
                mbar_expect_tx(aFull, static_cast<unsigned>(kBlocks) * TC_A_BYTES);
                for (int kb = 0; kb < kBlocks; ++kb)
                    tma_load_2d(sA + kb * TC_A_BYTES, &mapX, aFull, kb * TC_BK, rt * TC_BM);
                for (int i = 0; i < numNodeTiles && ok; ++i)
                    for (int kb = 0; kb < kBlocks && ok; ++kb)
                    {
                        const int nt = (i + ntStart) % numNodeTiles;
                        ok = mbar_wait(&bEmpty[stage], phase ^ 1, err);
                        mbar_expect_tx(&bFull[stage], TC_B_BYTES);
                        tma_load_2d(sB + stage * TC_B_BYTES, &mapM, &bFull[stage], kb * TC_BK, nt * TC_BN);
                        if (++stage == TC_STAGES)
                        {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
            }
        }
    }
    else if (warp == 1)
    {
        // ===================================================== MMA issuer (one lane)
        if (lane == 0)
        {
            unsigned stage = 0, phase = 0, aPhase = 0, acc = 0, accPhase = 0;
            bool ok = true;
            for (int rt = blockIdx.x; rt < numRowTiles && ok; rt += gridDim.x)
            {
                ok = mbar_wait(aFull, aPhase, err);
                aPhase ^= 1;
                for (int nt = 0; nt < numNodeTiles && ok; ++nt)
                {
                    ok = mbar_wait(&tEmpty[acc], accPhase ^ 1, err); // epilogue drained this accumulator
                    tc_fence_after();
                    const unsigned tmemC = tmemBase + acc * TC_BN;
                    for (int kb = 0; kb < kBlocks && ok; ++kb)
                    {
                        ok = mbar_wait(&bFull[stage], phase, err);
                        tc_fence_after();
                        const unsigned aAddr = smem_u32(sA + kb * TC_A_BYTES), bAddr = smem_u32(sB + stage * TC_B_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                        {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row
                            const u64 da = umma_desc_sw128(aAddr + k * 32), db = umma_desc_sw128(bAddr + k * 32);
                            umma_f16(tmemC, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&bEmpty[stage]); // frees the B stage when the MMAs above have read it
                        if (++stage == TC_STAGES)
                        {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(&tFull[acc]); // accumulator complete
                    if (++acc == 2)
                    {
                        acc = 0;
                        accPhase ^= 1;
                    }
                }
                umma_commit(aEmpty); // all MMAs of this row tile are done reading A
            }
        }
    }
    else if (warp >= 4)
    {
        // ===================================================== epilogue: 8 warps, two threads per row of the tile
        // Warps 4-7 score columns 0..127 of every accumulator, warps 8-11 columns 128..255 (warp w may only touch the
        // TMEM lane quarter w % 4, so both halves cover all 128 rows).  Two resident warps per scheduler hide each
        // other's LDS / ALU latencies; each thread keeps its own running best and its own short list, merged per row
        // at the end of the row tile.
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;  // 0: columns 0..127, 1: columns 128..255
        const int rowInTile = q * 32 + lane;
        const int et = threadIdx.x - 128;  // 0..255 ; partner thread of the same row: et ^ 128
        uint2 *mine = lists + et;          // entry e of this thread at mine[e * 256]
        const unsigned mineAddr = smem_u32(mine);
        const unsigned mineEnd = mineAddr + TC_LIST * 2048u;
        unsigned acc = 0, accPhase = 0;
        bool ok = true;
        const float mx = maxNorm2[0];
        const float inf = __int_as_float(0x7f800000);

        // drop list entries that are no longer below the (tightened) threshold; warp-uniform control flow
        auto compact = [&](int &cnt, float thr) {
            const int maxc = __reduce_max_sync(0xffffffffu, cnt);
            int k = 0;
#pragma unroll 1
            for (int e = 0; e < maxc; ++e)
                if (e < cnt)
                {
                    const uint2 v = mine[e * 256];
                    if (__uint_as_float(v.x) < thr)
                    {
                        mine[k * 256] = v;
                        ++k;
                    }
                }
            cnt = k;
        };

        for (int rt = blockIdx.x; rt < numRowTiles && ok; rt += gridDim.x)
        {
            const long long row = static_cast<long long>(rt) * TC_BM + rowInTile;
            const float xn = row < rowsTotal ? xnorm2[row] : 0.0f;
            const float E = 0.0079f * sqrtf(xn * mx) + 1e-5f * (xn + mx);
            const float delta = 2.2f * E + 2e-4f * (xn + mx);
            float best = inf, thr = inf;
            int cnt = 0;
            bool ovf = false;
            unsigned wp = mineAddr; // shared address of the next free entry of this thread's list (2048 bytes apart)
            for (int i = 0; i < numNodeTiles && ok; ++i)
            {
                const int nt = (i + ntStart) % numNodeTiles;
                // per-node constants of this tile
                float *cn = sCn + acc * TC_BN;
                cn[et] = cnorm[nt * TC_BN + et];
                asm volatile("bar.sync 1, 256;" ::: "memory");
                ok = mbar_wait(&tFull[acc], accPhase, err);
                tc_fence_after();
                const unsigned taddr = tmemBase + (static_cast<unsigned>(q * 32) << 16) + acc * TC_BN + half * (TC_BN / 2);
                const float *cnh = cn + half * (TC_BN / 2);
                unsigned v[2][32];
                if ((stagger & 2) == 0) // bit 1 of the debug word: skip the column work (pipeline-only timing)
                {
                    tmem_ld32(taddr, v[0]);
#pragma unroll
                    for (int c = 0; c < TC_BN / 64; ++c)
                    {
                        tmem_wait_ld();
                        if (c + 1 < TC_BN / 64)
                            tmem_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                        cnt = static_cast<int>((wp - mineAddr) >> 11);
                        if (__any_sync(0xffffffffu, cnt > TC_LIST_HI))
                        {
                            compact(cnt, thr);
                            wp = mineAddr + (static_cast<unsigned>(cnt) << 11);
                        }
                        unsigned(&w)[32] = v[c & 1];
                        const unsigned nodeBase = static_cast<unsigned>(nt * TC_BN + half * (TC_BN / 2) + c * 32);
                        // All 32 scores of the chunk and the per-group minima first (independent instructions, no
                        // branch).  gm = groups of four columns in which this row may have something to append.
                        unsigned gm = 0;
                        float cm = inf;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                        {
                            const float4 k4 = *reinterpret_cast<const float4 *>(cnh + c * 32 + j4 * 4);
                            const float s0 = fmaf(-2.0f, __uint_as_float(w[j4 * 4 + 0]), k4.x);
                            const float s1 = fmaf(-2.0f, __uint_as_float(w[j4 * 4 + 1]), k4.y);
                            const float s2 = fmaf(-2.0f, __uint_as_float(w[j4 * 4 + 2]), k4.z);
                            const float s3 = fmaf(-2.0f, __uint_as_float(w[j4 * 4 + 3]), k4.w);
                            w[j4 * 4 + 0] = __float_as_uint(s0);
                            w[j4 * 4 + 1] = __float_as_uint(s1);
                            w[j4 * 4 + 2] = __float_as_uint(s2);
                            w[j4 * 4 + 3] = __float_as_uint(s3);
                            const float m4 = fminf(fminf(s0, s1), fminf(s2, s3));
                            gm |= (m4 < thr) ? (1u << j4) : 0u; // against the (valid, possibly stale) old threshold
                            cm = fminf(cm, m4);
                        }
                        // the chunk's own minimum tightens the threshold BEFORE anything is appended
                        best = fminf(best, cm);
                        thr = best + delta;
                        // one REDUX tells the whole warp which groups need the (predicated) appends; the branches
                        // below are on a warp-uniform value
                        const unsigned any = __reduce_or_sync(0xffffffffu, gm);
                        if (any)
                        {
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4)
                                if (any & (1u << j4))
                                {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (__uint_as_float(w[j4 * 4 + k]) < thr)
                                        {
                                            if (wp < mineEnd)
                                            {
                                                sts64(wp, w[j4 * 4 + k], nodeBase + j4 * 4 + k);
                                                wp += 2048;
                                            }
                                            else
                                                ovf = true; // list full: this row takes the exact scan
                                        }
                                }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tEmpty[acc]);
                if (++acc == 2)
                {
                    acc = 0;
                    accPhase ^= 1;
                }
            }
            // merge the two halves of the row: common best -> final threshold -> filter both lists -> candidates
            cnt = static_cast<int>((wp - mineAddr) >> 11);
            sBest[et] = best;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float thrF = fminf(best, sBest[et ^ 128]) + delta;
            compact(cnt, thrF);
            sCnt[et] = ovf ? 1000 : cnt;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int other = sCnt[et ^ 128], total = sCnt[et] + other;
            if (row < rowsTotal)
            {
                const bool bad = total > TC_TOPK || total == 0;
                if (!bad)
                {
                    const int off = half ? other : 0;
                    for (int e = 0; e < cnt; ++e)
                        candOut[row * TC_TOPK + off + e] = mine[e * 256].y;
                }
                if (half == 0)
                    countOut[row] = bad ? TC_OVERFLOW : static_cast<unsigned>(total);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory"); // sBest / sCnt are reused by the next row tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ operand preparation

// rows of f32 -> bf16 (zero padded to Kpad), optional |row|^2 (f32, any order: only used by the guard)
__global__ void to_bf16_rows_kernel(const float *__restrict__ src, long long rows, int D, int srcStride, __nv_bfloat16 *__restrict__ dst, int Kpad,
                                    float *__restrict__ norm2)
{
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows)
        return;
    float s = 0.0f;
    for (int k = lane; k < Kpad; k += 32)
    {
        const float v = k < D ? src[row * srcStride + k] : 0.0f;
        dst[row * Kpad + k] = __float2bfloat16_rn(v);
        s += v * v;
    }
    for (int o = 16; o; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && norm2)
        norm2[row] = s;
}

// per-node constant of the score: |m_p|^2 for nodes that may win, +inf for padding and for nodes below min_hits
// (node 0 always competes: it seeds findRestrictedBmu regardless of its hit count, src/Som.cpp:316-322)
__global__ void node_const_kernel(const float *__restrict__ norm2, const u64 *__restrict__ hits, u64 minHits, int N, int Npad, float *__restrict__ cnorm,
                                  float *__restrict__ maxNorm2)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Npad)
        return;
    float c = __int_as_float(0x7f800000);
    if (p < N && (p == 0 || minHits == 0 || hits[p] >= minHits))
        c = norm2[p];
    cnorm[p] = c;
    if (p < N)
        atomicMax(reinterpret_cast<int *>(maxNorm2), __float_as_int(norm2[p])); // non-negative floats order like ints
}

// ------------------------------------------------------------------------------------------------ exact rescore + guard

// Re-evaluation of the candidates with the reference's sequential f32 chain.  Rows keep 1-3 candidates on average, so a
// thread per (row, slot) would leave nine lanes in ten idle: a CTA takes 128 rows, compacts their (row, candidate) items into
// a list (block scan of the counts) and its 256 threads walk the list, all lanes busy; the per-row winner is an atomic min of
// the (distance, node) key in shared memory — lowest index on ties, like the reference.  Rows whose list overflowed, or whose
// winner is NaN, go to the fallback list (exact full scan).
constexpr int RS_ROWS = 128;
constexpr int RS_THREADS = 256;
__global__ void __launch_bounds__(RS_THREADS) rescore_kernel(const float *__restrict__ x, long long rows, int D, const float *__restrict__ mean, int rowStride,
                                                             const unsigned *__restrict__ cand, const unsigned *__restrict__ count, unsigned *__restrict__ outBmu,
                                                             float *__restrict__ outDist, unsigned *__restrict__ fallbackRows, unsigned *__restrict__ fallbackCount)
{
    __shared__ unsigned sOff[RS_ROWS + 1];
    __shared__ unsigned sWarpTot[RS_ROWS / 32];
    __shared__ u64 sKey[RS_ROWS];
    __shared__ unsigned short sItem[RS_ROWS * TC_TOPK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = static_cast<long long>(blockIdx.x) * RS_ROWS;
    unsigned cnt = 0, c = 0;
    if (tid < RS_ROWS)
    {
        const long long row = row0 + tid;
        cnt = row < rows ? count[row] : 0u;
        c = cnt == TC_OVERFLOW ? 0u : cnt;
        sKey[tid] = ~0ull;
        unsigned inc = c; // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += v;
        }
        if (lane == 31)
            sWarpTot[warp] = inc;
        sOff[tid] = inc - c; // exclusive, warp-local for now
    }
    __syncthreads();
    if (tid < RS_ROWS)
    {
        unsigned base = 0;
        for (int w = 0; w < warp; ++w)
            base += sWarpTot[w];
        const unsigned off = sOff[tid] + base;
        for (unsigned j = 0; j < c; ++j)
            sItem[off + j] = static_cast<unsigned short>((tid << 4) | j);
        if (tid == RS_ROWS - 1)
            sOff[RS_ROWS] = off + c;
    }
    __syncthreads();
    const unsigned total = sOff[RS_ROWS];
    for (unsigned i = tid; i < total; i += RS_THREADS)
    {
        const unsigned it = sItem[i];
        const int r = static_cast<int>(it >> 4), j = static_cast<int>(it & 15u);
        const long long row = row0 + r;
        const unsigned node = cand[row * TC_TOPK + j];
        const float *m = mean + static_cast<size_t>(node) * rowStride, *xr = x + row * D;
        float s = 0.0f;
        int k = 0;
        if ((D & 3) == 0) // rows of x are 16-byte aligned: 128-bit loads, same sequential order of the adds
        {
            const float4 *m4 = reinterpret_cast<const float4 *>(m), *x4 = reinterpret_cast<const float4 *>(xr);
#pragma unroll 4
            for (int q = 0; q < (D >> 2); ++q)
            {
                const float4 a = m4[q], b = x4[q];
                float d = __fsub_rn(a.x, b.x);
                s = __fadd_rn(s, __fmul_rn(d, d));
                d = __fsub_rn(a.y, b.y);
                s = __fadd_rn(s, __fmul_rn(d, d));
                d = __fsub_rn(a.z, b.z);
                s = __fadd_rn(s, __fmul_rn(d, d));
                d = __fsub_rn(a.w, b.w);
                s = __fadd_rn(s, __fmul_rn(d, d));
            }
            k = D;
        }
        for (; k < D; ++k)
        {
            const float d = __fsub_rn(m[k], xr[k]);
            s = __fadd_rn(s, __fmul_rn(d, d));
        }
        atomicMin(&sKey[r], make_key(s, node, (s != s) ? 1u : 0u));
    }
    __syncthreads();
    if (tid < RS_ROWS && row0 + tid < rows)
    {
        const long long row = row0 + tid;
        const u64 key = sKey[tid];
        const bool nan = (key & 1ull) != 0;
        if (cnt == TC_OVERFLOW || nan || key == ~0ull)
            fallbackRows[atomicAdd(fallbackCount, 1u)] = static_cast<unsigned>(row);
        else
        {
            if (outBmu)
                outBmu[row] = key_node(key);
            if (outDist)
                outDist[row] = __uint_as_float(static_cast<unsigned>(key >> 32));
        }
    }
}

// Exact full scan for a FEW rows (the rows the candidate lists could not certify): one CTA per row, threads over
// nodes, every (row, node) distance summed sequentially in f32 like the reference.  Same key rule as K3.
// Exact full scan of the rows a slab's guard could not certify.  The row count lives in device memory (no host round trip
// between the slabs of a batch): a fixed grid walks the list, and CTA 0 adds the count to the batch total.
__global__ void __launch_bounds__(256) find_bmu_rowwise_kernel(const float *__restrict__ x, int D, const unsigned *__restrict__ rows,
                                                               const unsigned *__restrict__ countPtr, const float *__restrict__ mean, int rowStride, int N,
                                                               const u64 *__restrict__ hits, u64 minHits, unsigned *__restrict__ outBmu, float *__restrict__ outDist,
                                                               unsigned long long *__restrict__ total)
{
    extern __shared__ float xrow[];
    __shared__ u64 wkey[8];
    const unsigned count = *countPtr;
    if (blockIdx.x == 0 && threadIdx.x == 0 && count)
        atomicAdd(total, static_cast<unsigned long long>(count));
    for (unsigned i = blockIdx.x; i < count; i += gridDim.x)
    {
        const size_t r = rows[i];
        __syncthreads(); // xrow / wkey of the previous row are no longer read
        for (int k = threadIdx.x; k < D; k += blockDim.x)
            xrow[k] = x[r * D + k];
        __syncthreads();
        u64 best = ~0ull;
        for (int node = threadIdx.x; node < N; node += blockDim.x)
        {
            if (!(node == 0 || minHits == 0 || hits[node] >= minHits))
                continue;
            const float *m = mean + static_cast<size_t>(node) * rowStride;
            float s = 0.0f;
            int k = 0;
            for (; k + 4 <= D; k += 4)
            {
                const float4 a = *reinterpret_cast<const float4 *>(m + k);
                float q = __fsub_rn(a.x, xrow[k]);
                s = __fadd_rn(s, __fmul_rn(q, q));
                q = __fsub_rn(a.y, xrow[k + 1]);
                s = __fadd_rn(s, __fmul_rn(q, q));
                q = __fsub_rn(a.z, xrow[k + 2]);
                s = __fadd_rn(s, __fmul_rn(q, q));
                q = __fsub_rn(a.w, xrow[k + 3]);
                s = __fadd_rn(s, __fmul_rn(q, q));
            }
            for (; k < D; ++k)
            {
                const float q = __fsub_rn(m[k], xrow[k]);
                s = __fadd_rn(s, __fmul_rn(q, q));
            }
            best = u64_min(best, make_key(s, static_cast<unsigned>(node), (s != s) ? 1u : 0u));
        }
        best = warp_min_u64(best);
        if ((threadIdx.x & 31) == 0)
            wkey[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            for (int w = 1; w < 8; ++w)
                best = u64_min(best, wkey[w]);
            if (outBmu)
                outBmu[r] = key_node(best);
            if (outDist)
                outDist[r] = (best & 1ull) ? __uint_as_float(0x7fc00000u) : __uint_as_float(static_cast<unsigned>(best >> 32));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map(vsom_ctx *ctx, CUtensorMap *map, void *base, unsigned long long rows, int Kpad, int boxRows)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc)
        return set_error(ctx, VSOM_ERR_CUDA, "score_tc: cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kpad), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kpad) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(TC_BK), static_cast<cuuint32_t>(boxRows)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(ctx, VSOM_ERR_CUDA, "score_tc: cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")");
    return VSOM_OK;
}

bool score_tc_supported(const vsom_ctx *ctx) { return ctx->transform != VSOM_CLR && ctx->Dm <= TC_MAXK; }

// stage slots used here: 6 = bf16 map + node constants, 7 = bf16 rows of the current slab, 8 = per-row scratch
int launch_find_bmu_tc(vsom_ctx *ctx, const float *xDev, size_t n, uint64_t minHits, unsigned *outBmuDev, float *outDistDev, unsigned long long *fallbackRowsOut)
{
    if (!score_tc_supported(ctx))
        return set_error(ctx, VSOM_ERR_UNSUPPORTED, "score_tc: needs a Standard/Median transformation and Dm <= 256");
    if (fallbackRowsOut)
        *fallbackRowsOut = 0;
    if (n == 0)
        return VSOM_OK;
    const int D = ctx->Dm, N = ctx->N;
    const int Kpad = (D + TC_BK - 1) / TC_BK * TC_BK, kBlocks = Kpad / TC_BK;
    const int Npad = (N + TC_BN - 1) / TC_BN * TC_BN, nodeTiles = Npad / TC_BN;

    // ---- map side: bf16 copy, |m|^2, node constants, max |m|^2
    const size_t mbBytes = sizeof(__nv_bfloat16) * static_cast<size_t>(Npad) * Kpad;
    int rc = stage_reserve(ctx, 6, mbBytes + sizeof(float) * (2 * static_cast<size_t>(Npad) + 64));
    if (rc)
        return rc;
    __nv_bfloat16 *Mb = static_cast<__nv_bfloat16 *>(ctx->stage[6]);
    float *mnorm = reinterpret_cast<float *>(static_cast<unsigned char *>(ctx->stage[6]) + mbBytes);
    float *cnorm = mnorm + Npad;
    float *maxNorm2 = cnorm + Npad;
    VSOM_CUDA(ctx, cudaMemsetAsync(Mb, 0, mbBytes, ctx->stream));
    VSOM_CUDA(ctx, cudaMemsetAsync(maxNorm2, 0, sizeof(float), ctx->stream));
    to_bf16_rows_kernel<<<(N + 7) / 8, 256, 0, ctx->stream>>>(ctx->mean, N, D, ctx->rowStride, Mb, Kpad, mnorm);
    node_const_kernel<<<(Npad + 255) / 256, 256, 0, ctx->stream>>>(mnorm, ctx->hits, minHits, N, Npad, cnorm, maxNorm2);
    ctx->launches += 2;
    CUtensorMap mapM;
    rc = make_map(ctx, &mapM, Mb, static_cast<unsigned long long>(Npad), Kpad, TC_BN);
    if (rc)
        return rc;

    static bool attrSet = false;
    const int smemBytes = TcShared::TOTAL;
    if (!attrSet)
    {
        VSOM_CUDA(ctx, cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
        attrSet = true;
    }

    // ---- rows, in slabs of 4M (the bf16 staging buffer: 2 GB at K = 256).  Two streams: the context's stream converts a slab and runs the tensor-core search; an auxiliary
    // stream re-scores the slab's candidates (and scans the few rows its guard could not certify) while the next slab is
    // already being searched.  Candidate scratch is double-buffered by slab parity; nothing returns to the host in between.
    const char *slabEnv = getenv("VSOM_TC_SLAB_LOG2"), *ovlEnv = getenv("VSOM_TC_OVERLAP"); // experiment knobs
    const int slabLog2 = slabEnv ? atoi(slabEnv) : 22; // smaller slabs measured slower (per-slab ramp and tail of the persistent kernel)
    const bool overlap = ovlEnv ? atoi(ovlEnv) != 0 : true;
    const size_t slabRows = std::min<size_t>(n, static_cast<size_t>(1) << slabLog2);
    rc = stage_reserve(ctx, 7, sizeof(__nv_bfloat16) * slabRows * Kpad);
    if (rc)
        return rc;
    // per-row scratch: candidates, their count, |x|^2, fallback list; per set: fallback count
    const size_t perRow = sizeof(unsigned) * TC_TOPK + sizeof(unsigned) + sizeof(float) + sizeof(unsigned);
    const size_t setBytes = (perRow * slabRows + 256 + 255) & ~static_cast<size_t>(255);
    rc = stage_reserve(ctx, 8, 2 * setBytes + 256);
    if (rc)
        return rc;
    if (!ctx->auxStream)
    {
        VSOM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->auxStream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i)
        {
            VSOM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evScore[i], cudaEventDisableTiming));
            VSOM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evDone[i], cudaEventDisableTiming));
        }
    }
    __nv_bfloat16 *Xb = static_cast<__nv_bfloat16 *>(ctx->stage[7]);
    unsigned long long *totalDev = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(ctx->stage[8]) + 2 * setBytes);
    VSOM_CUDA(ctx, cudaMemsetAsync(totalDev, 0, sizeof(unsigned long long), ctx->stream));
    VSOM_CUDA(ctx, cudaMemsetAsync(ctx->errFlag, 0, sizeof(int), ctx->stream));

    const char *stg = getenv("VSOM_TC_STAGGER");
    const int stagger = stg ? atoi(stg) : 1;
    size_t slab = 0;
    for (size_t r0 = 0; r0 < n; r0 += slabRows, ++slab)
    {
        const int par = static_cast<int>(slab & 1);
        unsigned char *set = static_cast<unsigned char *>(ctx->stage[8]) + par * setBytes;
        unsigned *cand = reinterpret_cast<unsigned *>(set);
        unsigned *candCount = cand + slabRows * TC_TOPK;
        float *xnorm = reinterpret_cast<float *>(candCount + slabRows);
        unsigned *fbRows = reinterpret_cast<unsigned *>(xnorm + slabRows);
        unsigned *fbCount = fbRows + slabRows;

        const size_t rows = std::min(slabRows, n - r0);
        const float *xs = xDev + r0 * D;
        if (slab >= 2)
            VSOM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evDone[par], 0)); // slab - 2 is done with this scratch set
        to_bf16_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, ctx->stream>>>(xs, static_cast<long long>(rows), D, D, Xb, Kpad, xnorm);
        CUtensorMap mapX;
        rc = make_map(ctx, &mapX, Xb, rows, Kpad, TC_BM);
        if (rc)
            return rc;
        const int rowTiles = static_cast<int>((rows + TC_BM - 1) / TC_BM);
        const int grid = std::min(rowTiles, ctx->numSMs);
        VSOM_CUDA(ctx, cudaMemsetAsync(fbCount, 0, sizeof(unsigned), ctx->stream));
        score_tc_kernel<<<grid, TC_THREADS, smemBytes, ctx->stream>>>(mapX, mapM, cnorm, static_cast<int>(rows), rowTiles, nodeTiles, kBlocks, stagger, xnorm, maxNorm2, cand, candCount,
                                                                      ctx->errFlag);
        cudaStream_t rs = overlap ? ctx->auxStream : ctx->stream;
        VSOM_CUDA(ctx, cudaEventRecord(ctx->evScore[par], ctx->stream));
        VSOM_CUDA(ctx, cudaStreamWaitEvent(rs, ctx->evScore[par], 0));
        rescore_kernel<<<static_cast<unsigned>((rows + RS_ROWS - 1) / RS_ROWS), RS_THREADS, 0, rs>>>(
            xs, static_cast<long long>(rows), D, ctx->mean, ctx->rowStride, cand, candCount, outBmuDev ? outBmuDev + r0 : nullptr,
            outDistDev ? outDistDev + r0 : nullptr, fbRows, fbCount);
        // rows the guard could not certify: exact full scan, count read on the device
        find_bmu_rowwise_kernel<<<2 * ctx->numSMs, 256, sizeof(float) * D, rs>>>(xs, D, fbRows, fbCount, ctx->mean, ctx->rowStride, N, ctx->hits, minHits,
                                                                                             outBmuDev ? outBmuDev + r0 : nullptr,
                                                                                             outDistDev ? outDistDev + r0 : nullptr, totalDev);
        VSOM_CUDA(ctx, cudaEventRecord(ctx->evDone[par], rs));
        ctx->launches += 4;
        VSOM_CUDA(ctx, cudaGetLastError());
    }
    // the context's stream owns the results again
    for (int i = 0; i < 2 && static_cast<size_t>(i) < slab; ++i)
        VSOM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evDone[i], 0));
    unsigned long long totalFallback = 0;
    int flag = 0;
    VSOM_CUDA(ctx, cudaMemcpyAsync(&totalFallback, totalDev, sizeof(totalFallback), cudaMemcpyDeviceToHost, ctx->stream));
    VSOM_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->errFlag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VSOM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag)
        return set_error(ctx, VSOM_ERR_TIMEOUT, "score_tc: pipeline barrier timed out (kernel bug)");
    VSOM_CUDA(ctx, cudaGetLastError());
    if (fallbackRowsOut)
        *fallbackRowsOut = totalFallback;
    ctx->lastFallbackRows = totalFallback;
    return VSOM_OK;
}

} // namespace vsom

// umatrix.cu — K4: the U-matrix stencil.
//
// Replaces Som::updateUMatrix (src/Som.cpp:999-1111): for every node the sigma-weighted distance
// Som::euclidianWeightedDistRaw (src/Som.cpp:143-157) to each existing neighbour of its 8-neighbourhood,
// straight neighbours weighted 1, diagonal ones 0.3, summed left to right in double in the textual order of
// the source and divided by 8 / 5 / 3 (interior / edge / corner).  Raw(p, u) always divides by p's OWN
// clamped sigma, so the stencil is not symmetric.
//
// One thread owns one (node, neighbour) pair and sums its Dm terms in f32 in the context's summation order (sequential
// like the stand-in dot product, or Eigen's SSE2 packet order: eight interleaved chains, common.cuh) — every Raw value
// and every U value is bit-identical to the reference.  A CTA covers 16 consecutive nodes of one grid row; the eight
// threads of a node are eight consecutive lanes.  Rows are read straight from global memory with 128-bit read-only
// loads (a node's eight threads ask for the same words of its own rows: one transaction; neighbouring nodes share
// neighbour rows, L1 serves the reuse).
//
// The arithmetic is what bounds this kernel: the reference formula has one IEEE division per element AND neighbour
// (div.rn.f32 is ~9 issue slots).  Here the eight neighbour threads of a node share ONE correctly rounded reciprocal per
// element (each lane computes y = RN(1 / sM) for the elements k = lane mod 8 and hands it round by shuffle) and every
// division becomes Markstein's correction step:  q = RN(d y),  r = d - sM q (exact, one FMA),  d / sM = RN(q + r y),
// which IS the correctly rounded quotient when y is the correctly rounded reciprocal and nothing overflows or underflows
// on the way (Markstein 1990; Cornea, Harrison & Tang, "Scientific computing on Itanium-based systems").  Range argument,
// with sM in [1e-5, 2^100] (checked once per block of eight; the clamp gives the lower end):
//   * |d| >= 2^100, infinities and NaN take div.rn itself (one predicate per element);
//   * the residual r is a multiple of 2^(e_d - 46); whenever the TERM a^2 can be non-zero at all (|a| >= 2^-75, i.e.
//     |d| >= 2^-75 sM >= 2^-92) that is >= 2^-138, representable, so r is exact and a is correctly rounded;
//   * for smaller |d| both the fast and the exact quotient are below 2^-83 in magnitude and their squares round to +0:
//     the term is +0 either way (and d == 0 gives a == 0 on both paths).
// So every TERM is bit-identical to the reference's although the division is 3 issue slots instead of ~9.  The U-matrix
// parity tests (tests/test_gpu_parity.py) cover maps with zero / tiny / huge sigma and differences.
//
// Rows are addressed through per-grid-row pointer tables, so the same kernel serves node-sharded contexts: the rows of
// other ranks that border this rank's row blocks are fetched into a halo buffer over NVLink first (capi.cu).
#include "common.cuh"

namespace vsom
{

constexpr int UT = 16; // nodes per CTA

// neighbour order of the reference's sums: W, E, S, N, then the diagonals NW, SW, NE, SE
// (src/Som.cpp:1017-1025; the edge and corner cases drop the missing terms and keep this order).
__constant__ int kDi[8] = {0, 0, +1, -1, -1, +1, -1, +1};
__constant__ int kDj[8] = {-1, +1, 0, 0, -1, -1, +1, +1};

// sM = max(sigma, 1e-5) as the reference's select (:150; a NaN sigma stays NaN)
__device__ __forceinline__ float clamp_sigma(float sg) { return sg < 0.00001f ? 0.00001f : sg; }

// sM inside the range the correction step is proven for (a NaN sigma fails)
__device__ __forceinline__ bool rcp_ok(float sM) { return sM <= 1.2676506e30f; }

// one term of Raw(p, u): a = (m_p - m_u) / sM (:156), returned as a * a.  (m - v) * (valid * weights) / sM equals a bit for
// bit because valid * weights == 1.0f exactly (:1002-1003), so one division serves both factors of the dot product.
// `bad` collects the elements whose fast quotient is not proven (see the head of this file).
__device__ __forceinline__ float raw_term_fast(float mc, float mu, float sM, float y, bool &bad)
{
    const float d = __fsub_rn(mc, mu);
    const float q = __fmul_rn(d, y);
    const float r = __fmaf_rn(-sM, q, d);
    const float a = __fmaf_rn(r, y, q);
    bad = bad || !(fabsf(d) < 1.2676506e30f); // 2^100; NaN fails the comparison
    return __fmul_rn(a, a);
}
// the reference's own arithmetic, for the blocks the fast path cannot prove (rare; one branch per block of eight)
__device__ __forceinline__ void raw_terms_exact(const float (&mcv)[8], const float (&muv)[8], const float (&sMv)[8], float (&tv)[8])
{
#pragma unroll
    for (int e = 0; e < 8; ++e)
    {
        const float a = __fdiv_rn(__fsub_rn(mcv[e], muv[e]), sMv[e]);
        tv[e] = __fmul_rn(a, a);
    }
}

template <int ORDER>
__global__ void __launch_bounds__(UT * 8) umatrix_kernel(const float *const *__restrict__ meanRow, const float *const *__restrict__ sigmaRow,
                                                         const int *__restrict__ rowsY, int W, int H, int Dm, int rowStride, double *__restrict__ out)
{
    __shared__ float res[UT][8];

    const int tid = threadIdx.x, lane = tid & 31;
    const int i = rowsY[blockIdx.y]; // grid row
    const int j0 = blockIdx.x * UT;  // first grid column of the tile
    const int t = tid >> 3, nb = tid & 7;
    const int j = j0 + t;
    const int ni = i + kDi[nb], nj = j + kDj[nb];
    const bool node = j < W;
    const bool active = node && ni >= 0 && ni < H && nj >= 0 && nj < W;
    const int grp = lane & ~7; // first lane of this node's eight threads

    // every lane takes part in the shuffles; lanes without a neighbour read their own row as the neighbour (distance 0)
    const int jc = node ? j : 0;
    const float *mc = meanRow[i] + static_cast<size_t>(jc) * rowStride;
    const float *sg = sigmaRow[i] + static_cast<size_t>(jc) * rowStride;
    const float *mu = active ? meanRow[ni] + static_cast<size_t>(nj) * rowStride : mc;

    EigenSseSum acc;
    float s = 0.0f;
    const int blocks8 = Dm >> 3;
    const float4 *mc4 = reinterpret_cast<const float4 *>(mc), *mu4 = reinterpret_cast<const float4 *>(mu);
    // software pipeline: the next block's five loads are in flight while the current block is computed
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
    float sgn = 1.0f;
    if (blocks8 > 0)
    {
        a0 = __ldg(mc4);
        a1 = __ldg(mc4 + 1);
        b0 = __ldg(mu4);
        b1 = __ldg(mu4 + 1);
        sgn = __ldg(sg + nb);
    }
#pragma unroll 1
    for (int b8 = 0; b8 < blocks8; ++b8)
    {
        const float mcv[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, muv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float sMown = clamp_sigma(sgn); // this lane's element of the block
        if (b8 + 1 < blocks8)
        {
            a0 = __ldg(mc4 + 2 * b8 + 2);
            a1 = __ldg(mc4 + 2 * b8 + 3);
            b0 = __ldg(mu4 + 2 * b8 + 2);
            b1 = __ldg(mu4 + 2 * b8 + 3);
            sgn = __ldg(sg + 8 * b8 + 8 + nb);
        }
        const float yown = __frcp_rn(sMown);
        float tv[8], sMv[8];
        bool bad = ((__ballot_sync(0xffffffffu, !rcp_ok(sMown)) >> grp) & 0xffu) != 0; // some sM of this node's block is out of range
#pragma unroll
        for (int e = 0; e < 8; ++e)
        {
            sMv[e] = __shfl_sync(0xffffffffu, sMown, grp + e);
            tv[e] = raw_term_fast(mcv[e], muv[e], sMv[e], __shfl_sync(0xffffffffu, yown, grp + e), bad);
        }
        if (bad)
            raw_terms_exact(mcv, muv, sMv, tv);
        if (ORDER == VSOM_ORDER_EIGEN_SSE)
            acc.block(tv);
        else
        {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                s = __fadd_rn(s, tv[e]);
        }
    }
    // the last Dm % 8 elements
    {
        const int k0 = blocks8 << 3, nrest = Dm - k0;
        const float sMown = nb < nrest ? clamp_sigma(__ldg(sg + k0 + nb)) : 1.0f;
        const float yown = __frcp_rn(sMown);
        float tv[8], sMv[8], mcv[8], muv[8];
        bool bad = ((__ballot_sync(0xffffffffu, !rcp_ok(sMown)) >> grp) & 0xffu) != 0;
#pragma unroll
        for (int e = 0; e < 8; ++e)
        {
            sMv[e] = __shfl_sync(0xffffffffu, sMown, grp + e);
            mcv[e] = e < nrest ? __ldg(mc + k0 + e) : 0.0f;
            muv[e] = e < nrest ? __ldg(mu + k0 + e) : 0.0f;
            tv[e] = raw_term_fast(mcv[e], muv[e], sMv[e], __shfl_sync(0xffffffffu, yown, grp + e), bad);
        }
        if (bad)
            raw_terms_exact(mcv, muv, sMv, tv);
        if (ORDER == VSOM_ORDER_EIGEN_SSE)
            s = acc.finish(tv, nrest);
        else
            for (int e = 0; e < nrest; ++e)
                s = __fadd_rn(s, tv[e]);
    }
    res[t][nb] = s;
    __syncthreads();
    if (nb == 0 && node)
    {
        double u = 0.0;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const int qi = i + kDi[q], qj = j + kDj[q];
            if (qi >= 0 && qi < H && qj >= 0 && qj < W)
            {
                const double r = static_cast<double>(res[t][q]);
                u = __dadd_rn(u, q < 4 ? r : __dmul_rn(r, 0.3));
                ++cnt;
            }
        }
        out[static_cast<size_t>(i) * W + j] = __ddiv_rn(u, static_cast<double>(cnt));
    }
}

// ------------------------------------------------------------------------------------------------ tiled version
// The kernel above asks L1 for 32 different row segments per load instruction (one neighbour row per lane): ncu shows the
// L1/TEX data path 99 % busy at 0.13 of the HBM roofline.  Here a CTA owns a tile of UTY x UTX nodes and stages 32-element
// slices of the (UTY + 2) x (UTX + 2) mean rows and UTY x UTX sigma rows it needs into shared memory with coalesced 16-byte
// cp.async copies (one 128-byte row segment per eight lanes), double buffered; the (node, neighbour) threads then read
// their two rows from shared memory (row stride 36 floats: consecutive rows start 16 bytes apart modulo 128).  Neighbour
// rows are fetched once per tile instead of once per neighbour: 1.35 x the algorithmic bytes instead of ~5 x.
// Same arithmetic, same order, same shared reciprocal as above.  Tiles are runs of at most UTY consecutive grid rows that
// the context owns (node-sharded contexts own blocks of rows); every row a tile touches has an entry in the row tables.
constexpr int UTX = 16, UTY = 4, UKS = 32, USTR = 36;
constexpr int UTHREADS = UTX * UTY * 8;
constexpr int UMEANROWS = (UTY + 2) * (UTX + 2), UROWS = UMEANROWS + UTY * UTX;
constexpr int UCHUNKS = (UROWS * 8 + UTHREADS - 1) / UTHREADS; // 16-byte copies per thread and slice

template <int ORDER>
__global__ void __launch_bounds__(UTHREADS) umatrix_tiled_kernel(const float *const *__restrict__ meanRow, const float *const *__restrict__ sigmaRow,
                                                                 const int2 *__restrict__ tiles, int W, int H, int Dm, int rowStride, double *__restrict__ out)
{
    extern __shared__ __align__(16) float usmem[];
    float(*sbuf)[UROWS * USTR] = reinterpret_cast<float(*)[UROWS * USTR]>(usmem); // two slice buffers
    float(*res)[8] = reinterpret_cast<float(*)[8]>(usmem + 2 * UROWS * USTR);     // [UTX * UTY][8]

    const int tid = threadIdx.x, lane = tid & 31;
    const int2 tile = tiles[blockIdx.y]; // first grid row, number of rows (<= UTY)
    const int y0 = tile.x, ny = tile.y, x0 = blockIdx.x * UTX;
    const int nb = tid & 7, tx = (tid >> 3) % UTX, ty = tid / (8 * UTX);
    const int i = y0 + ty, j = x0 + tx, ni = i + kDi[nb], nj = j + kDj[nb];
    const bool node = ty < ny && j < W;
    const bool active = node && ni >= 0 && ni < H && nj >= 0 && nj < W;
    const int grp = lane & ~7;

    // the 16-byte copies this thread issues for every slice: source row (null: zeros) and destination offset
    const float *csrc[UCHUNKS];
    int cdst[UCHUNKS];
#pragma unroll
    for (int c = 0; c < UCHUNKS; ++c)
    {
        const int id = tid + c * UTHREADS, row = id >> 3, part = id & 7;
        csrc[c] = nullptr;
        cdst[c] = row < UROWS ? row * USTR + part * 4 : -1;
        if (row < UMEANROWS)
        {
            const int ry = row / (UTX + 2), rx = row - ry * (UTX + 2), y = y0 - 1 + ry, x = x0 - 1 + rx;
            if (ry < ny + 2 && y >= 0 && y < H && x >= 0 && x < W && meanRow[y])
                csrc[c] = meanRow[y] + static_cast<size_t>(x) * rowStride + part * 4;
        }
        else if (row < UROWS)
        {
            const int r2 = row - UMEANROWS, ry = r2 / UTX, rx = r2 - ry * UTX, y = y0 + ry, x = x0 + rx;
            if (ry < ny && x < W)
                csrc[c] = sigmaRow[y] + static_cast<size_t>(x) * rowStride + part * 4;
        }
    }
    auto load_slice = [&](int sl, float *buf) {
        const int k0 = sl * UKS;
#pragma unroll
        for (int c = 0; c < UCHUNKS; ++c)
            if (cdst[c] >= 0)
            {
                const int part4 = (cdst[c] % USTR); // element offset of the chunk inside the slice
                if (csrc[c] && k0 + part4 < rowStride)
                    cp_async16(buf + cdst[c], csrc[c] + k0);
                else
                    *reinterpret_cast<float4 *>(buf + cdst[c]) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int mcRow = (ty + 1) * (UTX + 2) + tx + 1;
    const int muRow = active ? (ty + 1 + kDi[nb]) * (UTX + 2) + tx + 1 + kDj[nb] : mcRow; // no neighbour: distance to itself (0)
    const int sgRow = UMEANROWS + ty * UTX + tx;

    u64 acc2[4] = {0ull, 0ull, 0ull, 0ull}; // the eight chains of EigenSseSum as four packed pairs (+0.0f each)
    float s = 0.0f;
    const int blocks8 = Dm >> 3, nrest = Dm & 7;
    const int nSlices = (Dm + UKS - 1) / UKS;
    load_slice(0, sbuf[0]);
    for (int sl = 0; sl < nSlices; ++sl)
    {
        if (sl + 1 < nSlices)
        {
            load_slice(sl + 1, sbuf[(sl + 1) & 1]);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float *buf = sbuf[sl & 1];
        const float *mc = buf + mcRow * USTR, *mu = buf + muRow * USTR, *sg = buf + sgRow * USTR;
        const int fullHere = min(UKS / 8, blocks8 - sl * (UKS / 8)); // full blocks of eight inside this slice
#pragma unroll
        for (int b8 = 0; b8 < UKS / 8; ++b8)
            if (b8 < fullHere)
            {
                // elements (0,1) (2,3) (4,5) (6,7) of the block as packed f32x2 pairs: the five rounded operations of a term cost
                // five issue slots per TWO elements (FADD2 / FMUL2 / FFMA2), same bits as the scalar sequence in raw_term_fast
                const ulonglong2 a0 = *reinterpret_cast<const ulonglong2 *>(mc + 8 * b8), a1 = *reinterpret_cast<const ulonglong2 *>(mc + 8 * b8 + 4);
                const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(mu + 8 * b8), b1 = *reinterpret_cast<const ulonglong2 *>(mu + 8 * b8 + 4);
                const u64 mc2[4] = {a0.x, a0.y, a1.x, a1.y}, mu2[4] = {b0.x, b0.y, b1.x, b1.y};
                const float sMown = clamp_sigma(sg[8 * b8 + nb]); // this lane's element of the block
                const float yown = __frcp_rn(sMown), nsMown = -sMown;
                u64 t2[4];
                float nsMv[8];
                bool bad = ((__ballot_sync(0xffffffffu, !rcp_ok(sMown)) >> grp) & 0xffu) != 0;
#pragma unroll
                for (int p = 0; p < 4; ++p)
                {
                    nsMv[2 * p] = __shfl_sync(0xffffffffu, nsMown, grp + 2 * p);
                    nsMv[2 * p + 1] = __shfl_sync(0xffffffffu, nsMown, grp + 2 * p + 1);
                    const u64 y2 = f2_pack(__shfl_sync(0xffffffffu, yown, grp + 2 * p), __shfl_sync(0xffffffffu, yown, grp + 2 * p + 1));
                    const u64 d2 = f2_sub(mc2[p], mu2[p]);
                    const u64 q2 = f2_mul(d2, y2);
                    const u64 r2 = f2_fma(f2_pack(nsMv[2 * p], nsMv[2 * p + 1]), q2, d2);
                    const u64 av = f2_fma(r2, y2, q2);
                    t2[p] = f2_mul_halves(av, av); // feeds the chain's addition: scalar products (see f2_mul_halves)
                    float dl, dh;
                    f2_unpack(d2, dl, dh);
                    bad = bad || !(fabsf(dl) < 1.2676506e30f) || !(fabsf(dh) < 1.2676506e30f); // 2^100; NaN fails the comparison
                }
                if (bad)
                {
                    float mcv[8], muv[8], sMv[8], tv[8];
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                    {
                        f2_unpack(mc2[p], mcv[2 * p], mcv[2 * p + 1]);
                        f2_unpack(mu2[p], muv[2 * p], muv[2 * p + 1]);
                        sMv[2 * p] = -nsMv[2 * p];
                        sMv[2 * p + 1] = -nsMv[2 * p + 1];
                    }
                    raw_terms_exact(mcv, muv, sMv, tv);
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        t2[p] = f2_pack(tv[2 * p], tv[2 * p + 1]);
                }
                if (ORDER == VSOM_ORDER_EIGEN_SSE)
                {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        acc2[p] = f2_add(acc2[p], t2[p]); // EigenSseSum::block on chains (2p, 2p + 1)
                }
                else
                {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                    {
                        float tl, th;
                        f2_unpack(t2[p], tl, th);
                        s = __fadd_rn(__fadd_rn(s, tl), th);
                    }
                }
            }
        if (sl == nSlices - 1)
        {
            // the last Dm % 8 elements sit behind this slice's full blocks (zero padded up to the slice's end)
            const int o = (blocks8 << 3) - sl * UKS; // in [0, 24] when there is a rest
            float tv[8], sMv[8], mcv[8], muv[8];
            const float sMown = (nrest && nb < nrest) ? clamp_sigma(sg[o + nb]) : 1.0f;
            const float yown = __frcp_rn(sMown);
            bool bad = ((__ballot_sync(0xffffffffu, !rcp_ok(sMown)) >> grp) & 0xffu) != 0;
#pragma unroll
            for (int e = 0; e < 8; ++e)
            {
                sMv[e] = __shfl_sync(0xffffffffu, sMown, grp + e);
                mcv[e] = (nrest && e < nrest) ? mc[o + e] : 0.0f;
                muv[e] = (nrest && e < nrest) ? mu[o + e] : 0.0f;
                tv[e] = raw_term_fast(mcv[e], muv[e], sMv[e], __shfl_sync(0xffffffffu, yown, grp + e), bad);
            }
            if (bad)
                raw_terms_exact(mcv, muv, sMv, tv);
            if (ORDER == VSOM_ORDER_EIGEN_SSE)
            {
                EigenSseSum acc;
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    f2_unpack(acc2[p], acc.a[2 * p], acc.a[2 * p + 1]);
                s = acc.finish(tv, nrest);
            }
            else
                for (int e = 0; e < nrest; ++e)
                    s = __fadd_rn(s, tv[e]);
        }
        __syncthreads(); // this buffer is refilled by the next iteration's load
    }
    res[tid >> 3][nb] = s;
    __syncthreads();
    if (nb == 0 && node)
    {
        double u = 0.0;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const int qi = i + kDi[q], qj = j + kDj[q];
            if (qi >= 0 && qi < H && qj >= 0 && qj < W)
            {
                const double r = static_cast<double>(res[tid >> 3][q]);
                u = __dadd_rn(u, q < 4 ? r : __dmul_rn(r, 0.3));
                ++cnt;
            }
        }
        out[static_cast<size_t>(i) * W + j] = __ddiv_rn(u, static_cast<double>(cnt));
    }
}

// ------------------------------------------------------------------------------------------------ tiled, Eigen order, element-major
// In Eigen's packet order the eight chains of a dot product are indexed by the element's position inside its block of eight
// (EigenSseSum).  So lane e of a node's eight lanes can own ELEMENT e of every block for ALL eight neighbours: its own-node
// value, sigma and reciprocal are private (no shuffles at all), each neighbour costs one conflict-free LDS.32, and the lane's
// eight accumulators ARE chain e of the eight pairs.  The kernel above needs 4 LDS.128 + 16 SHFL per block and warp and is bound
// by the shared-memory / shuffle pipe (32 of its cycles per block and warp, measured: halving the FP issue slots changed nothing);
// this one needs 10 LDS.32.  Neighbours are processed as packed f32x2 pairs.  The chains meet once, at the end (finish()).
// The sequential order has no such decomposition (its single chain runs across the elements) and keeps the kernel above.
constexpr int UKSE = 64;  // elements per slice (two barriers and one round of copy bookkeeping per slice)
constexpr int USTRE = UKSE + 8; // row stride: rows of consecutive nodes 8 banks apart -> the 4 x 8 lanes of a warp hit 32 different banks
constexpr int UCHUNKSE = (UROWS * (UKSE / 4) + UTHREADS - 1) / UTHREADS; // 16-byte copies per thread and slice

__global__ void __launch_bounds__(UTHREADS, 2) umatrix_tiled_eigen_kernel(const float *const *__restrict__ meanRow, const float *const *__restrict__ sigmaRow,
                                                                       const int2 *__restrict__ tiles, int W, int H, int Dm, int rowStride, double *__restrict__ out)
{
    extern __shared__ __align__(16) float usmem[];
    float(*sbuf)[UROWS * USTRE] = reinterpret_cast<float(*)[UROWS * USTRE]>(usmem); // two slice buffers; reused for the chains at the end
    float(*res)[8] = reinterpret_cast<float(*)[8]>(usmem + 2 * UROWS * USTRE);      // [UTX * UTY][8]

    const int tid = threadIdx.x;
    const int2 tile = tiles[blockIdx.y]; // first grid row, number of rows (<= UTY)
    const int y0 = tile.x, ny = tile.y, x0 = blockIdx.x * UTX;
    const int e = tid & 7, tx = (tid >> 3) % UTX, ty = tid / (8 * UTX);
    const int i = y0 + ty, j = x0 + tx;
    const bool node = ty < ny && j < W;

    // per 16-byte copy of a slice: source (null: the chunk is zero-filled), destination offset, and the number of elements the
    // chunk may start at before it leaves the row (<= 0: never copy)
    const float *csrc[UCHUNKSE];
    int cdst[UCHUNKSE], clim[UCHUNKSE];
#pragma unroll
    for (int c = 0; c < UCHUNKSE; ++c)
    {
        const int id = tid + c * UTHREADS, row = id / (UKSE / 4), part = id % (UKSE / 4);
        csrc[c] = nullptr;
        cdst[c] = row < UROWS ? row * USTRE + part * 4 : -1;
        if (row < UMEANROWS)
        {
            const int ry = row / (UTX + 2), rx = row - ry * (UTX + 2), y = y0 - 1 + ry, x = x0 - 1 + rx;
            if (ry < ny + 2 && y >= 0 && y < H && x >= 0 && x < W && meanRow[y])
                csrc[c] = meanRow[y] + static_cast<size_t>(x) * rowStride + part * 4;
        }
        else if (row < UROWS)
        {
            const int r2 = row - UMEANROWS, ry = r2 / UTX, rx = r2 - ry * UTX, y = y0 + ry, x = x0 + rx;
            if (ry < ny && x < W)
                csrc[c] = sigmaRow[y] + static_cast<size_t>(x) * rowStride + part * 4;
        }
        clim[c] = csrc[c] ? rowStride - part * 4 : 0; // copy while k0 < clim (0: no source, or no such chunk)
    }
    // Rows without a source (outside the grid, absent halo) are zeroed ONCE; a chunk past the end of the rows is simply not copied:
    // what it would hold is never read (full blocks end at 8 (Dm / 8), the rest block reads its first Dm % 8 elements).
    for (int q = tid; q < 2 * UROWS * USTRE / 4; q += UTHREADS)
        reinterpret_cast<float4 *>(usmem)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    auto load_slice = [&](int sl, float *buf) {
        const int k0 = sl * UKSE;
#pragma unroll
        for (int c = 0; c < UCHUNKSE; ++c)
            if (k0 < clim[c])
                cp_async16(buf + cdst[c], csrc[c] + k0);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // BYTE offsets of this lane's element inside a slice buffer (the buffer base stays in a uniform register: LDS [R + UR + imm])
    const unsigned mcOff = static_cast<unsigned>(((ty + 1) * (UTX + 2) + tx + 1) * USTRE + e) * 4u;
    const unsigned sgOff = static_cast<unsigned>((UMEANROWS + ty * UTX + tx) * USTRE + e) * 4u;
    unsigned muOff[8];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
    {
        const int ni = i + kDi[nb], nj = j + kDj[nb];
        const bool act = node && ni >= 0 && ni < H && nj >= 0 && nj < W;
        muOff[nb] = act ? static_cast<unsigned>(((ty + 1 + kDi[nb]) * (UTX + 2) + tx + 1 + kDj[nb]) * USTRE + e) * 4u : mcOff; // no neighbour: distance to itself (0)
    }

    // one block of eight elements (this lane: element e) against the eight neighbours -> four packed terms.  The pointers already
    // include the slice buffer, the row and e; `o` is a compile-time offset in the unrolled callers (LDS [reg + imm]).
    auto at = [](const float *buf, unsigned byteOff, int o) { return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(buf) + byteOff + 4 * o); };
    auto block_terms = [&](const float *buf, int o, u64 (&t2)[4]) {
        const float mcv = at(buf, mcOff, o), sM = clamp_sigma(at(buf, sgOff, o));
        // sM in [1e-5, 2^100]: inside the range where __frcp_rn is its fast path (MUFU.RCP + one FMA-residual Newton step, the
        // correctly rounded reciprocal for biased exponents 1..252); written out to drop the range test and the slow-path call.
        // Outside (NaN, > 2^100) the block takes the exact path below and y is not used.
        float y0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(sM));
        const float y = __fmaf_rn(y0, -__fmaf_rn(sM, y0, -1.0f), y0);
        const u64 mc2 = f2_pack(mcv, mcv), y2 = f2_pack(y, y), nsM2 = f2_pack(-sM, -sM);
        bool bad = !rcp_ok(sM);
        float muv[8];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
            muv[nb] = at(buf, muOff[nb], o);
#pragma unroll
        for (int p = 0; p < 4; ++p)
        {
            const u64 d2 = f2_sub(mc2, f2_pack(muv[2 * p], muv[2 * p + 1]));
            const u64 q2 = f2_mul(d2, y2);
            const u64 r2 = f2_fma(nsM2, q2, d2);
            const u64 av = f2_fma(r2, y2, q2);
            t2[p] = f2_mul_halves(av, av); // feeds the chain's addition: scalar products (see f2_mul_halves)
            float dl, dh;
            f2_unpack(d2, dl, dh);
            bad = bad || !(fabsf(dl) < 1.2676506e30f) || !(fabsf(dh) < 1.2676506e30f); // 2^100; NaN fails the comparison
        }
        if (bad) // the reference's own arithmetic for what the correction step is not proven for (rare, per lane)
        {
#pragma unroll
            for (int p = 0; p < 4; ++p)
            {
                const float a0 = __fdiv_rn(__fsub_rn(mcv, muv[2 * p]), sM), a1 = __fdiv_rn(__fsub_rn(mcv, muv[2 * p + 1]), sM);
                t2[p] = f2_pack(__fmul_rn(a0, a0), __fmul_rn(a1, a1));
            }
        }
    };

    u64 acc2[4] = {0ull, 0ull, 0ull, 0ull}; // chain e of the pairs (0,1) (2,3) (4,5) (6,7)
    u64 rest2[4] = {0ull, 0ull, 0ull, 0ull}; // term of element e of the trailing partial block (e < Dm % 8)
    const int blocks8 = Dm >> 3, nrest = Dm & 7;
    const int nSlices = (Dm + UKSE - 1) / UKSE;
    load_slice(0, sbuf[0]);
    for (int sl = 0; sl < nSlices; ++sl)
    {
        if (sl + 1 < nSlices)
        {
            load_slice(sl + 1, sbuf[(sl + 1) & 1]);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        }
        else
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float *buf = sbuf[sl & 1];
        const int fullHere = min(UKSE / 8, blocks8 - sl * (UKSE / 8)); // full blocks of eight inside this slice
#pragma unroll
        for (int b8 = 0; b8 < UKSE / 8; ++b8)
            if (b8 < fullHere)
            {
                u64 t2[4];
                block_terms(buf, 8 * b8, t2);
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    acc2[p] = f2_add(acc2[p], t2[p]);
            }
        if (sl == nSlices - 1 && e < nrest)
            block_terms(buf, (blocks8 << 3) - sl * UKSE, rest2); // the last Dm % 8 elements sit behind this slice's full blocks
        __syncthreads(); // this buffer is refilled by the next iteration's load
    }
    // the chains meet: [node][neighbour][chain] and [node][neighbour][rest element] through shared memory, then finish() per pair
    float *fin = usmem, *finRest = usmem + UTX * UTY * 64;
    {
        const int nodeAt = (tid >> 3) * 64;
#pragma unroll
        for (int p = 0; p < 4; ++p)
        {
            float lo, hi;
            f2_unpack(acc2[p], lo, hi);
            fin[nodeAt + (2 * p) * 8 + e] = lo;
            fin[nodeAt + (2 * p + 1) * 8 + e] = hi;
            f2_unpack(rest2[p], lo, hi);
            finRest[nodeAt + (2 * p) * 8 + e] = lo;
            finRest[nodeAt + (2 * p + 1) * 8 + e] = hi;
        }
    }
    __syncthreads();
    {
        // this thread: pair (node, neighbour e)
        EigenSseSum acc;
        float rest[8];
        const int at = (tid >> 3) * 64 + e * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c)
        {
            acc.a[c] = fin[at + c];
            rest[c] = finRest[at + c];
        }
        const float sres = acc.finish(rest, nrest);
        __syncthreads(); // res lies behind the slice buffers, but keep the phases apart anyway
        res[tid >> 3][e] = sres;
    }
    __syncthreads();
    if (e == 0 && node)
    {
        double u = 0.0;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const int qi = i + kDi[q], qj = j + kDj[q];
            if (qi >= 0 && qi < H && qj >= 0 && qj < W)
            {
                const double r = static_cast<double>(res[tid >> 3][q]);
                u = __dadd_rn(u, q < 4 ? r : __dmul_rn(r, 0.3));
                ++cnt;
            }
        }
        out[static_cast<size_t>(i) * W + j] = __ddiv_rn(u, static_cast<double>(cnt));
    }
}

int launch_umatrix_tiles(vsom_ctx *ctx, const float *const *meanRowDev, const float *const *sigmaRowDev, const int2 *tilesDev, int nTiles)
{
    if (ctx->W < 2 || ctx->H < 2)
        return set_error(ctx, VSOM_ERR_INVALID, "updateUMatrix needs width >= 2 and height >= 2 (the reference indexes out of bounds otherwise)");
    if (nTiles <= 0)
        return VSOM_OK;
    dim3 grid((ctx->W + UTX - 1) / UTX, nTiles);
    const int smem = static_cast<int>(sizeof(float) * (2 * UROWS * USTR + UTX * UTY * 8));
    static const bool lanesPerPair = [] { const char *e = getenv("VSOM_UMATRIX_EIGEN_KERNEL"); return e && atoi(e) == 0; }(); // 0: the (node, neighbour) kernel also in Eigen order
    if (ctx->order == VSOM_ORDER_EIGEN_SSE && !lanesPerPair)
    {
        const int smemE = static_cast<int>(sizeof(float) * (2 * UROWS * USTRE + UTX * UTY * 8));
        static_assert(2 * UROWS * USTRE >= 2 * UTX * UTY * 64, "the chains are exchanged through the slice buffers");
        VSOM_CUDA(ctx, cudaFuncSetAttribute(umatrix_tiled_eigen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemE));
        umatrix_tiled_eigen_kernel<<<grid, UTHREADS, smemE, ctx->stream>>>(meanRowDev, sigmaRowDev, tilesDev, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    }
    else if (ctx->order == VSOM_ORDER_EIGEN_SSE)
    {
        VSOM_CUDA(ctx, cudaFuncSetAttribute(umatrix_tiled_kernel<VSOM_ORDER_EIGEN_SSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umatrix_tiled_kernel<VSOM_ORDER_EIGEN_SSE><<<grid, UTHREADS, smem, ctx->stream>>>(meanRowDev, sigmaRowDev, tilesDev, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    }
    else
    {
        VSOM_CUDA(ctx, cudaFuncSetAttribute(umatrix_tiled_kernel<VSOM_ORDER_REFERENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        umatrix_tiled_kernel<VSOM_ORDER_REFERENCE><<<grid, UTHREADS, smem, ctx->stream>>>(meanRowDev, sigmaRowDev, tilesDev, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    }
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

int umatrix_tile_rows() { return UTY; }

// meanRowDev / sigmaRowDev: [H] device pointers to the first node of every grid row this context can read (own rows, and
// for node-sharded contexts the halo rows; sigma only for the own rows); rowsDev: the nRows grid rows to compute.
// The result lands in ctx->umatrix at the node's GLOBAL index.
int launch_umatrix_rows(vsom_ctx *ctx, const float *const *meanRowDev, const float *const *sigmaRowDev, const int *rowsDev, int nRows)
{
    if (ctx->W < 2 || ctx->H < 2)
        return set_error(ctx, VSOM_ERR_INVALID, "updateUMatrix needs width >= 2 and height >= 2 (the reference indexes out of bounds otherwise)");
    if (nRows <= 0)
        return VSOM_OK;
    dim3 grid((ctx->W + UT - 1) / UT, nRows);
    if (ctx->order == VSOM_ORDER_EIGEN_SSE)
        umatrix_kernel<VSOM_ORDER_EIGEN_SSE><<<grid, UT * 8, 0, ctx->stream>>>(meanRowDev, sigmaRowDev, rowsDev, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    else
        umatrix_kernel<VSOM_ORDER_REFERENCE><<<grid, UT * 8, 0, ctx->stream>>>(meanRowDev, sigmaRowDev, rowsDev, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

} // namespace vsom

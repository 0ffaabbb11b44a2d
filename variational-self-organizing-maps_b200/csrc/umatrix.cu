// umatrix.cu — K4: the U-matrix stencil.
//
// Replaces Som::updateUMatrix (src/Som.cpp:999-1111): for every node the sigma-weighted distance
// Som::euclidianWeightedDistRaw (src/Som.cpp:143-157) to each existing neighbour of its 8-neighbourhood,
// straight neighbours weighted 1, diagonal ones 0.3, summed left to right in double in the textual order of
// the source and divided by 8 / 5 / 3 (interior / edge / corner).  Raw(p, u) always divides by p's OWN
// clamped sigma, so the stencil is not symmetric.
//
// One thread owns one (node, neighbour) pair and sums its Dm terms sequentially in f32 — the order of the
// reference's dot product — so every Raw value and every U value is bit-identical to the reference.  A CTA
// covers 16 consecutive nodes of one grid row.  The rows are read straight from global memory with 128-bit
// read-only loads: a node's eight threads ask for the same words of its own mean and sigma rows (one
// transaction), neighbouring nodes share neighbour rows, and L1 serves the reuse.  (A first version staged 32-wide
// slices of three grid rows through shared memory; ncu showed 54 % of its instructions in the staging loop's
// address arithmetic, issue slots 76 % busy: 2.78 ms at 512x512x784.)
#include "common.cuh"

namespace vsom
{

constexpr int UT = 16; // nodes per CTA

// neighbour order of the reference's sums: W, E, S, N, then the diagonals NW, SW, NE, SE
// (src/Som.cpp:1017-1025; the edge and corner cases drop the missing terms and keep this order).
__constant__ int kDi[8] = {0, 0, +1, -1, -1, +1, -1, +1};
__constant__ int kDj[8] = {-1, +1, 0, 0, -1, -1, +1, +1};

// one term of Raw(p, u): a = (m_p - m_u) / sM with sM = max(sigma_p, 1e-5) (:150), accumulated as a * a (:156).
// (m - v) * (valid * weights) / sM equals a bit for bit because valid * weights == 1.0f exactly (:1002-1003), so one
// IEEE division serves both factors of the dot product.
__device__ __forceinline__ void raw_term(float &s, float mc, float mu, float sg)
{
    const float sM = sg < 0.00001f ? 0.00001f : sg;
    const float a = __fdiv_rn(__fsub_rn(mc, mu), sM);
    s = __fadd_rn(s, __fmul_rn(a, a));
}

__global__ void __launch_bounds__(UT * 8) umatrix_kernel(const float *__restrict__ mean, const float *__restrict__ sigma, int W, int H, int Dm,
                                                         int rowStride, double *__restrict__ out)
{
    __shared__ float res[UT][8];

    const int tid = threadIdx.x;
    const int i = blockIdx.y;       // grid row
    const int j0 = blockIdx.x * UT; // first grid column of the tile
    const int t = tid >> 3, nb = tid & 7;
    const int j = j0 + t;
    const int ni = i + kDi[nb], nj = j + kDj[nb];
    const bool active = j < W && ni >= 0 && ni < H && nj >= 0 && nj < W;

    float s = 0.0f;
    if (active)
    {
        const float *mc = mean + (static_cast<size_t>(i) * W + j) * rowStride;
        const float *sg = sigma + (static_cast<size_t>(i) * W + j) * rowStride;
        const float *mu = mean + (static_cast<size_t>(ni) * W + nj) * rowStride;
        const int quads = Dm >> 2; // rows start 16-byte aligned (rowStride is a multiple of 4 floats)
        const float4 *mc4 = reinterpret_cast<const float4 *>(mc), *sg4 = reinterpret_cast<const float4 *>(sg), *mu4 = reinterpret_cast<const float4 *>(mu);
#pragma unroll 2
        for (int q = 0; q < quads; ++q)
        {
            const float4 a = __ldg(mc4 + q), b = __ldg(mu4 + q), c = __ldg(sg4 + q);
            raw_term(s, a.x, b.x, c.x);
            raw_term(s, a.y, b.y, c.y);
            raw_term(s, a.z, b.z, c.z);
            raw_term(s, a.w, b.w, c.w);
        }
        for (int k = quads << 2; k < Dm; ++k)
            raw_term(s, __ldg(mc + k), __ldg(mu + k), __ldg(sg + k));
    }
    res[t][nb] = s;
    __syncthreads();
    if (nb == 0 && j < W)
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

int launch_umatrix(vsom_ctx *ctx)
{
    if (ctx->W < 2 || ctx->H < 2)
        return set_error(ctx, VSOM_ERR_INVALID, "updateUMatrix needs width >= 2 and height >= 2 (the reference indexes out of bounds otherwise)");
    dim3 grid((ctx->W + UT - 1) / UT, ctx->H);
    umatrix_kernel<<<grid, UT * 8, 0, ctx->stream>>>(ctx->mean, ctx->sigma, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

} // namespace vsom

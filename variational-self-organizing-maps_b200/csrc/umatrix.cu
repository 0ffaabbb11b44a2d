// umatrix.cu — K4: the U-matrix stencil.
//
// Replaces Som::updateUMatrix (src/Som.cpp:999-1111): for every node the sigma-weighted distance
// Som::euclidianWeightedDistRaw (src/Som.cpp:143-157) to each existing neighbour of its 8-neighbourhood,
// straight neighbours weighted 1, diagonal ones 0.3, summed left to right in double in the textual order of
// the source and divided by 8 / 5 / 3 (interior / edge / corner).  Raw(p, u) always divides by p's OWN
// clamped sigma, so the stencil is not symmetric.
//
// One thread owns one node and carries its eight Raw(p, neighbour) sums as eight independent sequential f32 chains
// (the order of the reference's dot product, so every Raw value and every U value is bit-identical to the reference;
// the eight IEEE divisions per slice element overlap instead of waiting for each other).  A CTA covers an 8 x 16 block of
// nodes; the block plus its one-node halo (10 x 18 rows of the mean plane) and the block's sigma rows are staged through
// shared memory in 32-wide slices of the model vector (coalesced 128-byte segments, halo re-read factor 1.4).
#include "common.cuh"

namespace vsom
{

constexpr int UTX = 16, UTY = 8; // nodes per CTA: 16 columns x 8 rows
constexpr int UKC = 32;          // slice of the model vector
constexpr int UHX = UTX + 2, UHY = UTY + 2;

// neighbour order of the reference's sums: W, E, S, N, then the diagonals NW, SW, NE, SE
// (src/Som.cpp:1017-1025; the edge and corner cases drop the missing terms and keep this order).
__constant__ int kDi[8] = {0, 0, +1, -1, -1, +1, -1, +1};
__constant__ int kDj[8] = {-1, +1, 0, 0, -1, -1, +1, +1};

__global__ void __launch_bounds__(UTX * UTY) umatrix_kernel(const float *__restrict__ mean, const float *__restrict__ sigma, int W, int H, int Dm,
                                                            int rowStride, double *__restrict__ out)
{
    __shared__ float mt[UHY * UHX][UKC + 1];
    __shared__ float st[UTY * UTX][UKC + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % UTX, ty = tid / UTX;
    const int i0 = blockIdx.y * UTY, j0 = blockIdx.x * UTX;
    const int i = i0 + ty, j = j0 + tx;
    const int centre = (ty + 1) * UHX + tx + 1;

    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
        s[q] = 0.0f;

    for (int k0 = 0; k0 < Dm; k0 += UKC)
    {
        __syncthreads();
        const int k = k0 + lane;
        for (int item = warp; item < UHY * UHX + UTY * UTX; item += UTX * UTY / 32)
        {
            if (item < UHY * UHX)
            {
                const int gi = i0 + item / UHX - 1, gj = j0 + item % UHX - 1;
                float v = 0.0f;
                if (gi >= 0 && gi < H && gj >= 0 && gj < W && k < Dm)
                    v = mean[(static_cast<size_t>(gi) * W + gj) * rowStride + k];
                mt[item][lane] = v;
            }
            else
            {
                const int c = item - UHY * UHX;
                const int gi = i0 + c / UTX, gj = j0 + c % UTX;
                float v = 1.0f;
                if (gi < H && gj < W && k < Dm)
                    v = sigma[(static_cast<size_t>(gi) * W + gj) * rowStride + k];
                st[c][lane] = v;
            }
        }
        __syncthreads();
        const int kend = Dm - k0 < UKC ? Dm - k0 : UKC;
        const float *mc = mt[centre];
        const float *sg = st[tid];
        for (int kk = 0; kk < kend; ++kk)
        {
            const float sM = sg[kk] < 0.00001f ? 0.00001f : sg[kk]; // :150
            const float c = mc[kk];
#pragma unroll
            for (int q = 0; q < 8; ++q)
            {
                // a = (m - v) / sM and b = ((m - v) * valid*weights) / sM with valid*weights == 1.0f exactly (:1002-1003):
                // d * 1.0f == d bit for bit, so b == a and one IEEE division serves both factors of the dot product.
                // Neighbours outside the grid read zeros here and are dropped in the combination below.
                const float d = __fsub_rn(c, mt[centre + kDi[q] * UHX + kDj[q]][kk]);
                const float a = __fdiv_rn(d, sM);
                s[q] = __fadd_rn(s[q], __fmul_rn(a, a)); // :156
            }
        }
    }
    if (i < H && j < W)
    {
        double u = 0.0;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const int qi = i + kDi[q], qj = j + kDj[q];
            if (qi >= 0 && qi < H && qj >= 0 && qj < W)
            {
                const double r = static_cast<double>(s[q]);
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
    dim3 grid((ctx->W + UTX - 1) / UTX, (ctx->H + UTY - 1) / UTY);
    umatrix_kernel<<<grid, UTX * UTY, 0, ctx->stream>>>(ctx->mean, ctx->sigma, ctx->W, ctx->H, ctx->Dm, ctx->rowStride, ctx->umatrix);
    VSOM_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VSOM_OK;
}

} // namespace vsom

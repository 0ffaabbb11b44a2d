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
// covers 16 consecutive nodes of one grid row; the three grid rows it touches are staged through shared
// memory in 32-wide slices of the model vector (coalesced 128-byte segments).
#include "common.cuh"

namespace vsom
{

constexpr int UT = 16;  // nodes per CTA
constexpr int UKC = 32; // slice of the model vector

// neighbour order of the reference's sums: W, E, S, N, then the diagonals NW, SW, NE, SE
// (src/Som.cpp:1017-1025; the edge and corner cases drop the missing terms and keep this order).
__constant__ int kDi[8] = {0, 0, +1, -1, -1, +1, -1, +1};
__constant__ int kDj[8] = {-1, +1, 0, 0, -1, -1, +1, +1};

__global__ void __launch_bounds__(UT * 8) umatrix_kernel(const float *__restrict__ mean, const float *__restrict__ sigma, int W, int H, int Dm,
                                                         int rowStride, double *__restrict__ out)
{
    __shared__ float mt[3][UT + 2][UKC + 1];
    __shared__ float st[UT][UKC + 1];
    __shared__ float res[UT][8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.y;       // grid row
    const int j0 = blockIdx.x * UT; // first grid column of the tile
    const int t = tid >> 3, nb = tid & 7;
    const int j = j0 + t;
    const int ni = i + kDi[nb], nj = j + kDj[nb];
    const bool active = j < W && ni >= 0 && ni < H && nj >= 0 && nj < W;

    float s = 0.0f;
    for (int k0 = 0; k0 < Dm; k0 += UKC)
    {
        __syncthreads();
        const int k = k0 + lane;
        for (int item = warp; item < 3 * (UT + 2) + UT; item += UT * 8 / 32)
        {
            if (item < 3 * (UT + 2))
            {
                const int r = item / (UT + 2), c = item % (UT + 2);
                const int gi = i + r - 1, gj = j0 + c - 1;
                float v = 0.0f;
                if (gi >= 0 && gi < H && gj >= 0 && gj < W && k < Dm)
                    v = mean[(static_cast<size_t>(gi) * W + gj) * rowStride + k];
                mt[r][c][lane] = v;
            }
            else
            {
                const int c = item - 3 * (UT + 2);
                const int gj = j0 + c;
                float v = 1.0f;
                if (gj < W && k < Dm)
                    v = sigma[(static_cast<size_t>(i) * W + gj) * rowStride + k];
                st[c][lane] = v;
            }
        }
        __syncthreads();
        if (active)
        {
            const int kend = Dm - k0 < UKC ? Dm - k0 : UKC;
            const float *mc = mt[1][t + 1];
            const float *mu = mt[1 + kDi[nb]][t + 1 + kDj[nb]];
            const float *sg = st[t];
            for (int kk = 0; kk < kend; ++kk)
            {
                const float sM = sg[kk] < 0.00001f ? 0.00001f : sg[kk];  // :150
                const float d = __fsub_rn(mc[kk], mu[kk]);
                // a = (m - v) / sM and b = ((m - v) * valid*weights) / sM with valid*weights == 1.0f exactly (:1002-1003):
                // d * 1.0f == d bit for bit, so b == a and one IEEE division serves both factors of the dot product
                const float a = __fdiv_rn(d, sM);
                s = __fadd_rn(s, __fmul_rn(a, a));                         // :156
            }
        }
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

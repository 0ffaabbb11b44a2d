/* TEST INFRASTRUCTURE — CPU restatement of the reference VSOM hot path.  See vsom_oracle.h for the
 * usage rules and the pin status.  Compiled with -msse2 -ffp-contract=off: every float operation
 * below is one IEEE binary32 operation, in the order the reference (with sequentially-summed dot
 * products) performs it.  Citations are file:line in /root/reference.
 */
#include "vsom_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- construction / state */

static int model_length(int Din, int transform)
{
    /* Transformation::Length — src/Transformation.cpp:31-35 (Standard), :69-73 (Median), :162-165 (CLR: J*(J-1)) */
    return transform == ORACLE_CLR ? Din * (Din - 1) : Din;
}

vsom_oracle *oracle_create(int W, int H, int Din, int transform)
{
    /* Som::Construct — src/Som.cpp:11-48: all planes zero. */
    vsom_oracle *o = (vsom_oracle *)calloc(1, sizeof(*o));
    o->W = W;
    o->H = H;
    o->N = W * H;
    o->Din = Din;
    o->transform = transform;
    o->Dm = model_length(Din, transform);
    o->P = transform == ORACLE_CLR ? o->Dm / 2 : 0;
    size_t nd = (size_t)o->N * (size_t)o->Dm;
    o->mean = (float *)calloc(nd ? nd : 1, sizeof(float));
    o->S = (float *)calloc(nd ? nd : 1, sizeof(float));
    o->sigma = (float *)calloc(nd ? nd : 1, sizeof(float));
    o->weight = (float *)calloc((size_t)o->N, sizeof(float));
    o->hits = (uint64_t *)calloc((size_t)o->N, sizeof(uint64_t));
    o->umatrix = (double *)calloc((size_t)o->N, sizeof(double));
    return o;
}

void oracle_destroy(vsom_oracle *o)
{
    if (!o)
        return;
    free(o->mean);
    free(o->S);
    free(o->sigma);
    free(o->weight);
    free(o->hits);
    free(o->umatrix);
    free(o);
}

int oracle_depth(const vsom_oracle *o) { return o->Dm; }

void oracle_random_initialize(vsom_oracle *o, int seed, float sigma)
{
    /* Som::randomInitialize — src/Som.cpp:977-997: srand(seed); node-major then dimension;
     * map = ((float)(rand() % (int)(2000*sigma)) - 1000.f*sigma) / 1000.f ; everything else zero. */
    srand((unsigned)seed);
    for (int i = 0; i < o->N; ++i)
    {
        for (int n = 0; n < o->Dm; ++n)
        {
            size_t k = (size_t)i * (size_t)o->Dm + (size_t)n;
            o->mean[k] = ((float)(rand() % (int)(2000 * sigma)) - (1000.f * sigma)) / 1000.f;
            o->sigma[k] = 0.0f;
            o->S[k] = 0.0f;
        }
        o->weight[i] = 0.0f;
        o->hits[i] = 0;
        o->umatrix[i] = 0.0;
    }
}

void oracle_get_state(const vsom_oracle *o, float *mean, float *S, float *sigma, float *weight, uint64_t *hits)
{
    size_t nd = (size_t)o->N * (size_t)o->Dm;
    if (mean)
        memcpy(mean, o->mean, nd * sizeof(float));
    if (S)
        memcpy(S, o->S, nd * sizeof(float));
    if (sigma)
        memcpy(sigma, o->sigma, nd * sizeof(float));
    if (weight)
        memcpy(weight, o->weight, (size_t)o->N * sizeof(float));
    if (hits)
        memcpy(hits, o->hits, (size_t)o->N * sizeof(uint64_t));
}

void oracle_set_state(vsom_oracle *o, const float *mean, const float *S, const float *sigma, const float *weight, const uint64_t *hits)
{
    size_t nd = (size_t)o->N * (size_t)o->Dm;
    if (mean)
        memcpy(o->mean, mean, nd * sizeof(float));
    if (S)
        memcpy(o->S, S, nd * sizeof(float));
    if (sigma)
        memcpy(o->sigma, sigma, nd * sizeof(float));
    if (weight)
        memcpy(o->weight, weight, (size_t)o->N * sizeof(float));
    if (hits)
        memcpy(o->hits, hits, (size_t)o->N * sizeof(uint64_t));
}

/* ---------------------------------------------------------------- Transformation functors */

/* Comparer — residual r of length Dr (= Dm for Standard/Median, P for CLR), written to out.
 * Standard / Median: model - value (src/Transformation.cpp:7-8, :45-46).
 * CLR: pairs (i<j) in row-major upper-triangle order, x' = v_i, y' = v_j, A = model[0..P), B = model[P..2P);
 *      r = (A*x' + B) - y'  (src/Transformation.cpp:87-104).  `dispersion` and `valueWeight` are ignored
 *      by all three shipped Comparers. */
static int comparer(const vsom_oracle *o, const float *v, const float *m, float *out)
{
    if (o->transform != ORACLE_CLR)
    {
        for (int k = 0; k < o->Dm; ++k)
            out[k] = m[k] - v[k];
        return o->Dm;
    }
    const int P = o->P;
    int q = 0;
    for (int i = 0; i < o->Din; ++i)
        for (int j = i + 1; j < o->Din; ++j)
        {
            float ax = m[q] * v[i];
            float axb = ax + m[P + q];
            out[q] = axb - v[j];
            ++q;
        }
    return P;
}

/* Stepper — delta of length Dm.
 * Standard: value - model (src/Transformation.cpp:11-12).  Median: sign(value - model) (:49-50).
 * CLR: inner = (A*x' + B) - y'; delta = [ (-2*inner)*x' || -2*inner ] (src/Transformation.cpp:107-142;
 *      `-2*inner.array()*xPrime.array()` groups as ((-2)*inner)*x'). */
static void stepper(const vsom_oracle *o, const float *v, const float *m, float *out)
{
    if (o->transform == ORACLE_STANDARD)
    {
        for (int k = 0; k < o->Dm; ++k)
            out[k] = v[k] - m[k];
    }
    else if (o->transform == ORACLE_MEDIAN)
    {
        for (int k = 0; k < o->Dm; ++k)
        {
            float d = v[k] - m[k];
            out[k] = (d != d) ? d : (float)((0.0f < d) - (d < 0.0f));
        }
    }
    else
    {
        const int P = o->P;
        int q = 0;
        for (int i = 0; i < o->Din; ++i)
            for (int j = i + 1; j < o->Din; ++j)
            {
                float ax = m[q] * v[i];
                float axb = ax + m[P + q];
                float inner = axb - v[j];
                float m2 = -2.0f * inner;
                out[q] = m2 * v[i];
                out[P + q] = m2;
                ++q;
            }
    }
}

/* ---------------------------------------------------------------- summation order of dot / squaredNorm */

void oracle_set_order(vsom_oracle *o, int order) { o->order = order; }
int oracle_get_order(const vsom_oracle *o) { return o->order; }

/* The one order-dependent operation of the path: the f32 sum of the terms of `a.dot(b)` / `squaredNorm()`
 * (src/Som.cpp:140, :156, :781, :804, :1167).
 *   ORACLE_ORDER_SEQUENTIAL: t[0] + t[1] + ... left to right — what the reference does when compiled against this
 *     repository's stand-in Eigen header (no Eigen on disk here).
 *   ORACLE_ORDER_EIGEN_SSE: what real Eigen 3.3 / 3.4 does under the reference's release flags (-msse2 only,
 *     build/Makefile:18; README.md:64-69 installs libeigen3-dev).  Restated from Eigen's published sources, which are
 *     NOT in /root/reference (system package, unpinned): Eigen/src/Core/Redux.h, redux_impl<Func, Evaluator,
 *     LinearVectorizedTraversal, NoUnrolling>::run with PacketType = Packet4f, and predux<Packet4f> of
 *     Eigen/src/Core/arch/SSE/PacketMath.h (the non-SSE3 branch).  The reduced object is an expression
 *     (cwiseProduct / cwiseAbs2), so the aligned start is 0.  In words: eight interleaved chains — lanes 0..3 of the even
 *     packets and lanes 0..3 of the odd packets — over the first size/8*8 elements; the two 4-wide accumulators are added
 *     lane by lane; if 4..7 elements remain, one more packet is added; the four lanes are combined as
 *     (l0 + l2) + (l1 + l3); the last size%4 elements are added one by one. */
static float sum_terms(int order, const float *t, int n)
{
    if (order == ORACLE_ORDER_EIGEN_SSE && n >= 4)
    {
        const int aligned2 = n / 8 * 8, aligned = n / 4 * 4;
        float p0[4], p1[4];
        for (int j = 0; j < 4; ++j)
            p0[j] = t[j];
        if (aligned > 4)
        {
            for (int j = 0; j < 4; ++j)
                p1[j] = t[4 + j];
            for (int i = 8; i < aligned2; i += 8)
                for (int j = 0; j < 4; ++j)
                {
                    p0[j] = p0[j] + t[i + j];
                    p1[j] = p1[j] + t[i + 4 + j];
                }
            for (int j = 0; j < 4; ++j)
                p0[j] = p0[j] + p1[j];
            if (aligned > aligned2)
                for (int j = 0; j < 4; ++j)
                    p0[j] = p0[j] + t[aligned2 + j];
        }
        float lo = p0[0] + p0[2], hi = p0[1] + p0[3];
        float res = lo + hi;
        for (int i = aligned; i < n; ++i)
            res = res + t[i];
        return res;
    }
    float s = 0.0f;
    for (int k = 0; k < n; ++k)
        s = s + t[k];
    return s;
}

/* r[k] <- r[k] * r[k] (one rounding each), then the ordered sum */
static float sum_squares(int order, float *r, int n)
{
    for (int k = 0; k < n; ++k)
        r[k] = r[k] * r[k];
    return sum_terms(order, r, n);
}

/* ---------------------------------------------------------------- distances */

double oracle_dist(const vsom_oracle *o, size_t pos, const float *v)
{
    /* Som::euclidianWeightedDist — src/Som.cpp:124-141: builds sM and valid*weights, passes them to the
     * Comparer (which ignores them) and returns comparer.dot(comparer): f32 accumulate, returned as double. */
    float *r = (float *)malloc(sizeof(float) * (size_t)(o->Dm > 0 ? o->Dm : 1));
    int n = comparer(o, v, o->mean + pos * (size_t)o->Dm, r);
    float s = sum_squares(o->order, r, n);
    free(r);
    return (double)s;
}

double oracle_dist_f64(const vsom_oracle *o, size_t pos, const float *v)
{
    /* Same f32 residuals, f64 accumulation: the order-independent value used to classify near-ties
     * (SURVEY.md section 8c). */
    float *r = (float *)malloc(sizeof(float) * (size_t)(o->Dm > 0 ? o->Dm : 1));
    int n = comparer(o, v, o->mean + pos * (size_t)o->Dm, r);
    double s = 0.0;
    for (int k = 0; k < n; ++k)
        s += (double)r[k] * (double)r[k];
    free(r);
    return s;
}

void oracle_all_dists(const vsom_oracle *o, const float *v, double *out)
{
    for (int p = 0; p < o->N; ++p)
        out[p] = oracle_dist(o, (size_t)p, v);
}

double oracle_dist_raw(const vsom_oracle *o, size_t pos, const float *u)
{
    /* Som::euclidianWeightedDistRaw — src/Som.cpp:143-157:
     *   sM = sigma < 1e-5f ? 1e-5f : sigma;  vw = valid*weights (= 1*1 from updateUMatrix :1002-1003)
     *   return ((m - u)/sM) . (((m - u)*vw)/sM)   (f32 dot) */
    const float *m = o->mean + pos * (size_t)o->Dm;
    const float *sg = o->sigma + pos * (size_t)o->Dm;
    float *t = (float *)malloc(sizeof(float) * (size_t)(o->Dm > 0 ? o->Dm : 1));
    for (int k = 0; k < o->Dm; ++k)
    {
        float sM = sg[k] < 0.00001f ? 0.00001f : sg[k];
        float vw = 1.0f * 1.0f;
        float d = m[k] - u[k];
        float a = d / sM;
        float b = (d * vw) / sM;
        t[k] = a * b;
    }
    float s = sum_terms(o->order, t, o->Dm);
    free(t);
    return (double)s;
}

/* ---------------------------------------------------------------- BMU searches */

uint32_t oracle_find_bmu_one(const vsom_oracle *o, const float *v)
{
    /* Som::findBmu — src/Som.cpp:291-309: seed with node 0, strict '<' in double, lowest index wins. */
    double minDist = oracle_dist(o, 0, v);
    size_t minIndex = 0;
    for (size_t i = 0; i < (size_t)o->N; ++i)
    {
        double cur = oracle_dist(o, i, v);
        if (cur < minDist)
        {
            minDist = cur;
            minIndex = i;
        }
    }
    return (uint32_t)minIndex;
}

uint32_t oracle_find_restricted_bmu_one(const vsom_oracle *o, const float *v, uint64_t minHits)
{
    /* Som::findRestrictedBmu — src/Som.cpp:313-332: seeds with node 0 REGARDLESS of its hit count. */
    double minDist = oracle_dist(o, 0, v);
    size_t minIndex = 0;
    for (size_t i = 0; i < (size_t)o->N; ++i)
    {
        double cur = oracle_dist(o, i, v);
        if (cur < minDist && o->hits[i] >= minHits)
        {
            minDist = cur;
            minIndex = i;
        }
    }
    return (uint32_t)minIndex;
}

static uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static uint64_t u64max(uint64_t a, uint64_t b) { return a > b ? a : b; }

uint32_t oracle_find_local_bmu(const vsom_oracle *o, const float *v, uint64_t start)
{
    /* Som::findLocalBmu — src/Som.cpp:335-454, with its size_t arithmetic kept: the "-1" offsets are
     * 2^64-1, min(x+off, W-1) then max(.,0) => at x==0 the left neighbour is x==W-1 (same for y);
     * the Y-direction continuation block never measures anything (startX = -1uz makes the loop empty). */
    const uint64_t W = (uint64_t)o->W, H = (uint64_t)o->H;
    uint64_t lastBMU = start;
    double minDist = oracle_dist(o, (size_t)lastBMU, v);
    uint64_t minIndex = lastBMU;
    const uint64_t M1 = (uint64_t)-1;
    const uint64_t fx[8] = {M1, 0, 1, 1, 1, 0, M1, M1};
    const uint64_t fy[8] = {1, 1, 1, 0, M1, M1, M1, 0};
    uint64_t lastMeasured = lastBMU;
    for (;;)
    {
        uint64_t lmX = lastMeasured % W, lmY = lastMeasured / W;
        uint64_t lbX = lastBMU % W, lbY = lastBMU / W;
        if (lastMeasured == lastBMU)
        {
            for (int i = 0; i < 8; ++i)
            {
                uint64_t cx = u64max(u64min(lmX + fx[i], W - 1), 0);
                uint64_t cy = u64max(u64min(lmY + fy[i], H - 1), 0);
                double cur = oracle_dist(o, (size_t)(cy * W + cx), v);
                if (cur < minDist)
                {
                    minDist = cur;
                    minIndex = cy * W + cx;
                }
            }
            if (minIndex == lastBMU)
                return (uint32_t)minIndex;
            lastMeasured = minIndex;
        }
        else
        {
            if (lmX - lbX) /* moving in X */
            {
                for (int i = -1; i < 2; ++i)
                {
                    uint64_t cx = u64max(u64min(lmX + lmX - lbX, W - 1), 0);
                    uint64_t cy = u64max(u64min(lmY + (uint64_t)(int64_t)i, H - 1), 0);
                    double cur = oracle_dist(o, (size_t)(cy * W + cx), v);
                    if (cur < minDist)
                    {
                        minDist = cur;
                        minIndex = cy * W + cx;
                    }
                }
            }
            if (lmY - lbY) /* moving in Y */
            {
                uint64_t startX, endX;
                if (lmX - lbX > 0)
                {
                    startX = M1;
                    endX = 0;
                }
                else
                {
                    startX = M1;
                    endX = 1;
                }
                for (uint64_t i = startX; i < endX + 1; ++i) /* never entered: startX == 2^64-1 */
                {
                    uint64_t cx = u64max(u64min(lmX + i, W - 1), 0);
                    uint64_t cy = u64max(u64min(lmY + lmY - lbY, H - 1), 0);
                    double cur = oracle_dist(o, (size_t)(cy * W + cx), v);
                    if (cur < minDist)
                    {
                        minDist = cur;
                        minIndex = cy * W + cx;
                    }
                }
            }
            if (minIndex == lastMeasured)
                return (uint32_t)minIndex;
            lastBMU = lastMeasured;
            lastMeasured = minIndex;
        }
    }
}

void oracle_find_bmu(const vsom_oracle *o, const float *x, size_t n, uint32_t *outBmu, float *outDist)
{
    for (size_t r = 0; r < n; ++r)
    {
        const float *v = x + r * (size_t)o->Din;
        uint32_t b = oracle_find_bmu_one(o, v);
        if (outBmu)
            outBmu[r] = b;
        if (outDist)
            outDist[r] = (float)oracle_dist(o, b, v);
    }
}

void oracle_find_restricted_bmu(const vsom_oracle *o, const float *x, size_t n, uint64_t minHits, uint32_t *outBmu)
{
    for (size_t r = 0; r < n; ++r)
        outBmu[r] = oracle_find_restricted_bmu_one(o, x + r * (size_t)o->Din, minHits);
}

void oracle_find_restricted_bmd(const vsom_oracle *o, const float *v, uint64_t minHits, double *out)
{
    /* Som::findRestrictedBmd — src/Som.cpp:457-487: exp(-d*d/2) on the (already squared) distance,
     * normalised by the sum C (no guard for C == 0). */
    double C = 0;
    for (int i = 0; i < o->N; ++i)
    {
        if (o->hits[i] >= minHits)
        {
            double d = oracle_dist(o, (size_t)i, v);
            out[i] = exp(-d * d / 2);
            C += out[i];
        }
        else
            out[i] = 0;
    }
    for (int i = 0; i < o->N; ++i)
        out[i] /= C;
}

/* ---------------------------------------------------------------- online training */

double oracle_neighbourhood_weight(uint64_t cx, uint64_t cy, uint64_t bx, uint64_t by, double sigma)
{
    /* Som::calculateNeighbourhoodWeight — src/Som.cpp:949-975 (division order kept). */
    if (sigma > 1.0)
    {
        double x = (double)cx, y = (double)cy, X = (double)bx, Y = (double)by;
        return exp(-((x - X) * (x - X) / 2.0 / sigma / sigma + (y - Y) * (y - Y) / 2.0 / sigma / sigma));
    }
    else if (cx == bx && cy == by)
        return 1.0;
    return 0.0;
}

static void train_single(vsom_oracle *o, const float *v, double eta, double sigma, uint64_t *lastBMU, int decay, uint32_t *outBmu,
                         float *outDist, float *outResid2, float *delta, float *delta2)
{
    /* Som::trainSingle — src/Som.cpp:885-947. */
    const int Dm = o->Dm;
    /* :889-892  global search iff sigma > SIGMA_SWITCH_TO_LOCAL (= 1, include/SOM.hpp:37) */
    uint32_t bmu = sigma > 1 ? oracle_find_bmu_one(o, v) : oracle_find_local_bmu(o, v, *lastBMU);
    uint64_t bx = bmu % (uint32_t)o->W, by = bmu / (uint32_t)o->W;
    *lastBMU = by * (uint64_t)o->W + bx; /* :895 */

    /* :899-903  asymmetric window, exclusive end */
    double lo;
    lo = (double)bx - 2.5 * sigma;
    size_t startX = (size_t)(lo > 0. ? lo : 0.);
    lo = (double)by - 2.5 * sigma;
    size_t startY = (size_t)(lo > 0. ? lo : 0.);
    double hi;
    hi = (double)bx + 2.5 * sigma;
    size_t endX = (size_t)(hi < (double)o->W ? hi : (double)o->W);
    hi = (double)by + 2.5 * sigma;
    size_t endY = (size_t)(hi < (double)o->H ? hi : (double)o->H);

    for (size_t j = startY; j < endY; ++j)
        for (size_t i = startX; i < endX; ++i)
        {
            size_t p = j * (size_t)o->W + i;
            float *m = o->mean + p * (size_t)Dm;
            float *S = o->S + p * (size_t)Dm;
            float *sg = o->sigma + p * (size_t)Dm;
            stepper(o, v, m, delta);                                             /* :912 */
            double nw = oracle_neighbourhood_weight(i, j, bx, by, sigma);        /* :915 */
            if (decay == ORACLE_EXPONENTIAL)
            {
                o->weight[p] = o->weight[p] + (float)(nw * eta);                 /* :924 */
                float c = (float)(nw * eta);                                     /* :925 double scalar -> float */
                for (int k = 0; k < Dm; ++k)
                {
                    float t = c * delta[k];
                    m[k] = m[k] + t;
                }
            }
            else
            {
                o->weight[p] = o->weight[p] + (float)nw;                         /* :930 */
                double tw = o->weight[p] == 0 ? 1.0 : nw / (double)o->weight[p]; /* :933 */
                float c = (float)tw;
                stepper(o, v, m, delta2);                                        /* :935 (same value as delta) */
                for (int k = 0; k < Dm; ++k)
                {
                    float t = c * delta2[k];
                    m[k] = m[k] + t;
                }
            }
            double tempWeight = o->weight[p] == 0 ? 0.000001 : (double)o->weight[p]; /* :939 */
            stepper(o, v, m, delta2);                                            /* :941 Stepper on the NEW mean */
            float nwf = (float)nw;
            float twf = (float)tempWeight;
            for (int k = 0; k < Dm; ++k)
            {
                float dd = delta[k] * delta2[k];
                float t = nwf * dd;
                S[k] = S[k] + t;
                float q = S[k] / twf;                                            /* :942 */
                sg[k] = sqrtf(fabsf(q));
            }
        }
    /* :946  residual and distance on the UPDATED map */
    if (outBmu)
        *outBmu = bmu;
    int n = comparer(o, v, o->mean + (size_t)bmu * (size_t)Dm, delta);
    float s = sum_squares(o->order, delta, n);
    if (outResid2)
        *outResid2 = s; /* residual.squaredNorm() — src/Som.cpp:1167 */
    if (outDist)
        *outDist = (float)oracle_dist(o, bmu, v);
}

void oracle_train_rows(vsom_oracle *o, const float *x, size_t n, double eta, double sigma, int decay, uint64_t *lastBMU,
                       uint32_t *outBmu, float *outDist, float *outResid2)
{
    /* inner loop of Som::trainBasicSom — src/Som.cpp:1161-1171 (trainSingle + addBmu :1189-1192). */
    float *d1 = (float *)malloc(sizeof(float) * (size_t)(o->Dm > 0 ? o->Dm : 1));
    float *d2 = (float *)malloc(sizeof(float) * (size_t)(o->Dm > 0 ? o->Dm : 1));
    for (size_t r = 0; r < n; ++r)
    {
        uint64_t last = lastBMU ? lastBMU[r] : 0;
        uint32_t b;
        train_single(o, x + r * (size_t)o->Din, eta, sigma, &last, decay, &b, outDist ? outDist + r : NULL,
                     outResid2 ? outResid2 + r : NULL, d1, d2);
        o->hits[b] += 1;
        if (outBmu)
            outBmu[r] = b;
        if (lastBMU)
            lastBMU[r] = last;
    }
    free(d1);
    free(d2);
}

void oracle_train(vsom_oracle *o, const float *x, size_t n, size_t chunkRows, size_t epochs, double eta0, double etaDecay,
                  double sigma0, double sigmaDecay, int decay, int umatrixAfterEpoch, float *outMse)
{
    /* Som::trainBasicSom — src/Som.cpp:1135-1187 over the DataSet chunk protocol (src/DataSet.cpp:108-160):
     * every chunk load zeroes lastBMU (:136-137); sample order == loader order (shuffle result unused). */
    if (chunkRows == 0 || chunkRows > n)
        chunkRows = n;
    float *resid2 = (float *)malloc(sizeof(float) * (chunkRows ? chunkRows : 1));
    uint64_t *last = (uint64_t *)malloc(sizeof(uint64_t) * (chunkRows ? chunkRows : 1));
    for (size_t e = 0; e < epochs; ++e)
    {
        double eta = eta0 * exp(-etaDecay * (double)e);     /* :1145 */
        double sigma = sigma0 * exp(-sigmaDecay * (double)e); /* :1146 */
        if (sigma < 1.0)
            sigma = 1.0;                                     /* :1148-1149 */
        float mse = 0.0f;
        size_t chunks = 0;
        for (size_t begin = 0; begin < n; begin += chunkRows)
        {
            size_t rows = n - begin < chunkRows ? n - begin : chunkRows;
            memset(last, 0, sizeof(uint64_t) * rows);
            oracle_train_rows(o, x + begin * (size_t)o->Din, rows, eta, sigma, decay, last, NULL, NULL, resid2);
            for (size_t j = 0; j < rows; ++j)
                mse += resid2[j] / (float)rows;              /* :1167 */
            ++chunks;
        }
        mse /= (float)chunks;                                /* :1175 */
        if (outMse)
            outMse[e] = mse;
        if (umatrixAfterEpoch)
            oracle_update_umatrix(o, NULL);                  /* :1183-1184 */
    }
    free(resid2);
    free(last);
}

/* ---------------------------------------------------------------- batch-map trainer */

static void quirky_xy(const vsom_oracle *o, uint64_t index, uint64_t *x, uint64_t *y)
{
    /* SomIndex(const Som&, index) — src/SomIndex.cpp:15-18: y divides by the map HEIGHT (only right for square maps). */
    *x = index % (uint64_t)o->W;
    *y = (index - index % (uint64_t)o->W) / (uint64_t)o->H;
}

float oracle_batch_epoch(vsom_oracle *o, const float *x, size_t n, double sigma, int isFirst, uint64_t *lastBMU)
{
    /* Som::trainBatchSomEpoch — src/Som.cpp:756-879, with the serial execution order the reference has when its
     * parallel algorithms run on the sequential backend (no TBB): rows in order, then neurons in order. */
    const int Dm = o->Dm;
    float mse = 0.0f;
    float *r = (float *)malloc(sizeof(float) * (size_t)(Dm > 0 ? Dm : 1));
    /* phase A (:763-806): BMU per row (global on the first epoch, local walk from lastBMU afterwards), hit count,
     * mean squared residual accumulated in f32 in row order */
    for (size_t j = 0; j < n; ++j)
    {
        const float *v = x + j * (size_t)o->Din;
        uint32_t idx = isFirst ? oracle_find_bmu_one(o, v) : oracle_find_local_bmu(o, v, lastBMU[j]);
        lastBMU[j] = idx;
        o->hits[idx] += 1;
        int len = comparer(o, v, o->mean + (size_t)idx * (size_t)Dm, r);
        float s = sum_squares(o->order, r, len);
        mse = mse + s / (float)n;
    }
    /* phase B (:809-877): every neuron re-estimates its model as the incrementally weighted mean of ALL rows
     * (West / Finch), weights = neighbourhood of the row's BMU, and its sigma from the weighted squared steps */
    float *cur = (float *)malloc(sizeof(float) * (size_t)(Dm > 0 ? Dm : 1));
    float *S = (float *)malloc(sizeof(float) * (size_t)(Dm > 0 ? Dm : 1));
    float *delta = (float *)malloc(sizeof(float) * (size_t)(Dm > 0 ? Dm : 1));
    float *newMean = (float *)malloc(sizeof(float) * (size_t)o->N * (size_t)(Dm > 0 ? Dm : 1));
    for (int p = 0; p < o->N; ++p)
    {
        uint64_t cx, cy;
        quirky_xy(o, (uint64_t)p, &cx, &cy);
        float sumW = 0.f;
        for (int k = 0; k < Dm; ++k)
        {
            cur[k] = 0.0f;
            S[k] = 0.0f;
        }
        for (size_t j = 0; j < n; ++j)
        {
            uint64_t bx, by;
            quirky_xy(o, lastBMU[j], &bx, &by);
            float w = (float)oracle_neighbourhood_weight(cx, cy, bx, by, sigma);
            sumW = sumW + w;                                        /* Eq. 47 */
            stepper(o, x + j * (size_t)o->Din, cur, delta);         /* Stepper(x, currentModel) == Stepper(x, lastModel) */
            float c = w / sumW;
            for (int k = 0; k < Dm; ++k)
            {
                float d = delta[k];
                float t = c * d;
                cur[k] = cur[k] + t;                                /* Eq. 53 */
                float wd = w * d;
                float wdd = wd * d;
                S[k] = S[k] + wdd;                                  /* Eq. 68: w * Stepper(x, last) * delta, left to right */
            }
        }
        for (int k = 0; k < Dm; ++k)
        {
            newMean[(size_t)p * (size_t)Dm + (size_t)k] = cur[k];
            float q = S[k] / sumW;
            o->sigma[(size_t)p * (size_t)Dm + (size_t)k] = sqrtf(q); /* Eq. 69, no abs */
        }
        o->weight[p] = sumW;
    }
    /* neurons are independent of each other's new value (they only read the rows), so writing them at the end equals
     * the reference's in-place assignment */
    memcpy(o->mean, newMean, sizeof(float) * (size_t)o->N * (size_t)Dm);
    free(newMean);
    free(delta);
    free(S);
    free(cur);
    free(r);
    return mse;
}

void oracle_train_batch(vsom_oracle *o, const float *x, size_t n, size_t chunkRows, size_t epochs, double sigma0, double sigmaDecay,
                        int umatrixAfterEpoch, float *outMse)
{
    /* Som::trainBatchSom — src/Som.cpp:716-754: returns (not continues) as soon as sigma < 1; lastBMU is zeroed by
     * every chunk load (src/DataSet.cpp:136-137). */
    if (chunkRows == 0 || chunkRows > n)
        chunkRows = n;
    uint64_t *last = (uint64_t *)malloc(sizeof(uint64_t) * (chunkRows ? chunkRows : 1));
    for (size_t e = 0; e < epochs; ++e)
    {
        double sigma = sigma0 * exp(-sigmaDecay * (double)e);
        if (sigma < 1.0)
            break;
        float mse = 0.0f;
        size_t chunks = 0;
        for (size_t begin = 0; begin < n; begin += chunkRows)
        {
            size_t rows = n - begin < chunkRows ? n - begin : chunkRows;
            memset(last, 0, sizeof(uint64_t) * rows);
            mse += oracle_batch_epoch(o, x + begin * (size_t)o->Din, rows, sigma, e == 0, last);
            ++chunks;
        }
        mse /= (float)chunks;
        if (outMse)
            outMse[e] = mse;
        if (umatrixAfterEpoch)
            oracle_update_umatrix(o, NULL);
    }
    free(last);
}

/* ---------------------------------------------------------------- scoring */

double oracle_evaluate(const vsom_oracle *o, const float *x, size_t n)
{
    /* Som::evaluate — src/Som.cpp:490-523 for all-continuous columns: the binary cross-entropy vector is
     * multiplied by binary(=0) so its norm is exactly 0; what remains is the f64 running mean of the
     * BMU distance, evaluated as  error += 1./(i+1.) * (dist + 0 - error)  (:519). */
    double error = 0;
    for (size_t i = 0; i < n; ++i)
    {
        const float *v = x + i * (size_t)o->Din;
        uint32_t b = oracle_find_bmu_one(o, v);
        error += 1. / ((double)i + 1.0) * (oracle_dist(o, b, v) + sqrt((double)0.0f) - error);
    }
    return error;
}

int oracle_measure_similarity(const vsom_oracle *o, const float *x, size_t n, int numSigmas, uint64_t minHits)
{
    /* Som::measureSimilarity — src/Som.cpp:631-714 (all columns valid): rows 0..n-1, then the row with the
     * largest per-dimension deviation is visited again and only that visit can clear `success`.
     * Quirks kept: reversed sigma clamp (:658), signed delta compared to a stored |delta| (:686-688). */
    if (n == 0)
        return 1;
    int success = 1;
    float maxValue = -99999999.f;
    size_t maxRow = 0;
    int last = 0;
    const int D = o->Din;
    const float k = (float)numSigmas;
    for (size_t i = 0; i < n + 1; i++)
    {
        if (i == n)
        {
            i = maxRow;
            last = 1;
        }
        const float *v = x + i * (size_t)D;
        uint32_t pos = oracle_find_restricted_bmu_one(o, v, minHits);
        const float *m = o->mean + (size_t)pos * (size_t)o->Dm;
        const float *sg = o->sigma + (size_t)pos * (size_t)o->Dm;
        for (int d = 0; d < D; ++d)
        {
            float sM = sg[d] > 0.00001f ? 0.00001f : sg[d];
            float delta = ((v[d] - m[d]) / sM) / k;
            float lo = m[d] - sM * k;
            float hi = m[d] + sM * k;
            if (delta > maxValue)
            {
                maxValue = (float)fabs((double)delta);
                maxRow = i;
            }
            if (last)
                if (v[d] < lo || v[d] > hi)
                    success = 0;
        }
        if (last)
            break;
    }
    return success;
}

/* ---------------------------------------------------------------- U-matrix */

void oracle_update_umatrix(vsom_oracle *o, double *out)
{
    /* Som::updateUMatrix — src/Som.cpp:999-1111: 9-way case split; straight neighbours weight 1, diagonal
     * neighbours 0.3; divided by 8 / 5 / 3; the sum is formed left to right in double in the textual order
     * of the source.  Raw(p, u) always divides by p's OWN sigma. */
    const int W = o->W, H = o->H, Dm = o->Dm;
    const double df = 0.3;
#define RAW(di, dj) oracle_dist_raw(o, (size_t)(i * W + j), o->mean + (size_t)((i + (di)) * W + (j + (dj))) * (size_t)Dm)
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
        {
            double U;
            if (j > 0 && i > 0 && j < W - 1 && i < H - 1)
                U = (RAW(0, -1) + RAW(0, +1) + RAW(+1, 0) + RAW(-1, 0) + RAW(-1, -1) * df + RAW(+1, -1) * df + RAW(-1, +1) * df +
                     RAW(+1, +1) * df) / 8;
            else if (i == 0 && j > 0 && j < W - 1)
                U = (RAW(0, -1) + RAW(0, +1) + RAW(+1, 0) + RAW(+1, -1) * df + RAW(+1, +1) * df) / 5;
            else if (i == H - 1 && j > 0 && j < W - 1)
                U = (RAW(0, -1) + RAW(0, +1) + RAW(-1, 0) + RAW(-1, -1) * df + RAW(-1, +1) * df) / 5;
            else if (j == 0 && i > 0 && i < H - 1)
                U = (RAW(0, +1) + RAW(+1, 0) + RAW(-1, 0) + RAW(-1, +1) * df + RAW(+1, +1) * df) / 5;
            else if (j == W - 1 && i > 0 && i < H - 1)
                U = (RAW(0, -1) + RAW(+1, 0) + RAW(-1, 0) + RAW(-1, -1) * df + RAW(+1, -1) * df) / 5;
            else if (j == 0 && i == 0)
                U = (RAW(0, +1) + RAW(+1, 0) + RAW(+1, +1) * df) / 3;
            else if (j == W - 1 && i == 0)
                U = (RAW(0, -1) + RAW(+1, 0) + RAW(+1, -1) * df) / 3;
            else if (j == 0 && i == H - 1)
                U = (RAW(0, +1) + RAW(-1, 0) + RAW(-1, +1) * df) / 3;
            else if (j == W - 1 && i == H - 1)
                U = (RAW(0, -1) + RAW(-1, 0) + RAW(-1, -1) * df) / 3;
            else
                U = 0;
            o->umatrix[i * W + j] = U;
        }
#undef RAW
    if (out)
        memcpy(out, o->umatrix, sizeof(double) * (size_t)o->N);
}

/* ---------------------------------------------------------------- SomIndex build */

void oracle_build_index(const uint32_t *bmu, size_t n, int N, uint64_t *counts, uint64_t *offsets, uint32_t *rowIds)
{
    /* The "SomIndex build" of the north-star: what Som::addBmu (src/Som.cpp:1189-1192) accumulates
     * (histogram of BMU ids) plus the rows grouped by BMU in ascending row order — the grouping the
     * batch-map trainer scans per neuron (src/Som.cpp:845-868).  offsets has N+1 entries. */
    for (int p = 0; p < N; ++p)
        counts[p] = 0;
    for (size_t r = 0; r < n; ++r)
        counts[bmu[r]] += 1;
    offsets[0] = 0;
    for (int p = 0; p < N; ++p)
        offsets[p + 1] = offsets[p] + counts[p];
    uint64_t *cur = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    memcpy(cur, offsets, sizeof(uint64_t) * (size_t)N);
    for (size_t r = 0; r < n; ++r)
        rowIds[cur[bmu[r]]++] = (uint32_t)r;
    free(cur);
}

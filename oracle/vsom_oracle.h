/* TEST INFRASTRUCTURE — CPU restatement of the reference VSOM hot path (plain C).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the CHECKER.  The product path (libvsom_b200.so) never links, loads or
 * calls anything declared here and has no CPU fallback.
 *
 * Parity status: PINNED.  oracle/vsom_oracle.c is checked bit-for-bit against (a) the known-answer
 * vectors of SURVEY.md Appendix F (tests/golden/appendix_f.json) and (b) the reference's own
 * translation units compiled unmodified into oracle/_ref/libvsom_ref.so (tests/test_oracle_pin.py,
 * tests/golden/*.npz produced by oracle/gen_golden.py).  The reference's own test-suite holds no
 * numeric assertion for this path, so those two are the pins.
 *
 * ORACLE_ORDER_EIGEN_SSE (the summation order of dot() under real Eigen + SSE2) restates a THIRD-PARTY algorithm: Eigen is
 * a system package of the reference (README.md:64-69, version unpinned; 3.3.x and 3.4.0 share the code in question), absent
 * from /root/reference and from this image.  Its pin is therefore indirect: the same order is restated a second time,
 * independently, inside include/compat/Eigen/Dense (-DVSOM_COMPAT_EIGEN_SSE_REDUX), the reference's own translation units
 * are compiled against that (oracle/_ref/libvsom_ref_sse.so), and the two are checked against each other bit for bit
 * (tests/test_oracle_pin.py).  "Parity unpinned against a genuine Eigen build" for that mode — no Eigen exists here to
 * run.
 *
 * All citations are file:line in /root/reference.
 */
#ifndef VSOM_ORACLE_H
#define VSOM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_STANDARD = 0, ORACLE_MEDIAN = 1, ORACLE_CLR = 2 };
enum { ORACLE_EXPONENTIAL = 0, ORACLE_INVERSE = 1 };
/* summation order of dot() / squaredNorm(): see sum_terms() in vsom_oracle.c.  Values match vsom_reduction_order. */
enum { ORACLE_ORDER_SEQUENTIAL = 0, ORACLE_ORDER_EIGEN_SSE = 2 };

typedef struct vsom_oracle
{
    int W, H, N;   /* grid; node p = y*W + x (src/Som.cpp:119,191,895) */
    int Din;       /* sample length */
    int Dm;        /* model-vector length = Transformation::Length(Din) (src/Transformation.cpp:31-35,162-165) */
    int P;         /* CLR pair count Din*(Din-1)/2, else 0 */
    int transform; /* ORACLE_* */
    float *mean, *S, *sigma; /* N x Dm row-major: Som::map, SMap, sigmaMap (include/SOM.hpp:56-58) */
    float *weight;           /* N: weightMap */
    uint64_t *hits;          /* N: bmuHits */
    double *umatrix;         /* N */
    int order;               /* ORACLE_ORDER_* (default sequential) */
} vsom_oracle;

vsom_oracle *oracle_create(int W, int H, int Din, int transform);
void oracle_destroy(vsom_oracle *o);
int oracle_depth(const vsom_oracle *o);
void oracle_set_order(vsom_oracle *o, int order);
int oracle_get_order(const vsom_oracle *o);
void oracle_random_initialize(vsom_oracle *o, int seed, float sigma);
void oracle_get_state(const vsom_oracle *o, float *mean, float *S, float *sigma, float *weight, uint64_t *hits);
void oracle_set_state(vsom_oracle *o, const float *mean, const float *S, const float *sigma, const float *weight, const uint64_t *hits);

double oracle_neighbourhood_weight(uint64_t cx, uint64_t cy, uint64_t bx, uint64_t by, double sigma);
double oracle_dist(const vsom_oracle *o, size_t pos, const float *v);     /* f32 sequential accumulate */
double oracle_dist_f64(const vsom_oracle *o, size_t pos, const float *v); /* same residuals, f64 accumulate (near-tie classifier) */
void oracle_all_dists(const vsom_oracle *o, const float *v, double *out);
double oracle_dist_raw(const vsom_oracle *o, size_t pos, const float *u);
uint32_t oracle_find_bmu_one(const vsom_oracle *o, const float *v);
uint32_t oracle_find_local_bmu(const vsom_oracle *o, const float *v, uint64_t start);
uint32_t oracle_find_restricted_bmu_one(const vsom_oracle *o, const float *v, uint64_t minHits);
void oracle_find_bmu(const vsom_oracle *o, const float *x, size_t n, uint32_t *outBmu, float *outDist);
void oracle_find_restricted_bmu(const vsom_oracle *o, const float *x, size_t n, uint64_t minHits, uint32_t *outBmu);
void oracle_find_restricted_bmd(const vsom_oracle *o, const float *v, uint64_t minHits, double *out);

void oracle_train_rows(vsom_oracle *o, const float *x, size_t n, double eta, double sigma, int decay, uint64_t *lastBMU,
                       uint32_t *outBmu, float *outDist, float *outResid2);
void oracle_train(vsom_oracle *o, const float *x, size_t n, size_t chunkRows, size_t epochs, double eta0, double etaDecay,
                  double sigma0, double sigmaDecay, int decay, int umatrixAfterEpoch, float *outMse);
/* Batch-map trainer (a "next" row): one chunk-epoch, and the whole trainBatchSom schedule. */
float oracle_batch_epoch(vsom_oracle *o, const float *x, size_t n, double sigma, int isFirst, uint64_t *lastBMU);
void oracle_train_batch(vsom_oracle *o, const float *x, size_t n, size_t chunkRows, size_t epochs, double sigma0, double sigmaDecay,
                        int umatrixAfterEpoch, float *outMse);
double oracle_evaluate(const vsom_oracle *o, const float *x, size_t n);
int oracle_measure_similarity(const vsom_oracle *o, const float *x, size_t n, int numSigmas, uint64_t minHits);
void oracle_update_umatrix(vsom_oracle *o, double *out);
/* SomIndex build: histogram of BMU ids and rows grouped by BMU (stable): offsets[N+1], rowIds[n]. */
void oracle_build_index(const uint32_t *bmu, size_t n, int N, uint64_t *counts, uint64_t *offsets, uint32_t *rowIds);

#ifdef __cplusplus
}
#endif
#endif

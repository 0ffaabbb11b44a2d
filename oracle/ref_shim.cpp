// TEST INFRASTRUCTURE — not part of the product path.
//
// extern "C" shim around the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libvsom_ref.so).  It lets the Python tests, the golden-vector
// generator and bench.py's CPU-baseline legs drive the reference's own Som / DataSet / Transformation
// through their public API, one call at a time, on caller-supplied row-major float buffers.
// Nothing here re-implements the algorithm: every function forwards to a reference method
// (cited per function, paths relative to /root/reference).
#include "SOM.hpp"
#include "DataSet.hpp"
#include "IDataLoader.hpp"
#include "Transformation.hpp"

#include <cstdint>
#include <cstring>
#include <iostream>
#include <optional>
#include <streambuf>
#include <string>
#include <vector>

namespace
{
// The reference prints progress from inside its hot loops (src/Som.cpp:1151,1169-1170,
// src/DataSet.cpp:159); silence std::cout while a shim call runs.
struct NullBuf : std::streambuf
{
    int overflow(int c) override { return c; }
};
struct Quiet
{
    NullBuf nb;
    std::streambuf *old;
    Quiet() : old(std::cout.rdbuf(&nb)) {}
    ~Quiet() { std::cout.rdbuf(old); }
};

// In-memory row source implementing the reference's loader interface (include/IDataLoader.hpp:18-48).
// Serves `rows` in loader order, `chunk` rows per load(); wraps to the start after the last chunk.
class MemLoader : public IDataLoader
{
  public:
    MemLoader(const float *x, size_t n, size_t depth, size_t chunk, const int *validMask = nullptr)
        : _x{x}, _valid{validMask}, _n{n}, _depth{depth}, _chunk{chunk ? chunk : n}, _weights(depth, 1.0f), _binary(depth, 0),
          _continuous(depth, 1), _names(depth)
    {
        for (size_t i = 0; i < depth; ++i)
            _names[i] = "c" + std::to_string(i);
    }
    size_t load() override
    {
        const size_t begin = m_currentIndex;
        const size_t end = std::min(begin + _chunk, _n);
        data.clear();
        data.reserve(end - begin);
        for (size_t r = begin; r < end; ++r)
        {
            RowData row{Eigen::VectorXf(_depth), std::vector<int>(_depth, 1)};
            std::memcpy(row.values.data(), _x + r * _depth, _depth * sizeof(float));
            if (_valid)
                for (size_t d = 0; d < _depth; ++d)
                    row.valid[d] = _valid[r * _depth + d];
            data.push_back(std::move(row));
        }
        m_currentIndex = (end >= _n) ? 0 : end;
        return data.size();
    }
    std::vector<RowData> getPreview(size_t) override { return {}; }
    bool open(const char *) override { return true; }
    std::vector<std::string> findAllColumns() override { return _names; }
    void setColumnSpec(const std::vector<ColumnSpec>) noexcept override {}
    const std::vector<ColumnSpec> getColumnSpec() noexcept override { return {}; }
    float getWeight(size_t i) override { return _weights[i]; }
    const std::vector<float> getWeights() const noexcept override { return _weights; }
    const std::vector<int> &getBinary() const noexcept override { return _binary; }
    const std::vector<int> &getContinuous() const noexcept override { return _continuous; }
    std::string getName(size_t i) const noexcept override { return _names[i]; }
    const std::vector<std::string> getNames() const noexcept override { return _names; }
    size_t getDepth() const noexcept override { return _depth; }
    bool isAtStartOfDataStream() const noexcept override { return m_currentIndex == 0; }

  private:
    const float *_x;
    const int *_valid;
    size_t _n, _depth, _chunk;
    std::vector<float> _weights;
    std::vector<int> _binary, _continuous;
    std::vector<std::string> _names;
};

Transformation makeTransform(int kind)
{
    std::vector<std::string> names;
    switch (kind)
    {
    case 1:
        return Transformation::StandardMedianEstimator(names); // src/Transformation.cpp:41-77
    case 2:
        return Transformation::CombinatorialLinearRegression(names); // src/Transformation.cpp:79-167
    default:
        return Transformation::Standard(names); // src/Transformation.cpp:3-39
    }
}

// Exposes the protected state of the reference Som (include/SOM.hpp:55-64).
struct RefSom : Som
{
    using Som::Som;
    int dIn{0};
    void getState(float *mean, float *S, float *sigma, float *weight, uint64_t *hits) const
    {
        const size_t N = width * height;
        for (size_t p = 0; p < N; ++p)
        {
            if (mean)
                std::memcpy(mean + p * depth, map[p].data(), depth * sizeof(float));
            if (S)
                std::memcpy(S + p * depth, SMap[p].data(), depth * sizeof(float));
            if (sigma)
                std::memcpy(sigma + p * depth, sigmaMap[p].data(), depth * sizeof(float));
            if (weight)
                weight[p] = weightMap[p];
            if (hits)
                hits[p] = bmuHits[p];
        }
    }
    void setState(const float *mean, const float *S, const float *sigma, const float *weight, const uint64_t *hits)
    {
        const size_t N = width * height;
        for (size_t p = 0; p < N; ++p)
        {
            if (mean)
                std::memcpy(map[p].data(), mean + p * depth, depth * sizeof(float));
            if (S)
                std::memcpy(SMap[p].data(), S + p * depth, depth * sizeof(float));
            if (sigma)
                std::memcpy(sigmaMap[p].data(), sigma + p * depth, depth * sizeof(float));
            if (weight)
                weightMap[p] = weight[p];
            if (hits)
                bmuHits[p] = hits[p];
        }
    }
    // One reference online step + the caller-side bookkeeping of trainBasicSom (src/Som.cpp:1163-1167).
    void step(const Eigen::VectorXf &v, const Eigen::VectorXf &valid, const Eigen::VectorXf &weights, double eta, double sigma,
              size_t &lastBMU, int decay, uint32_t *bmu, float *dist, float *resid2)
    {
        auto r = trainSingle(v, valid, weights, eta, sigma, lastBMU, static_cast<WeigthDecayFunction>(decay)); // src/Som.cpp:885-947
        addBmu(r.bmu);                                                                                          // src/Som.cpp:1189-1192
        if (bmu)
            *bmu = static_cast<uint32_t>(getIndex(r.bmu));
        if (dist)
            *dist = r.distanceError;
        if (resid2)
            *resid2 = r.residual.squaredNorm();
    }
    std::vector<float> mse() const { return metrics.MeanSquaredError; }
};

Eigen::VectorXf rowOf(const float *x, size_t r, int d)
{
    Eigen::VectorXf v(d);
    std::memcpy(v.data(), x + r * static_cast<size_t>(d), static_cast<size_t>(d) * sizeof(float));
    return v;
}
} // namespace

extern "C"
{
    void *ref_create(int W, int H, int dIn, int transformKind)
    {
        Quiet q;
        auto t = makeTransform(transformKind);
        // Callers size the model vector with Transformation::Length (tests/performance/perf_tests.cpp:338-339).
        auto *s = new RefSom(static_cast<size_t>(W), static_cast<size_t>(H), t.Length(static_cast<size_t>(dIn)), t);
        s->dIn = dIn;
        return s;
    }
    void ref_destroy(void *h) { delete static_cast<RefSom *>(h); }
    int ref_depth(void *h) { return static_cast<int>(static_cast<RefSom *>(h)->getDepth()); }
    void ref_random_initialize(void *h, int seed, float sigma) { static_cast<RefSom *>(h)->randomInitialize(seed, sigma); } // src/Som.cpp:977-997
    void ref_get_state(void *h, float *mean, float *S, float *sigma, float *weight, uint64_t *hits)
    {
        static_cast<RefSom *>(h)->getState(mean, S, sigma, weight, hits);
    }
    void ref_set_state(void *h, const float *mean, const float *S, const float *sigma, const float *weight, const uint64_t *hits)
    {
        static_cast<RefSom *>(h)->setState(mean, S, sigma, weight, hits);
    }

    // n consecutive trainSingle + addBmu calls at fixed (eta, sigma): the inner loop of trainBasicSom
    // (src/Som.cpp:1161-1171).  lastBMU[r] is read and written per row like DataSet::getLastBMU(j).
    void ref_train_rows(void *h, const float *x, size_t n, double eta, double sigma, int decay, uint64_t *lastBMU, uint32_t *outBmu,
                        float *outDist, float *outResid2)
    {
        Quiet q;
        auto *s = static_cast<RefSom *>(h);
        const int d = s->dIn;
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(d);
        for (size_t r = 0; r < n; ++r)
        {
            size_t last = lastBMU ? static_cast<size_t>(lastBMU[r]) : 0;
            s->step(rowOf(x, r, d), ones, ones, eta, sigma, last, decay, outBmu ? outBmu + r : nullptr, outDist ? outDist + r : nullptr,
                    outResid2 ? outResid2 + r : nullptr);
            if (lastBMU)
                lastBMU[r] = last;
        }
    }

    // Som::train over a chunk-streaming DataSet (src/Som.cpp:1113-1187, src/DataSet.cpp:108-160).
    void ref_train(void *h, const float *x, size_t n, size_t chunkRows, size_t epochs, double eta0, double etaDecay, double sigma0,
                   double sigmaDecay, int decay, int umatrixAfterEpoch, float *outMse)
    {
        Quiet q;
        auto *s = static_cast<RefSom *>(h);
        MemLoader loader(x, n, static_cast<size_t>(s->dIn), chunkRows);
        DataSet ds(loader);
        s->train(ds, epochs, eta0, etaDecay, sigma0, sigmaDecay, static_cast<Som::WeigthDecayFunction>(decay), umatrixAfterEpoch != 0);
        if (outMse)
        {
            auto m = s->mse();
            for (size_t i = 0; i < epochs && i < m.size(); ++i)
                outMse[i] = m[i];
        }
    }

    // findBmu + euclidianWeightedDist per row (src/Som.cpp:291-309, 124-141).
    void ref_find_bmu(void *h, const float *x, size_t n, uint32_t *outBmu, float *outDist)
    {
        auto *s = static_cast<RefSom *>(h);
        const int d = s->dIn;
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(d);
        for (size_t r = 0; r < n; ++r)
        {
            auto v = rowOf(x, r, d);
            SomIndex b = s->findBmu(v, ones, ones);
            if (outBmu)
                outBmu[r] = static_cast<uint32_t>(s->getIndex(b));
            if (outDist)
                outDist[r] = static_cast<float>(s->euclidianWeightedDist(b, v, ones, ones));
        }
    }
    // Distance of one row to every node (src/Som.cpp:124-141), as the double the reference returns.
    void ref_all_dists(void *h, const float *v, double *out)
    {
        auto *s = static_cast<RefSom *>(h);
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(s->dIn);
        auto vv = rowOf(v, 0, s->dIn);
        const size_t N = s->getWidth() * s->getHeight();
        for (size_t p = 0; p < N; ++p)
            out[p] = s->euclidianWeightedDist(p, vv, ones, ones);
    }
    // findLocalBmu (src/Som.cpp:335-454) from a given start node.
    uint32_t ref_find_local_bmu(void *h, const float *v, uint64_t start)
    {
        auto *s = static_cast<RefSom *>(h);
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(s->dIn);
        size_t st = static_cast<size_t>(start);
        return static_cast<uint32_t>(s->getIndex(s->findLocalBmu(rowOf(v, 0, s->dIn), ones, st, ones)));
    }
    // findRestrictedBmu (src/Som.cpp:313-332).
    void ref_find_restricted_bmu(void *h, const float *x, size_t n, uint64_t minHits, uint32_t *outBmu)
    {
        auto *s = static_cast<RefSom *>(h);
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(s->dIn);
        for (size_t r = 0; r < n; ++r)
            outBmu[r] = static_cast<uint32_t>(s->getIndex(s->findRestrictedBmu(rowOf(x, r, s->dIn), ones, static_cast<size_t>(minHits), ones)));
    }
    // findRestrictedBmd (src/Som.cpp:457-487) for one row.
    void ref_find_restricted_bmd(void *h, const float *v, uint64_t minHits, double *out)
    {
        auto *s = static_cast<RefSom *>(h);
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(s->dIn);
        auto d = s->findRestrictedBmd(rowOf(v, 0, s->dIn), ones, static_cast<size_t>(minHits), ones);
        std::memcpy(out, d.data(), d.size() * sizeof(double));
    }
    // evaluate (src/Som.cpp:490-523) on one loaded chunk holding all n rows.
    double ref_evaluate(void *h, const float *x, size_t n)
    {
        Quiet q;
        auto *s = static_cast<RefSom *>(h);
        MemLoader loader(x, n, static_cast<size_t>(s->dIn), n);
        DataSet ds(loader);
        ds.loadNextDataFromStream(); // tests/performance/perf_tests.cpp:152-153
        return s->evaluate(ds);
    }
    // measureSimilarity (src/Som.cpp:631-714).
    int ref_measure_similarity(void *h, const float *x, size_t n, int numSigmas, uint64_t minHits)
    {
        Quiet q;
        auto *s = static_cast<RefSom *>(h);
        MemLoader loader(x, n, static_cast<size_t>(s->dIn), n);
        DataSet ds(loader);
        ds.loadNextDataFromStream();
        return s->measureSimilarity(&ds, numSigmas, static_cast<size_t>(minHits));
    }
    // updateUMatrix + getUMatrix (src/Som.cpp:999-1111, 159-162).
    void ref_update_umatrix(void *h, double *out)
    {
        auto *s = static_cast<RefSom *>(h);
        s->updateUMatrix(Eigen::VectorXf::Ones(static_cast<Eigen::Index>(s->getDepth())));
        auto u = s->getUMatrix();
        std::memcpy(out, u.getData().data(), u.getData().size() * sizeof(double));
    }
    // euclidianWeightedDistRaw (src/Som.cpp:143-157); v has the MODEL length (a neighbour's mean).
    double ref_dist_raw(void *h, uint64_t pos, const float *v)
    {
        auto *s = static_cast<RefSom *>(h);
        const int dm = static_cast<int>(s->getDepth());
        const Eigen::VectorXf ones = Eigen::VectorXf::Ones(dm);
        size_t p = static_cast<size_t>(pos);
        return s->euclidianWeightedDistRaw(p, rowOf(v, 0, dm), ones, ones);
    }
    // calculateNeighbourhoodWeight (src/Som.cpp:949-975).
    double ref_neighbourhood_weight(uint64_t cx, uint64_t cy, uint64_t bx, uint64_t by, double sigma)
    {
        size_t a = cx, b = cy, c = bx, d = by;
        return Som::calculateNeighbourhoodWeight(a, b, c, d, sigma);
    }
    // Batch-map trainer (src/Som.cpp:716-879) — a "next" row; exposed so a later round can pin it.
    void ref_train_batch(void *h, const float *x, size_t n, size_t chunkRows, size_t epochs, double sigma0, double sigmaDecay, int umatrixAfterEpoch,
                         float *outMse)
    {
        Quiet q;
        auto *s = static_cast<RefSom *>(h);
        MemLoader loader(x, n, static_cast<size_t>(s->dIn), chunkRows);
        DataSet ds(loader);
        s->train(ds, epochs, 0.0, 0.0, sigma0, sigmaDecay, Som::WeigthDecayFunction::BatchMap, umatrixAfterEpoch != 0);
        if (outMse)
        {
            auto m = s->mse();
            for (size_t i = 0; i < epochs && i < m.size(); ++i)
                outMse[i] = m[i];
        }
    }
    const char *ref_eigen_kind()
    {
#if defined(VSOM_COMPAT_EIGEN_DENSE) && defined(VSOM_COMPAT_EIGEN_SSE_REDUX)
        return "compat-standin(Eigen SSE2 packet redux order)";
#elif defined(VSOM_COMPAT_EIGEN_DENSE)
        return "compat-standin(sequential dot)";
#else
        return "real-eigen";
#endif
    }
}

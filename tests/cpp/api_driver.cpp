// api_driver.cpp — one program written against the reference's public C++ API (SOM.hpp, DataSet.hpp,
// Transformation.hpp, IDataLoader.hpp).  It is compiled twice from this same source:
//   * against /root/reference's headers + its own translation units  -> oracle/_ref/api_driver_ref  (CPU reference)
//   * against this repository's include/ + libvsom_host.so            -> tests/cpp/api_driver_b200   (B200)
// and tests/test_gpu_host_api.py requires the two output files to be identical byte for byte.
//
// usage: api_driver <case.bin> <out.bin>
// case.bin: int32 W,H,Din,transform,decay,epochs,chunk,seed,n,trainRows ; float64 eta0,etaDecay,sigma0,sigmaDecay ;
//           float32 x[n*Din]
// trainRows > 0 selects the LIGHT layout for large maps: training sees only the first trainRows rows, and the output holds
// the per-epoch MSE, hit counts, U-matrix, evaluate(), measureSimilarity() and the restricted BMU of every one of the n rows
// (no plane dumps, no checkpoint) — the scoring calls then run over thousands of rows, i.e. on the tensor-core path of the
// B200 build.
#include "DataSet.hpp"
#include "IDataLoader.hpp"
#include "SOM.hpp"
#include "Transformation.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <streambuf>
#include <string>
#include <vector>

namespace
{
class RowsLoader : public IDataLoader
{
  public:
    RowsLoader(const float *x, size_t n, size_t depth, size_t chunk) : _x{x}, _n{n}, _depth{depth}, _chunk{chunk ? chunk : n}, _w(depth, 1.0f), _bin(depth, 0), _cont(depth, 1), _names(depth, "c") {}
    size_t load() override
    {
        const size_t a = m_currentIndex, b = std::min(a + _chunk, _n);
        data.clear();
        for (size_t r = a; r < b; ++r)
        {
            RowData row{Eigen::VectorXf(_depth), std::vector<int>(_depth, 1)};
            std::memcpy(row.values.data(), _x + r * _depth, _depth * sizeof(float));
            data.push_back(row);
        }
        m_currentIndex = b >= _n ? 0 : b;
        return data.size();
    }
    std::vector<RowData> getPreview(size_t) override { return {}; }
    bool open(const char *) override { return true; }
    std::vector<std::string> findAllColumns() override { return _names; }
    void setColumnSpec(const std::vector<ColumnSpec>) noexcept override {}
    const std::vector<ColumnSpec> getColumnSpec() noexcept override { return {}; }
    float getWeight(size_t i) override { return _w[i]; }
    const std::vector<float> getWeights() const noexcept override { return _w; }
    const std::vector<int> &getBinary() const noexcept override { return _bin; }
    const std::vector<int> &getContinuous() const noexcept override { return _cont; }
    std::string getName(size_t i) const noexcept override { return _names[i]; }
    const std::vector<std::string> getNames() const noexcept override { return _names; }
    size_t getDepth() const noexcept override { return _depth; }
    bool isAtStartOfDataStream() const noexcept override { return m_currentIndex == 0; }

  private:
    const float *_x;
    size_t _n, _depth, _chunk;
    std::vector<float> _w;
    std::vector<int> _bin, _cont;
    std::vector<std::string> _names;
};

struct NullBuf : std::streambuf
{
    int overflow(int c) override { return c; }
};
template <typename T> void put(FILE *f, const T *p, size_t n) { std::fwrite(p, sizeof(T), n, f); }
} // namespace

int main(int argc, char **argv)
{
    if (argc != 3)
        return 2;
    FILE *in = std::fopen(argv[1], "rb");
    if (!in)
        return 3;
    int32_t h[10];
    double sched[4];
    if (std::fread(h, sizeof(int32_t), 10, in) != 10 || std::fread(sched, sizeof(double), 4, in) != 4)
        return 4;
    const int W = h[0], H = h[1], Din = h[2], transform = h[3], decay = h[4], epochs = h[5], chunk = h[6], seed = h[7], n = h[8], trainRows = h[9];
    const bool light = trainRows > 0;
    std::vector<float> x(static_cast<size_t>(n) * Din);
    if (std::fread(x.data(), sizeof(float), x.size(), in) != x.size())
        return 5;
    std::fclose(in);

    NullBuf nb;
    std::streambuf *old = std::cout.rdbuf(&nb); // the reference prints progress from its hot loop

    std::vector<std::string> names;
    Transformation t = transform == 1 ? Transformation::StandardMedianEstimator(names)
                       : transform == 2 ? Transformation::CombinatorialLinearRegression(names)
                                        : Transformation::Standard(names);
    Som som(static_cast<size_t>(W), static_cast<size_t>(H), t.Length(static_cast<size_t>(Din)), t);
    som.randomInitialize(seed, 1.0f);
    RowsLoader loader(x.data(), static_cast<size_t>(light ? trainRows : n), static_cast<size_t>(Din), static_cast<size_t>(chunk));
    DataSet ds(loader);
    som.train(ds, static_cast<size_t>(epochs), sched[0], sched[1], sched[2], sched[3], static_cast<Som::WeigthDecayFunction>(decay), true);

    FILE *out = std::fopen(argv[2], "wb");
    if (!out)
        return 6;
    const auto metrics = som.getMetrics();
    put(out, metrics.MeanSquaredError.data(), metrics.MeanSquaredError.size());
    const size_t N = som.getWidth() * som.getHeight();
    for (size_t p = 0; p < N && !light; ++p)
    {
        const Eigen::VectorXf m = som.getNeuron(p), s = som.getSigmaNeuron(p);
        put(out, m.data(), static_cast<size_t>(m.size()));
        put(out, s.data(), static_cast<size_t>(s.size()));
    }
    const Eigen::VectorXf wm = som.getWeigthMap();
    put(out, wm.data(), static_cast<size_t>(wm.size()));
    const std::vector<size_t> hits = som.getBmuHits();
    for (size_t v : hits)
    {
        const uint64_t u = v;
        put(out, &u, 1);
    }
    const UMatrix um = som.getUMatrix(); // updated after the last epoch
    put(out, um.getData().data(), um.getData().size());
    const double umin = um.getMinValue(), umax = um.getMaxValue();
    put(out, &umin, 1);
    put(out, &umax, 1);

    // scoring on the first chunk
    RowsLoader loader2(x.data(), static_cast<size_t>(n), static_cast<size_t>(Din), static_cast<size_t>(n));
    DataSet ds2(loader2);
    ds2.loadNextDataFromStream();
    if (transform != 2) // Som::evaluate mixes model- and sample-sized vectors for CLR in the reference
    {
        const double e = som.evaluate(ds2);
        put(out, &e, 1);
        const int32_t sim = som.measureSimilarity(&ds2, 3, 0);
        put(out, &sim, 1);
    }
    const Eigen::VectorXf ones = Eigen::VectorXf::Ones(Din);
    // ---- restricted BMU of EVERY row (the per-row loop a user of the reference writes; Som::mapDataSet in one device pass)
    //      and the rows grouped by BMU (SomIndex build)
    {
        std::vector<size_t> bmuAll(static_cast<size_t>(n));
        std::vector<float> distAll(static_cast<size_t>(n));
#ifdef VSOM_B200_API
        som.mapDataSet(ds2, bmuAll, distAll, 2);
#else
        for (int r = 0; r < n; ++r)
        {
            Eigen::VectorXf v(Din);
            std::memcpy(v.data(), x.data() + static_cast<size_t>(r) * Din, sizeof(float) * Din);
            const SomIndex b = som.findRestrictedBmu(v, ones, 2, ones);
            bmuAll[static_cast<size_t>(r)] = som.getIndex(b);
            distAll[static_cast<size_t>(r)] = light ? 0.0f : static_cast<float>(som.euclidianWeightedDist(b, v, ones, ones));
        }
#endif
        for (int r = 0; r < n; ++r)
        {
            const uint64_t b = bmuAll[static_cast<size_t>(r)];
            put(out, &b, 1);
            if (!light)
                put(out, &distAll[static_cast<size_t>(r)], 1);
        }
        std::vector<size_t> counts(N, 0), offsets(N + 1, 0);
        std::vector<unsigned> rowIds(static_cast<size_t>(n));
#ifdef VSOM_B200_API
        som.buildIndex(bmuAll, counts, offsets, rowIds);
#else
        for (size_t b : bmuAll)
            counts[b] += 1;
        for (size_t p = 0; p < N; ++p)
            offsets[p + 1] = offsets[p] + counts[p];
        std::vector<size_t> fill(offsets.begin(), offsets.end() - 1);
        for (int r = 0; r < n; ++r)
            rowIds[fill[bmuAll[static_cast<size_t>(r)]]++] = static_cast<unsigned>(r);
#endif
        for (size_t p = 0; p < N; ++p)
        {
            const uint64_t c = counts[p], o = offsets[p + 1];
            put(out, &c, 1);
            put(out, &o, 1);
        }
        put(out, rowIds.data(), rowIds.size());
    }
    if (light)
    {
        std::fclose(out);
        std::cout.rdbuf(old);
        return 0;
    }
    // ---- Som::trainSingle + Som::addBmu by hand (the caller-side loop of src/Som.cpp:1161-1171) on a copy of the map:
    //      the hit must be counted once per sample; global-BMU and local-BMU (sigma == 1) regimes
    {
        Som copy(som);
        size_t last = 0;
        for (int r = 0; r < std::min(n, 12); ++r)
        {
            Eigen::VectorXf v(Din);
            std::memcpy(v.data(), x.data() + static_cast<size_t>(r) * Din, sizeof(float) * Din);
            const auto ret = copy.trainSingle(v, ones, ones, transform == 2 ? 0.005 : 0.2, r < 8 ? 2.0 : 1.0, last, static_cast<Som::WeigthDecayFunction>(decay == 2 ? 0 : decay));
            copy.addBmu(ret.bmu);
            const uint64_t bi = copy.getIndex(ret.bmu), lb = last;
            put(out, &bi, 1);
            put(out, &lb, 1);
            put(out, &ret.distanceError, 1);
            put(out, ret.residual.data(), static_cast<size_t>(ret.residual.size()));
        }
        for (size_t v : copy.getBmuHits())
        {
            const uint64_t u = v;
            put(out, &u, 1);
        }
        const Eigen::VectorXf m0 = copy.getNeuron(size_t{0}), w2 = copy.getWeigthMap();
        put(out, m0.data(), static_cast<size_t>(m0.size()));
        put(out, w2.data(), static_cast<size_t>(w2.size()));
    }
    // ---- findLocalBmu from several start nodes (column 0 / row 0 wrap quirks included), findRestrictedBmd
    {
        const size_t starts[6] = {0, static_cast<size_t>(W - 1), N - 1, static_cast<size_t>(H - 1) * W, N / 2, static_cast<size_t>(W)};
        for (int r = 0; r < std::min(n, 6); ++r)
        {
            Eigen::VectorXf v(Din);
            std::memcpy(v.data(), x.data() + static_cast<size_t>(r) * Din, sizeof(float) * Din);
            for (size_t st : starts)
            {
                const uint64_t li = som.getIndex(som.findLocalBmu(v, ones, st, ones));
                put(out, &li, 1);
            }
            const std::vector<double> bmd = som.findRestrictedBmd(v, ones, r % 3, ones);
            put(out, bmd.data(), bmd.size());
        }
    }
    for (int r = 0; r < std::min(n, 16); ++r)
    {
        Eigen::VectorXf v(Din);
        std::memcpy(v.data(), x.data() + static_cast<size_t>(r) * Din, sizeof(float) * Din);
        const SomIndex b = som.findBmu(v, ones, ones);
        const SomIndex rb = som.findRestrictedBmu(v, ones, 2, ones);
        const uint64_t bi = som.getIndex(b), rbi = som.getIndex(rb);
        const double d = som.euclidianWeightedDist(b, v, ones, ones);
        put(out, &bi, 1);
        put(out, &rbi, 1);
        put(out, &d, 1);
    }
    // persistence: the checkpoint text itself, and a map rebuilt from it
    {
        const std::string ckpt = std::string(argv[2]) + ".som";
        som.save(ckpt.c_str());
        FILE *ck = std::fopen(ckpt.c_str(), "rb");
        std::vector<char> text;
        char buf[4096];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof(buf), ck)) > 0)
            text.insert(text.end(), buf, buf + got);
        std::fclose(ck);
        const uint64_t bytes = text.size();
        put(out, &bytes, 1);
        put(out, text.data(), text.size());
        if (transform == 0) // the file constructor always installs the default (Standard) transformation
        {
            Som again(ckpt.c_str());
            const uint64_t shape[3] = {again.getWidth(), again.getHeight(), again.getDepth()};
            put(out, shape, 3);
            for (size_t p = 0; p < N; ++p)
            {
                const Eigen::VectorXf m = again.getNeuron(p), s = again.getSigmaNeuron(p);
                put(out, m.data(), static_cast<size_t>(m.size()));
                put(out, s.data(), static_cast<size_t>(s.size()));
            }
            const Eigen::VectorXf w2 = again.getWeigthMap();
            put(out, w2.data(), static_cast<size_t>(w2.size()));
            for (size_t v : again.getBmuHits())
            {
                const uint64_t u = v;
                put(out, &u, 1);
            }
            const UMatrix u2 = again.getUMatrix();
            put(out, u2.getData().data(), u2.getData().size());
            const double e2 = again.evaluate(ds2); // scoring on the reloaded (6-decimal) map
            put(out, &e2, 1);
        }
    }
    const double nw = Som::calculateNeighbourhoodWeight(size_t{3}, size_t{1}, size_t{1}, size_t{2}, 2.0);
    put(out, &nw, 1);
    std::fclose(out);
    std::cout.rdbuf(old);
    return 0;
}

// mnist_loader_test.cpp — drives MnistDataLoader + DataSet (host code only, no GPU): prints, per load() of one pass over the
// stream and of the first load of the next pass, the row count, the stream flag and a checksum of the chunk, then column
// metadata.  tests/test_mnist_loader.py writes small IDX files and checks the output against numpy.
// usage: mnist_loader_test <dir> <maxLoadCount>
#include "DataSet.hpp"
#include "MnistDataLoader.hpp"

#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv)
{
    if (argc != 3)
        return 2;
    const size_t maxLoad = static_cast<size_t>(std::atoll(argv[2]));
    MnistDataLoader loader(maxLoad ? std::optional<size_t>(maxLoad) : std::nullopt);
    loader.open(argv[1]);
    DataSet ds(loader);
    std::printf("depth %zu name0 %s name783 %s name784 %s name793 %s cols %zu\n", ds.vectorLength(), ds.getName(0).c_str(), ds.getName(783).c_str(),
                ds.getName(784).c_str(), ds.getName(793).c_str(), loader.findAllColumns().size());
    int loads = 0;
    while (!ds.hasReadWholeDataStream() && loads < 64)
    {
        ds.loadNextDataFromStream();
        double sum = 0, wsum = 0;
        const float *rows = ds.contiguousRows();
        for (size_t r = 0; r < ds.size(); ++r)
            for (size_t k = 0; k < ds.vectorLength(); ++k)
            {
                sum += rows[r * ds.vectorLength() + k];
                wsum += static_cast<double>((r + 1) * (k + 1) % 1000) * rows[r * ds.vectorLength() + k];
            }
        int valid = 1;
        for (size_t r = 0; r < ds.size(); ++r)
            valid = valid && ds.getValidity(r).size() == 794 && ds.getValidity(r)[5] == 1 && ds.getData(r)[0] == rows[r * 794];
        std::printf("load %d rows %zu atStart %d sum %.1f wsum %.1f valid %d\n", loads, ds.size(), loader.isAtStartOfDataStream() ? 1 : 0, sum, wsum, valid);
        ++loads;
    }
    ds.resetStreamLoadPosition();
    ds.loadNextDataFromStream();
    std::printf("next pass rows %zu\n", ds.size());
    std::printf("preview %zu weight %.1f binary %d continuous %d spec %zu\n", ds.getPreviewData(3).size(), ds.getWeight(790), loader.getBinary()[0], loader.getContinuous()[793],
                loader.getColumnSpec().size());
    return 0;
}

// MnistDataLoader.hpp — the reference's MNIST row source (include/MnistDataLoader.hpp:10-51, src/MnistDataLoader.cpp:9-137)
// as a loader that INSTANTIATES: the reference's class leaves findAllColumns / setColumnSpec / getColumnSpec pure and reads
// the files through a third-party header (extern/mnistReader, not vendored), so it can neither be constructed nor built
// from the reference tree alone.  Same behaviour otherwise:
//   * open(path) names the directory that holds train-images-idx3-ubyte / train-labels-idx1-ubyte (IDX format);
//   * load() produces up to maxLoadCount rows (all when unset / 0) starting at the stream position, 28*28 pixel columns as
//     un-normalised floats 0..255 followed by the 10 one-hot label columns (src/MnistDataLoader.cpp:67-75), every column
//     valid; the stream position advances by the rows produced and wraps to 0 when a load comes back empty or holds the
//     whole training set (>= 60000 rows), exactly the rule of src/MnistDataLoader.cpp:49-53;
//   * names "XxY" for the pixels and "label:K" for the one-hot columns (:118-137); all weights 1, all columns continuous.
#pragma once

#include "IDataLoader.hpp"

#include <optional>
#include <string>
#include <vector>

class MnistDataLoader : public IDataLoader
{
  protected:
    std::vector<float> _weights;
    std::vector<int> _isBinary;
    std::vector<int> _isContinuous;
    std::vector<std::string> _names;
    std::string _filePath;
    bool _verbose;

    void generateNames();
    // rows [first, first + limit) of the training files (limit 0: to the end) as RowData, 794 columns each
    std::vector<RowData> readRows(size_t first, size_t limit) const;

  public:
    MnistDataLoader(std::optional<size_t> maxLoadCount = std::nullopt, bool verbose = false)
        : IDataLoader{maxLoadCount}, _weights(28 * 28 + 10, 1.0f), _isBinary(28 * 28 + 10, 0), _isContinuous(28 * 28 + 10, 1), _names{}, _filePath{},
          _verbose{verbose}
    {
        generateNames();
    }

    size_t load() override;
    std::vector<RowData> getPreview(size_t count) override;
    bool open(const char *path) override;
    std::vector<std::string> findAllColumns() override { return _names; }
    void setColumnSpec(const std::vector<ColumnSpec> columnSpec) noexcept override;
    const std::vector<ColumnSpec> getColumnSpec() noexcept override;
    float getWeight(size_t index) override;
    const std::vector<float> getWeights() const noexcept override;
    const std::vector<int> &getBinary() const noexcept override;
    const std::vector<int> &getContinuous() const noexcept override;
    std::string getName(size_t index) const noexcept override;
    const std::vector<std::string> getNames() const noexcept override;
    size_t getDepth() const noexcept override;
    bool isAtStartOfDataStream() const noexcept override { return m_currentIndex == 0; }
};

// DataSet.hpp — chunk-streaming façade over an IDataLoader, API of the reference (include/DataSet.hpp:10-63,
// src/DataSet.cpp:8-176): loadNextDataFromStream() pulls one chunk, hasReadWholeDataStream() /
// resetStreamLoadPosition() drive the per-epoch loop of Som::trainBasicSom (src/Som.cpp:1155-1181), and every
// chunk load zeroes the per-row lastBMU (src/DataSet.cpp:136-137).  Sample order is loader order.
//
// Added for the device path: the chunk is also kept as ONE contiguous row-major float buffer
// (contiguousRows()), which is what crosses the C-ABI — the per-row Eigen vectors remain for API parity.
#pragma once

#include "IDataLoader.hpp"

#include <string>
#include <vector>

#include "Eigen/Dense"

class DataSet
{
  protected:
    struct DataRow
    {
        Eigen::VectorXf *data;
        std::vector<int> *valid;
        size_t *lastBMU;
    };
    std::vector<DataRow> allData;
    std::vector<Eigen::VectorXf> data;
    std::vector<std::vector<int>> valid;
    std::vector<size_t> index;
    std::vector<size_t> lastBMU;
    std::vector<float> packed; // n x depth, row-major: the staging view handed to the device
    IDataLoader &_loader;
    size_t depth, n, loadedNumberOfChunks;
    bool _verbose;

  public:
    DataSet(IDataLoader &dataLoader, bool verbose = false) : _loader{dataLoader}, depth{}, n{}, loadedNumberOfChunks{0}, _verbose{verbose} {}
    ~DataSet() = default;
    const std::vector<DataRow> getAll() const;
    std::vector<DataRow> getAll();
    std::vector<Eigen::VectorXf> getPreviewData(size_t count) const;
    Eigen::VectorXf getData(size_t index) const;
    const Eigen::VectorXi getValidity(size_t index) const;
    const Eigen::ArrayXi getBinary() const;
    const Eigen::ArrayXi getContinuous() const;
    const Eigen::VectorXf getWeights() const;
    float getWeight(size_t index);
    const std::vector<std::string> getNames() const noexcept;
    std::string getName(size_t) const;
    const std::vector<size_t> &getLastBMU() const noexcept;
    size_t &getLastBMU(size_t);
    size_t size() const;
    void addVector(Eigen::VectorXf);
    void display() const;
    void loadNextDataFromStream();
    size_t vectorLength() const;
    bool hasReadWholeDataStream() const noexcept;
    void resetStreamLoadPosition() noexcept;

    // device staging view of the loaded chunk (size() x vectorLength() floats) and its lastBMU column
    const float *contiguousRows() const noexcept { return packed.data(); }
    std::vector<size_t> &lastBmuColumn() noexcept { return lastBMU; }
};

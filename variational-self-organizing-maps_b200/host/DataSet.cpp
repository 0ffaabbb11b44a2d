// DataSet.cpp — host chunk streaming (reference src/DataSet.cpp:8-176) plus the contiguous staging view.
#include "DataSet.hpp"

#include <cassert>
#include <cstring>
#include <iostream>
#include <numeric>

const std::vector<DataSet::DataRow> DataSet::getAll() const { return allData; }
std::vector<DataSet::DataRow> DataSet::getAll() { return allData; }

std::vector<Eigen::VectorXf> DataSet::getPreviewData(size_t count) const
{
    std::vector<Eigen::VectorXf> out;
    for (auto &row : _loader.getPreview(count))
        out.push_back(row.values);
    return out;
}

// rows beyond the loaded chunk read as zeros (reference src/DataSet.cpp:34-48)
Eigen::VectorXf DataSet::getData(size_t at) const { return at < _loader.data.size() ? data[at] : Eigen::VectorXf::Zero(_loader.getDepth()); }
const Eigen::VectorXi DataSet::getValidity(size_t at) const
{
    if (at < n)
        return Eigen::Map<const Eigen::VectorXi>(valid[at].data(), valid[at].size());
    return Eigen::VectorXi::Zero(_loader.getDepth());
}
const Eigen::ArrayXi DataSet::getBinary() const { return Eigen::Map<const Eigen::ArrayXi>(_loader.getBinary().data(), _loader.getBinary().size()); }
const Eigen::ArrayXi DataSet::getContinuous() const
{
    return Eigen::Map<const Eigen::ArrayXi>(_loader.getContinuous().data(), _loader.getContinuous().size());
}
const Eigen::VectorXf DataSet::getWeights() const
{
    const std::vector<float> w = _loader.getWeights();
    Eigen::VectorXf out(w.size());
    std::memcpy(out.data(), w.data(), w.size() * sizeof(float));
    return out;
}
float DataSet::getWeight(size_t at) { return _loader.getWeight(at); }
const std::vector<std::string> DataSet::getNames() const noexcept { return _loader.getNames(); }
std::string DataSet::getName(size_t at) const { return _loader.getName(at); }
const std::vector<size_t> &DataSet::getLastBMU() const noexcept { return lastBMU; }
size_t &DataSet::getLastBMU(size_t at)
{
    assert(at < n);
    return lastBMU[at];
}
size_t DataSet::size() const { return n; }

void DataSet::addVector(Eigen::VectorXf v)
{
    if (static_cast<size_t>(v.rows()) != _loader.getDepth())
    {
        std::cout << "Added vector size does not correspond to data set depth!\n";
        return;
    }
    packed.insert(packed.end(), v.data(), v.data() + v.size());
    data.push_back(v);
    n += 1;
}

void DataSet::display() const { std::cout << "Number of samples: " << _loader.data.size() << "\nVector length: " << _loader.getDepth() << "\n"; }
size_t DataSet::vectorLength() const { return _loader.getDepth(); }
void DataSet::resetStreamLoadPosition() noexcept { loadedNumberOfChunks = 0; }
bool DataSet::hasReadWholeDataStream() const noexcept { return loadedNumberOfChunks > 0 && _loader.isAtStartOfDataStream(); }

void DataSet::loadNextDataFromStream()
{
    if (_loader.isAtStartOfDataStream())
        loadedNumberOfChunks = 0;
    n = _loader.load();
    depth = _loader.getDepth();
    data.clear();
    valid.clear();
    allData.clear();
    data.reserve(n);
    valid.reserve(n);
    allData.reserve(n);
    lastBMU.assign(n, 0); // every chunk load restarts the local BMU search at node 0 (reference src/DataSet.cpp:136-137)
    index.resize(n);
    std::iota(index.begin(), index.end(), size_t{0}); // the reference shuffles this and never reads it (:143-146): order = loader order
    packed.resize(n * depth);
    size_t r = 0;
    for (auto &row : _loader.data)
    {
        data.push_back(row.values);
        valid.push_back(row.valid);
        allData.push_back(DataRow{&data.back(), &valid.back(), &lastBMU[r]});
        std::memcpy(packed.data() + r * depth, row.values.data(), depth * sizeof(float));
        ++r;
    }
    ++loadedNumberOfChunks;
    if (_verbose)
        std::cout << "Loaded " << loadedNumberOfChunks << " number of chunks\n";
}

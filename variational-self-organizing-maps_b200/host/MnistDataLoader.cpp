// MnistDataLoader.cpp — IDX reader behind the reference's MNIST loader interface (see include/MnistDataLoader.hpp).
#include "MnistDataLoader.hpp"

#include <cstdint>
#include <cstdio>
#include <iostream>

namespace
{
// big-endian 32-bit word of an IDX header
bool readWord(std::FILE *f, uint32_t &out)
{
    unsigned char b[4];
    if (std::fread(b, 1, 4, f) != 4)
        return false;
    out = (static_cast<uint32_t>(b[0]) << 24) | (static_cast<uint32_t>(b[1]) << 16) | (static_cast<uint32_t>(b[2]) << 8) | b[3];
    return true;
}
struct File
{
    std::FILE *f;
    explicit File(const std::string &path) : f{std::fopen(path.c_str(), "rb")} {}
    ~File()
    {
        if (f)
            std::fclose(f);
    }
};
} // namespace

bool MnistDataLoader::open(const char *path)
{
    _filePath = path;
    return true;
}

std::vector<RowData> MnistDataLoader::readRows(size_t first, size_t limit) const
{
    std::vector<RowData> out;
    File images(_filePath + "/train-images-idx3-ubyte"), labels(_filePath + "/train-labels-idx1-ubyte");
    uint32_t magicI = 0, countI = 0, rowsI = 0, colsI = 0, magicL = 0, countL = 0;
    if (!images.f || !labels.f || !readWord(images.f, magicI) || !readWord(images.f, countI) || !readWord(images.f, rowsI) || !readWord(images.f, colsI) ||
        !readWord(labels.f, magicL) || !readWord(labels.f, countL) || magicI != 0x803 || magicL != 0x801 || rowsI != 28 || colsI != 28)
    {
        if (_verbose)
            std::cout << "MnistDataLoader: cannot read the IDX training files under " << _filePath << "\n";
        return out;
    }
    const size_t total = countI < countL ? countI : countL;
    if (first >= total)
        return out;
    const size_t n = (limit == 0 || first + limit > total) ? total - first : limit;
    std::vector<unsigned char> pix(n * 784), lab(n);
    if (std::fseek(images.f, static_cast<long>(16 + first * 784), SEEK_SET) != 0 || std::fread(pix.data(), 1, pix.size(), images.f) != pix.size() ||
        std::fseek(labels.f, static_cast<long>(8 + first), SEEK_SET) != 0 || std::fread(lab.data(), 1, lab.size(), labels.f) != lab.size())
        return out;
    out.reserve(n);
    for (size_t r = 0; r < n; ++r)
    {
        RowData row{Eigen::VectorXf(28 * 28 + 10), std::vector<int>(28 * 28 + 10, 1)};
        for (size_t k = 0; k < 784; ++k)
            row.values[static_cast<Eigen::Index>(k)] = static_cast<float>(pix[r * 784 + k]);
        for (size_t k = 0; k < 10; ++k) // one-hot label behind the image (src/MnistDataLoader.cpp:67-75)
            row.values[static_cast<Eigen::Index>(784 + k)] = lab[r] == k ? 1.0f : 0.0f;
        out.push_back(std::move(row));
    }
    return out;
}

std::vector<RowData> MnistDataLoader::getPreview(size_t count) { return readRows(0, count); }

size_t MnistDataLoader::load()
{
    data = readRows(m_currentIndex, m_maxLoadCount.value_or(0));
    const size_t numberOfSamples = data.size();
    // src/MnistDataLoader.cpp:49-53
    if (numberOfSamples == 0)
        m_currentIndex = 0;
    else if (numberOfSamples >= 60000)
        m_currentIndex = 0;
    else
        m_currentIndex += numberOfSamples;
    return numberOfSamples;
}

void MnistDataLoader::setColumnSpec(const std::vector<ColumnSpec> columnSpec) noexcept
{
    for (size_t c = 0; c < columnSpec.size() && c < _names.size(); ++c)
    {
        _names[c] = columnSpec[c].name;
        _weights[c] = columnSpec[c].weight;
        _isBinary[c] = columnSpec[c].isBinary;
        _isContinuous[c] = !columnSpec[c].isBinary;
    }
}

const std::vector<ColumnSpec> MnistDataLoader::getColumnSpec() noexcept
{
    std::vector<ColumnSpec> out;
    for (size_t c = 0; c < _names.size(); ++c)
        out.emplace_back(_names[c], _weights[c], _isBinary[c]);
    return out;
}

float MnistDataLoader::getWeight(size_t index) { return _weights.at(index); }
const std::vector<float> MnistDataLoader::getWeights() const noexcept { return _weights; }
const std::vector<int> &MnistDataLoader::getBinary() const noexcept { return _isBinary; }
const std::vector<int> &MnistDataLoader::getContinuous() const noexcept { return _isContinuous; }
std::string MnistDataLoader::getName(size_t index) const noexcept { return _names.at(index); }
const std::vector<std::string> MnistDataLoader::getNames() const noexcept { return _names; }
size_t MnistDataLoader::getDepth() const noexcept { return _names.size(); }

void MnistDataLoader::generateNames()
{
    _names.reserve(28 * 28 + 10);
    for (size_t index = 0; index < 28 * 28; ++index)
        _names.emplace_back(std::to_string(index % 28) + "x" + std::to_string(index / 28)); // :124-131
    for (size_t index = 0; index < 10; ++index)
        _names.emplace_back("label:" + std::to_string(index)); // :132-136
}

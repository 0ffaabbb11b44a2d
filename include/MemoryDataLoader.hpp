// MemoryDataLoader.hpp — an in-memory IDataLoader over a caller-owned row-major float buffer.
// Not in the reference (its loaders are SQLite / MNIST file readers, out of scope here); this is the loader
// the benchmarks and parity tests feed synthetic data through, serving `chunkRows` rows per load() in buffer
// order and wrapping to the start after the last chunk, which is the protocol DataSet expects
// (reference src/DataSet.cpp:113-121).
#pragma once

#include "IDataLoader.hpp"

#include <algorithm>
#include <cstring>

class MemoryDataLoader : public IDataLoader
{
  public:
    MemoryDataLoader(const float *rows, size_t rowCount, size_t depth, size_t chunkRows = 0, const int *validMask = nullptr)
        : _rows{rows}, _mask{validMask}, _rowCount{rowCount}, _depth{depth}, _chunk{chunkRows ? chunkRows : rowCount}, _weights(depth, 1.0f),
          _binary(depth, 0), _continuous(depth, 1), _names(depth)
    {
        for (size_t c = 0; c < depth; ++c)
            _names[c] = "c" + std::to_string(c);
    }

    size_t load() override
    {
        const size_t first = m_currentIndex, last = std::min(first + _chunk, _rowCount);
        data.clear();
        data.reserve(last - first);
        for (size_t r = first; r < last; ++r)
        {
            RowData row{Eigen::VectorXf(_depth), std::vector<int>(_depth, 1)};
            std::memcpy(row.values.data(), _rows + r * _depth, _depth * sizeof(float));
            if (_mask)
                std::copy(_mask + r * _depth, _mask + (r + 1) * _depth, row.valid.begin());
            data.push_back(std::move(row));
        }
        _chunkFirstRow = first;
        m_currentIndex = last >= _rowCount ? 0 : last;
        return data.size();
    }
    std::vector<RowData> getPreview(size_t count) override
    {
        std::vector<RowData> out;
        for (size_t r = 0; r < std::min(count, _rowCount); ++r)
        {
            RowData row{Eigen::VectorXf(_depth), std::vector<int>(_depth, 1)};
            std::memcpy(row.values.data(), _rows + r * _depth, _depth * sizeof(float));
            out.push_back(std::move(row));
        }
        return out;
    }
    bool open(const char *) override { return true; }
    std::vector<std::string> findAllColumns() override { return _names; }
    void setColumnSpec(const std::vector<ColumnSpec> spec) noexcept override
    {
        for (size_t c = 0; c < spec.size() && c < _depth; ++c)
        {
            _names[c] = spec[c].name;
            _weights[c] = spec[c].weight;
            _binary[c] = spec[c].isBinary;
            _continuous[c] = !spec[c].isBinary;
        }
    }
    const std::vector<ColumnSpec> getColumnSpec() noexcept override
    {
        std::vector<ColumnSpec> out;
        for (size_t c = 0; c < _depth; ++c)
            out.emplace_back(_names[c], _weights[c], _binary[c]);
        return out;
    }
    float getWeight(size_t index) override { return _weights[index]; }
    const std::vector<float> getWeights() const noexcept override { return _weights; }
    const std::vector<int> &getBinary() const noexcept override { return _binary; }
    const std::vector<int> &getContinuous() const noexcept override { return _continuous; }
    std::string getName(size_t index) const noexcept override { return _names[index]; }
    const std::vector<std::string> getNames() const noexcept override { return _names; }
    size_t getDepth() const noexcept override { return _depth; }
    bool isAtStartOfDataStream() const noexcept override { return m_currentIndex == 0; }

    // zero-copy fast path: the rows of the chunk produced by the last load(), contiguous and row-major
    const float *chunkRows() const noexcept { return _rows + _chunkFirstRow * _depth; }

  private:
    const float *_rows;
    const int *_mask;
    size_t _rowCount, _depth, _chunk, _chunkFirstRow{0};
    std::vector<float> _weights;
    std::vector<int> _binary, _continuous;
    std::vector<std::string> _names;
};

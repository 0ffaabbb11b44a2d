// UMatrix.cpp — accessors of the U-matrix value wrapper (reference src/UMatrix.cpp:6-43).
#include "UMatrix.hpp"

#include <algorithm>
#include <cassert>

double UMatrix::getValueAtIndex(size_t column, size_t row) const
{
    const size_t at = row * width + column;
    assert(at <= data.size());
    return data[at];
}
double UMatrix::getValueAtIndex(SomIndex position) const { return getValueAtIndex(position.getX(), position.getY()); }
const std::vector<double> &UMatrix::getData() const noexcept { return data; }
size_t UMatrix::getWidth() const noexcept { return width; }
size_t UMatrix::getHeight() const noexcept { return height; }
double UMatrix::getMinValue() const noexcept { return *std::min_element(data.begin(), data.end()); }
double UMatrix::getMaxValue() const noexcept { return *std::max_element(data.begin(), data.end()); }

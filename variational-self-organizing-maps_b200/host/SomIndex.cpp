// SomIndex.cpp — host value type; behaviour of reference src/SomIndex.cpp:10-45 (incl. the /height quirk).
#include "SomIndex.hpp"
#include "SOM.hpp"

SomIndex::SomIndex(size_t ix, size_t iy) noexcept : x{ix}, y{iy} {}

SomIndex::SomIndex(const Som &map, size_t index) noexcept
    : x{index % map.getWidth()}, y{(index - index % map.getWidth()) / map.getHeight()} // sic: reference src/SomIndex.cpp:15-18
{
}

size_t SomIndex::getSomIndex(const Som &som) { return y * som.getWidth() + x; }
size_t SomIndex::getX() const noexcept { return x; }
size_t SomIndex::getY() const noexcept { return y; }
void SomIndex::setX(size_t index) noexcept { x = index; }
void SomIndex::setY(size_t index) noexcept { y = index; }

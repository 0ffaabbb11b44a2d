// SomIndex.hpp — (x, y) position of a node; value type of the reference (include/SomIndex.hpp:7-27,
// src/SomIndex.cpp:10-45).  Linear index = y * width + x.
#pragma once

#include <stddef.h>

class Som;

class SomIndex
{
  protected:
    size_t x, y;

  public:
    SomIndex(size_t x, size_t y) noexcept;
    // NB: the reference divides by the map HEIGHT here (src/SomIndex.cpp:15-18), which is only right for
    // square maps; kept for parity.
    SomIndex(const Som &map, size_t index) noexcept;
    ~SomIndex() = default;
    size_t getSomIndex(const Som &som);
    size_t getX() const noexcept;
    size_t getY() const noexcept;
    void setX(size_t index) noexcept;
    void setY(size_t index) noexcept;

    bool operator==(const SomIndex &other) const { return x == other.x && y == other.y; }
};

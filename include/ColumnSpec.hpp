// ColumnSpec.hpp — per-column metadata of a data source, same shape as the reference's
// include/ColumnSpec.hpp:5-17 (name, weight, isBinary; all const, construct only from lvalues).
#pragma once

#include <string>

struct ColumnSpec
{
    ColumnSpec(const std::string columnName, const float columnWeight, const int binaryFlag)
        : name{columnName}, weight{columnWeight}, isBinary{binaryFlag}
    {
    }
    // the reference forbids building a spec from temporaries (include/ColumnSpec.hpp:11-12); keep that contract
    ColumnSpec(const std::string &&, const float &&, const int &&) = delete;
    ColumnSpec(std::string &&, float &&, int &&) = delete;

    const std::string name;
    const float weight;
    const int isBinary;
};

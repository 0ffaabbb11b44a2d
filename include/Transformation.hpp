// Transformation.hpp — the injected strategy object of the reference (include/Transformation.hpp:10-41):
// Comparer / Stepper / Displayer / Length as std::function fields plus a Name, and the three shipped
// factories (src/Transformation.cpp:3-39, :41-77, :79-167).
//
// Device note: a std::function cannot run inside a CUDA kernel.  Each factory therefore also stamps `Kind`,
// which selects the compile-time functor of the same arithmetic in the sm_100a kernels
// (vsom_transform_kind in vsom_b200.h).  A hand-built Transformation (Kind == Custom and a Name that is none
// of the shipped ones) is rejected by Som's device entry points with an error — there is no CPU fallback.
#pragma once

#include "Eigen/Dense"

#include <functional>
#include <sstream>
#include <string>
#include <vector>

struct Transformation
{
    using T = Eigen::VectorXf;
    enum class DeviceKind
    {
        Custom = -1,
        Standard = 0,
        MedianEstimator = 1,
        LinearRegression = 2
    };

    std::function<T(const T &value, const T &model, const T &dispersion, const T &valueWeight)> Comparer{
        [](const T &value, const T &model, const T &, const T &) { return model - value; }};
    std::function<T(const T &value, const T &model, const T &valueWeight)> Stepper{
        [](const T &value, const T &model, const T &) { return value - model; }};
    std::vector<std::string> names{};
    // (the reference's default Displayer captures `names` by reference and dangles after a copy,
    //  SURVEY.md App. B.12; this one captures nothing and reads the argument's length)
    std::function<std::vector<std::string>(const T &model)> Displayer{[](const T &model) {
        std::vector<std::string> out;
        for (Eigen::Index i = 0; i < model.size(); ++i)
        {
            std::stringstream ss;
            ss << model[i];
            out.push_back(ss.str());
        }
        return out;
    }};
    std::function<size_t(size_t vectorLength)> Length{[](size_t vectorLength) { return vectorLength; }};
    std::string Name{"Standard transformation"};
    DeviceKind Kind{DeviceKind::Standard};

    static Transformation Standard(const std::vector<std::string> &columnNames);
    static Transformation StandardMedianEstimator(const std::vector<std::string> &columnNames);
    static Transformation CombinatorialLinearRegression(const std::vector<std::string> &columnNames);

    // Kind, or the kind recognised from Name for objects built by code that does not know about Kind.
    DeviceKind deviceKind() const;
};

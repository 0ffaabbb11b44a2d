// Transformation.cpp — host-side factories of the three shipped strategies (reference
// src/Transformation.cpp:3-39, :41-77, :79-167).  The std::function bodies serve host callers
// (Som::trainSingle's residual, display helpers); the device kernels run the same arithmetic as
// compile-time functors selected by Kind (csrc/common.cuh residual<>, csrc/online_step.cu).
#include "Transformation.hpp"

namespace
{
using T = Transformation::T;

std::function<std::vector<std::string>(const T &)> plainDisplayer(const std::vector<std::string> &columns)
{
    return [columns](const T &model) {
        std::vector<std::string> lines;
        lines.reserve(columns.size());
        for (size_t c = 0; c < columns.size(); ++c)
        {
            std::stringstream ss;
            ss << columns[c] << " = " << model[static_cast<Eigen::Index>(c)];
            lines.push_back(ss.str());
        }
        return lines;
    };
}

// pair expansion of the CLR model (src/Transformation.cpp:94-101): q enumerates i<j row-major,
// x'_q = v_i, y'_q = v_j; model = [A | B]; returns inner_q = A_q x'_q + B_q - y'_q
T clrInner(const T &value, const T &model, T *xPrimeOut)
{
    const Eigen::Index pairs = model.size() / 2;
    T inner(pairs), xPrime(pairs);
    Eigen::Index q = 0;
    for (Eigen::Index i = 0; i < value.size(); ++i)
        for (Eigen::Index j = i + 1; j < value.size(); ++j, ++q)
        {
            xPrime[q] = value[i];
            inner[q] = (model[q] * value[i] + model[pairs + q]) - value[j];
        }
    if (xPrimeOut)
        *xPrimeOut = xPrime;
    return inner;
}
} // namespace

Transformation Transformation::Standard(const std::vector<std::string> &columnNames)
{
    Transformation t;
    t.names = columnNames;
    t.Displayer = plainDisplayer(columnNames);
    t.Name = "Standard transformation";
    t.Kind = DeviceKind::Standard;
    return t;
}

Transformation Transformation::StandardMedianEstimator(const std::vector<std::string> &columnNames)
{
    Transformation t;
    t.Stepper = [](const T &value, const T &model, const T &) {
        T step(model.size());
        for (Eigen::Index k = 0; k < model.size(); ++k)
        {
            const float d = value[k] - model[k];
            step[k] = (d != d) ? d : static_cast<float>((0.0f < d) - (d < 0.0f));
        }
        return step;
    };
    t.names = columnNames;
    t.Displayer = plainDisplayer(columnNames);
    t.Name = "Standard median estimator transformation";
    t.Kind = DeviceKind::MedianEstimator;
    return t;
}

Transformation Transformation::CombinatorialLinearRegression(const std::vector<std::string> &columnNames)
{
    Transformation t;
    t.Comparer = [](const T &value, const T &model, const T &, const T &) { return clrInner(value, model, nullptr); };
    t.Stepper = [](const T &value, const T &model, const T &) {
        T xPrime;
        const T inner = clrInner(value, model, &xPrime);
        const Eigen::Index pairs = inner.size();
        T delta(model.size());
        for (Eigen::Index q = 0; q < pairs; ++q)
        {
            const float m2 = -2.0f * inner[q]; // (-2 * inner) * x'  (src/Transformation.cpp:129-130)
            delta[q] = m2 * xPrime[q];
            delta[pairs + q] = m2;
        }
        return delta;
    };
    t.names = columnNames;
    t.Displayer = [columnNames](const T &model) {
        std::vector<std::string> lines;
        const Eigen::Index pairs = model.size() / 2;
        Eigen::Index q = 0;
        for (size_t i = 0; i < columnNames.size() && q < pairs; ++i)
            for (size_t j = i + 1; j < columnNames.size(); ++j, ++q)
            {
                std::stringstream ss;
                ss << columnNames[i] << " = " << model[q] << '*' << columnNames[j] << " + " << model[pairs + q];
                lines.push_back(ss.str());
            }
        return lines;
    };
    t.Length = [](size_t vectorLength) { return vectorLength * (vectorLength - 1u); };
    t.Name = "Linear regression";
    t.Kind = DeviceKind::LinearRegression;
    return t;
}

Transformation::DeviceKind Transformation::deviceKind() const
{
    if (Kind != DeviceKind::Custom)
    {
        // a default-constructed Transformation says Standard; trust Name when it names another shipped strategy
        if (Name == "Standard median estimator transformation")
            return DeviceKind::MedianEstimator;
        if (Name == "Linear regression")
            return DeviceKind::LinearRegression;
        return Kind;
    }
    if (Name == "Standard transformation")
        return DeviceKind::Standard;
    if (Name == "Standard median estimator transformation")
        return DeviceKind::MedianEstimator;
    if (Name == "Linear regression")
        return DeviceKind::LinearRegression;
    return DeviceKind::Custom;
}
